"""result writers and the configuration round trip.

CPU (host-only entry points of the C ABI):
  * nbvh_write_exr_host reproduces BYTE FOR BYTE the file the reference's own tinyexr writes when driven like write_exr
    (include/base/bitmap.h:97-182) -- tests/golden/ref_exr_7x5.exr from oracle/ref_tools/exr_golden.cpp on the .rgba input next to it;
    when the tool is present (authoring container) a larger image is also decoded back by tinyexr's LoadEXR
  * the timing CSV has the layout of FrameTimer::write_to_file (include/utils/frame_timer.h:33-58)
  * the result file names follow save_image / save_performance_metrics (neural_bvh_renderer_init.cu:328-372)
GPU:
  * nbvh_save_image_exr == accumulation / spp; refused under ema_accumulate
  * nbvh_load_config_json(nbvh_save_config_json(x)) restores every constant block bit for bit, and the saved text keeps the reference's keys."""
import json
import os
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, "tests", "golden")
TOOL = os.path.join(REPO, "oracle", "_ref", "exr_golden")


def test_exr_bytes_equal_reference_tinyexr(tmp_path):
    from nbvh_b200 import results
    rgba = np.fromfile(os.path.join(GOLD, "ref_exr_7x5.rgba"), "<f4").reshape(5, 7, 4)
    out = tmp_path / "o.exr"
    results.write_exr(str(out), rgba, 16.0)
    assert out.read_bytes() == open(os.path.join(GOLD, "ref_exr_7x5.exr"), "rb").read()
    back = results.read_exr(str(out))
    assert (back.view(np.uint32) == (rgba / np.float32(16.0)).view(np.uint32)).all()


def test_exr_rejects_bad_arguments(tmp_path):
    from nbvh_b200 import _capi, results
    lib = _capi.load_library()
    assert lib.nbvh_write_exr_host(None, 4, 4, None, 1.0) == -1
    with pytest.raises(OSError):
        results.write_exr(str(tmp_path / "no_such_dir" / "x.exr"), np.zeros((2, 2, 4), np.float32))
    with pytest.raises(ValueError):
        results.write_exr(str(tmp_path / "x.exr"), np.zeros((2, 2, 3), np.float32))


@pytest.mark.skipif(not os.path.exists(TOOL), reason="oracle/_ref/exr_golden not built (authoring container only)")
def test_exr_decodes_with_reference_tinyexr(tmp_path):
    from nbvh_b200 import results
    rng = np.random.default_rng(5)
    rgba = rng.uniform(-2, 50, (37, 61, 4)).astype(np.float32)
    out = tmp_path / "o.exr"
    results.write_exr(str(out), rgba, 3.0)
    r = subprocess.run([TOOL, "load", str(out), str(tmp_path / "back.rgba")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.split() == ["61", "37"]
    back = np.fromfile(tmp_path / "back.rgba", "<f4").reshape(37, 61, 4)
    assert (back.view(np.uint32) == (rgba / np.float32(3.0)).view(np.uint32)).all()


def test_frame_timer_csv_layout(tmp_path):
    from nbvh_b200 import results
    t = results.FrameTimer()
    t.start_frame(0)
    t.end_frame()
    t.end_frame()
    t.frame_times = [1.5, 0.123456789, 1234567.0]
    p = tmp_path / "a.csv"
    t.write_to_file(str(p))
    assert p.read_text() == "frame index,frame time (ms)\n0,1.5\n1,0.123457\n2,1.23457e+06\n"      # default ostream formatting
    t.write_to_file(str(p), losses=[0.25, 3.0, 1e-5])
    assert p.read_text() == "frame index,frame time (ms), loss\n0,1.5,0.25\n1,0.123457,3\n2,1.23457e+06,1e-05\n"
    with pytest.raises(ValueError):
        t.write_to_file(str(p), losses=[1.0])
    t.start_frame(0)
    assert t.frame_times == []


def test_result_paths():
    from nbvh_b200 import results
    assert results.result_path("/s/chess", "nodes_1", "neural", "image", 64) == "/s/chess/results/nodes_1_neural_spp_64.exr"
    assert results.result_path("/s/chess", "nodes_1", "neural", "learning", 0) == "/s/chess/results/nodes_1_neural_learning_spp_0.csv"
    assert results.result_path("/s/chess", "nodes_1", "neural", "inference", 8) == "/s/chess/results/nodes_1_neural_inference_spp_8.csv"


@pytest.mark.gpu
def test_save_image_is_accumulation_over_spp(tmp_path, synthetic):
    import torch
    import nbvh_b200
    from nbvh_b200 import results
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg, nbvh, tlas, tlas_idx, blas = synthetic
    mod = nbvh_b200.NeuralPathModule(cfg)
    mod.set_neural_bvh(nbvh)
    mod.set_scene(tlas, tlas_idx, blas)
    mod.init_params(1337)
    w, h = cfg["cuda_scene"]["fb_size"]
    accum, _ = mod.attach_framebuffer((w, h))
    for s in range(3):
        mod.neural_render(s, 0)
    torch.cuda.synchronize()
    p = tmp_path / "img.exr"
    mod.save_image(str(p), 3)
    img = results.read_exr(str(p))
    want = accum.cpu().numpy() / np.float32(3.0)
    assert img.shape == (h, w, 4) and want.std() > 0
    assert (img.view(np.uint32) == want.view(np.uint32)).all()
    sc = mod.get_constants(nbvh_b200.CONST_SCENE)
    sc.ema_accumulate = True
    mod.set_constants(nbvh_b200.CONST_SCENE, sc)
    with pytest.raises(nbvh_b200.NbvhError):
        mod.save_image(str(p), 3)
    mod.close()


def _blocks(mod):
    import nbvh_b200
    import ctypes as C
    out = {}
    for which in (nbvh_b200.CONST_RENDERER, nbvh_b200.CONST_MODULE, nbvh_b200.CONST_PT_OUTPUT, nbvh_b200.CONST_PATH, nbvh_b200.CONST_SCENE,
                  nbvh_b200.CONST_CAMERA, nbvh_b200.CONST_NETWORK, nbvh_b200.CONST_OPTIMIZER):
        pod = mod.get_constants(which)
        out[which] = bytes(C.string_at(C.addressof(pod), C.sizeof(pod)))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("source", ["synthetic", "chess"])
def test_save_config_round_trip(source, synthetic):
    import torch
    import nbvh_b200
    from nbvh_b200 import scene_io
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if source == "chess":
        root = scene_io.chess_fixture_dir()
        if root is None:
            pytest.skip("Chess fixture missing")
        cfg = open(os.path.join(root, scene_io.CHESS_BASE + "_neural.json")).read()
    else:
        cfg = synthetic[0]
    a = nbvh_b200.NeuralPathModule(cfg)
    a.max_depth = 5
    text = a.save_config()
    saved = json.loads(text)
    ref_cfg = json.loads(cfg) if isinstance(cfg, str) else cfg
    # the saved text keeps the reference's structure: every key it writes exists in the source config at the same place
    for sect in ("pt_output_constants", "path_module_constants", "bvh_module_constants"):
        assert set(saved["neural_bvh_renderer"]["neural_module"][sect]) >= set(ref_cfg["neural_bvh_renderer"]["neural_module"].get(sect, {}))
    assert set(saved["neural_bvh_renderer"]["bvh_constants"]) >= set(ref_cfg["neural_bvh_renderer"].get("bvh_constants", {}))
    assert saved["cuda_scene"]["fb_size"] == list(ref_cfg["cuda_scene"]["fb_size"])
    assert saved["neural_bvh_renderer"]["neural_module"]["max_depth"] == 5
    b = nbvh_b200.NeuralPathModule(text)
    assert b.max_depth == 5 and b.n_params() == a.n_params()
    ba, bb = _blocks(a), _blocks(b)
    for k in ba:
        assert ba[k] == bb[k], "constant block %d differs after the round trip" % k
    assert b.save_config() == text
    a.close()
    b.close()


@pytest.mark.gpu
def test_full_validation_run_writes_reference_layout_and_reloads(tmp_path, chess_root):
    """ValidationMode::Full, headless and short: 40 optimisation steps with tree-cut learning, 3 frames; the files have the reference's
    names and formats, and a fresh module created from the SAVED config + .nbvh + network JSON reproduces the frame bit for bit."""
    import torch
    import nbvh_b200
    from nbvh_b200 import results, validation
    from nbvh_b200.scene_io import CHESS_BASE
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    W, H = 160, 90
    cfg = json.load(open(os.path.join(chess_root, CHESS_BASE + "_neural.json")))
    cfg["cuda_scene"]["fb_size"] = [W, H]
    mod = nbvh_b200.NeuralPathModule(json.dumps(cfg))
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(chess_root, with_neural_mesh=True, with_attributes=True)
    mats, texs = nbvh_b200.procedural_materials(int(max(b["material_ids"].max() for b in bl)) + 1)
    mod.set_scene(tl, ti, bl, materials=mats, textures=texs)
    mod.max_depth = 1
    accum, _ = mod.attach_framebuffer((W, H))
    res = validation.run_full_validation(mod, cfg, str(tmp_path), max_training_steps=40, max_accumulated_spp=3, base=CHESS_BASE, variant="t")
    p = res["paths"]
    assert os.path.relpath(p["image"], tmp_path) == "results/%s_t_spp_3.exr" % CHESS_BASE
    assert os.path.relpath(p["learning_csv"], tmp_path) == "results/%s_t_learning_spp_0.csv" % CHESS_BASE
    assert os.path.relpath(p["inference_csv"], tmp_path) == "results/%s_t_inference_spp_3.csv" % CHESS_BASE
    lines = open(p["learning_csv"]).read().splitlines()
    assert lines[0] == "frame index,frame time (ms), loss" and len(lines) == 41
    assert abs(float(lines[1].split(",")[2]) - res["losses"][0]) <= 1e-5 * abs(res["losses"][0])
    assert len(open(p["inference_csv"]).read().splitlines()) == 4
    assert np.isfinite(res["losses"]).all() and np.mean(res["losses"][-10:]) < 0.8 * max(res["losses"])   # the first splits raise the loss, then it falls
    assert res["tree_cut_nodes"] > 1                               # the schedule split the root-only cut
    img = results.read_exr(p["image"])
    torch.cuda.synchronize()
    assert (img.view(np.uint32) == (accum.cpu().numpy() / np.float32(3.0)).view(np.uint32)).all()
    saved = json.load(open(p["config"]))
    for k in ("batch_size", "tree_cut_learning", "bvh_split_scheduler", "max_training_steps"):
        assert saved["neural_bvh_renderer"][k] == cfg["neural_bvh_renderer"][k]
    assert saved["neural_bvh_renderer"]["training_step"] == 40
    mod.neural_render(0, 0)
    torch.cuda.synchronize()
    first = accum.cpu().numpy().copy()
    mod.close()
    mod2 = nbvh_b200.NeuralPathModule(open(p["config"]).read())
    mod2.load_neural_bvh(p["nbvh"])
    mod2.load_network_model(p["network"])
    mod2.set_scene(tl, ti, bl, materials=mats, textures=texs)
    assert mod2.max_depth == 1 and len(mod2.neural_bvh_nodes()) == res["tree_cut_nodes"]
    accum2, _ = mod2.attach_framebuffer((W, H))
    mod2.neural_render(0, 0)
    torch.cuda.synchronize()
    assert first.std() > 0 and (accum2.cpu().numpy().view(np.uint32) == first.view(np.uint32)).all()
    mod2.close()
