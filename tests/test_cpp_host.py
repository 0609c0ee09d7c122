"""The C++ host side above the C ABI: include/nbvh_b200.hpp (header-only mirror of NeuralBVHModule / NeuralPathModule,
/root/reference/include/cuda/neural/neural_bvh_module.h:144-241) and nbvh_b200/host/nbvh_render_cli.cpp.

  * CPU: the header compiles stand-alone as C++17 with -Wall -Werror, the command line host links against libnbvh_b200.so, and
    without a device it fails loudly (exit code 1, a message) instead of falling back to anything.
  * GPU: the frame the C++ host writes for the shipped Chess files is byte-identical to the frame the Python host gets through the
    same C ABI (same config JSON, .nbvh, network_models JSON, scene arrays)."""
import json
import os
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(REPO, "nbvh_b200", "host", "nbvh_render_cli")


def _cli():
    from nbvh_b200 import build as b
    b.build()
    assert os.path.exists(CLI)
    return CLI


def test_header_compiles_standalone(tmp_path):
    src = tmp_path / "tu.cpp"
    src.write_text('#include "nbvh_b200.hpp"\nint main() { return sizeof(nbvh_b200::NeuralPathModule) == sizeof(void *) ? 0 : 1; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I" + os.path.join(REPO, "include"), str(src)])


def test_cli_links_and_fails_loudly_without_inputs(tmp_path):
    cli = _cli()
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    r = subprocess.run([cli, str(tmp_path), "nothing", str(tmp_path / "o.ppm")], capture_output=True, text=True)
    assert r.returncode == 1 and "nbvh_render_cli:" in r.stderr and not (tmp_path / "o.ppm").exists()


def _read_ppm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P6"
        w, h = map(int, f.readline().split())
        assert f.readline().strip() == b"255"
        return np.frombuffer(f.read(), np.uint8).reshape(h, w, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("spp", [1, 3])
def test_cli_frame_equals_python_host_frame(tmp_path, spp):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nbvh_b200 import scene_io
    from nbvh_b200._capi import CONST_SCENE
    root = scene_io.chess_fixture_dir()
    if root is None:
        pytest.skip("Chess fixture missing")
    W, H = 256, 144
    out = tmp_path / "cli.ppm"
    r = subprocess.run([_cli(), root, scene_io.CHESS_BASE, str(out), str(W), str(H), str(spp), "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["W"] == W and info["H"] == H and info["tree_cut_nodes"] > 20000 and info["rays"] >= W * H * spp
    cli_img = _read_ppm(out)

    mod = scene_io.load_chess_module(max_depth=1, envmap_eval=True)
    sc = mod.get_constants(CONST_SCENE)
    sc.skylike_background = True
    mod.set_constants(CONST_SCENE, sc)
    mod.reinitialise_inference_data((W, H))
    ldr = np.zeros((H, W, 4), np.uint8)
    for s in range(spp - 1):
        mod.neural_render(s, 0)
    mod.neural_render_host(spp - 1, 0, ldr)
    mod.close()
    assert ldr[..., :3].std() > 5                 # a real picture, not a constant frame
    assert (cli_img == ldr[..., :3]).all()


@pytest.mark.gpu
def test_cli_with_device_built_tlas_renders_the_same_frame(tmp_path):
    """Without the TLAS files the C++ host builds the TLAS on the device (nbvh_build_tlas): same frame, byte for byte."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nbvh_b200 import scene_io
    root = scene_io.chess_fixture_dir()
    if root is None:
        pytest.skip("Chess fixture missing")
    view = tmp_path / "scene"
    view.mkdir()
    for name in os.listdir(root):
        if not name.startswith("tlas_") and os.path.isfile(os.path.join(root, name)):
            os.symlink(os.path.join(root, name), view / name)
    imgs = []
    for d in (root, str(view)):
        out = tmp_path / ("a.ppm" if d == root else "b.ppm")
        r = subprocess.run([_cli(), d, scene_io.CHESS_BASE, str(out), "320", "180", "1", "1"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        imgs.append(_read_ppm(out))
    assert imgs[0].std() > 5 and (imgs[0] == imgs[1]).all()


@pytest.mark.gpu
def test_cli_exr_is_accumulation_over_spp(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nbvh_b200 import results, scene_io
    from nbvh_b200._capi import CONST_SCENE
    root = scene_io.chess_fixture_dir()
    if root is None:
        pytest.skip("Chess fixture missing")
    W, H, spp = 128, 72, 2
    out = tmp_path / "cli.exr"
    r = subprocess.run([_cli(), root, scene_io.CHESS_BASE, str(out), str(W), str(H), str(spp), "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    img = results.read_exr(str(out))
    mod = scene_io.load_chess_module(max_depth=1, envmap_eval=True)
    sc = mod.get_constants(CONST_SCENE)
    sc.skylike_background = True
    mod.set_constants(CONST_SCENE, sc)
    accum, _ = mod.attach_framebuffer((W, H))
    for s in range(spp):
        mod.neural_render(s, 0)
    torch.cuda.synchronize()
    want = accum.cpu().numpy() / np.float32(spp)
    mod.close()
    assert img.shape == (H, W, 4) and np.isfinite(img).all() and img[..., :3].std() > 0
    np.testing.assert_allclose(img, want, rtol=0, atol=1e-6)   # same kernels; fp32 atomics free (one sample per pixel per launch)


def test_train_cli_links_and_fails_loudly_without_inputs(tmp_path):
    _cli()
    cli = os.path.join(REPO, "nbvh_b200", "host", "nbvh_train_cli")
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    r = subprocess.run([cli, str(tmp_path), "nothing", str(tmp_path / "out")], capture_output=True, text=True)
    assert r.returncode == 1 and "nbvh_train_cli:" in r.stderr


@pytest.mark.gpu
def test_train_cli_matches_python_validation_run(tmp_path):
    """The C++ validation-mode host (optimisation with tree-cut learning, frames, result files) against the Python host on the same
    scene files: same tree-cut growth, the same first loss (same RNG keys and initial weights), loss curves within 2 % (the gradient
    atomics make runs differ in the last bits), the reference's file layout, and close images."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import nbvh_b200
    from nbvh_b200 import results, scene_io, validation
    root = scene_io.chess_fixture_dir()
    if root is None or not scene_io.has_shading_fixture(root):
        pytest.skip("Chess fixture missing")
    _cli()
    base, W, H, steps, spp = scene_io.CHESS_BASE, 160, 90, 45, 2
    out = tmp_path / "cpp"
    r = subprocess.run([os.path.join(REPO, "nbvh_b200", "host", "nbvh_train_cli"), root, base, str(out), str(steps), str(spp), str(W), str(H)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    for rel in ("results/%s_cpp_learning_spp_0.csv", "results/%s_cpp_spp_2.exr", "results/%s_cpp_inference_spp_2.csv", "nbvhs/%s.nbvh",
                "network_models/%s_network.json", "configs/%s_cpp.json"):
        assert (out / (rel % base)).exists(), rel
    rows = (out / ("results/%s_cpp_learning_spp_0.csv" % base)).read_text().splitlines()
    assert rows[0] == "frame index,frame time (ms), loss" and len(rows) == steps + 1
    cpp_losses = np.array([float(x.split(",")[2]) for x in rows[1:]])

    cfg = json.load(open(os.path.join(root, base + "_neural.json")))
    cfg["cuda_scene"]["fb_size"] = [W, H]
    mod = nbvh_b200.NeuralPathModule(json.dumps(cfg))
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root, with_neural_mesh=True, with_attributes=True)
    mats, textures, _ = nbvh_b200.load_shading_arrays(root, with_envmap=False)
    mod.set_scene(tl, ti, bl, materials=mats, textures=textures)
    mod.max_depth = 1
    mod.attach_framebuffer((W, H))
    res = validation.run_full_validation(mod, cfg, str(tmp_path / "py"), max_training_steps=steps, max_accumulated_spp=spp, base=base, variant="py")
    assert info["tree_cut_nodes"] == res["tree_cut_nodes"] > 1
    py_losses = np.array(res["losses"])
    assert abs(cpp_losses[0] - py_losses[0]) <= 2e-5 * py_losses[0]
    assert np.abs(cpp_losses - py_losses).max() <= 0.02 * py_losses.max()
    a = results.read_exr(str(out / ("results/%s_cpp_spp_2.exr" % base)))
    b = results.read_exr(res["paths"]["image"])
    assert a.shape == b.shape == (H, W, 4) and np.abs(a - b).mean() < 0.02
    # the C++ host's files load back: its tree-cut and weights render through the Python host
    mod.load_neural_bvh(str(out / ("nbvhs/%s.nbvh" % base)))
    mod.load_network_model(str(out / ("network_models/%s_network.json" % base)))
    assert len(mod.neural_bvh_nodes()) == info["tree_cut_nodes"]
    mod.close()
