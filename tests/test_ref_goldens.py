"""Parity against the reference's OWN kernels, run headless on a B200 (tests/golden/ref_chess_*.npz).

The goldens were produced by tools/make_ref_render_goldens.py driving oracle/ref_tools/ref_render_harness.cu: the
reference's device functions (camera_raygen, NeuralBVH2::intersect, CudaBVH2::intersect<false>), its wavefront
kernels (path_inference_kernels.cu, neural_bvh_module_kernels.cu) and its tiny-cuda-nn inference, compiled from
/root/reference for sm_100, on the shipped Chess scene / config / tree-cut / weights.

  * `-m "not gpu"`: the CPU oracle against those goldens -> pins the oracle by execution of the reference
  * `-m gpu`: the CUDA path (through the C ABI) against the same goldens

Tolerances: bit-exact (u32 compare) for rays, interval lists, node ids, encoded segment inputs and classical hits.
Network outputs: tcnn accumulates in fp16, nbvh_b200 / the oracle's mode 0 in fp32 -> max-abs NET_TOL_TCNN on the
8 post-sigmoid outputs.  Hit records: kinds may flip where visibility is within that tolerance of the 0.5 threshold."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G_SMALL = os.path.join(HERE, "golden", "ref_chess_256x144.npz")
G_FULL = os.path.join(HERE, "golden", "ref_chess_1080p.npz")
MAX_IV = 32
NET_TOL_TCNN = 2e-2      # SURVEY 8(d) parity gate: max-abs on post-sigmoid outputs vs tcnn (fp16 accumulate)
# The north star states the tolerance on LOGITS: saturated sigmoid outputs hide pre-activation error.  tcnn hands out fp16 post-sigmoid
# values, so logits are recovered as log(y / (1 - y)) where tcnn's y lies in [0.05, 0.95] (there one fp16 ulp of y is <= 0.011 in logit
# space); max-abs logit error of the fp32-accumulate network against tcnn's fp16-accumulate one:
LOGIT_TOL_TCNN = 0.05  # measured 0.026 (oracle and CUDA path alike: same fp32-accumulate arithmetic)
LOGIT_RANGE = (0.05, 0.95)


def logit_error(y, ref):
    """max-abs and 99.9th percentile |logit(y) - logit(ref)| over the outputs whose reference value is away from saturation."""
    y, ref = np.asarray(y, np.float64), np.asarray(ref, np.float64)
    m = (ref >= LOGIT_RANGE[0]) & (ref <= LOGIT_RANGE[1])
    yc = np.clip(y[m], 1e-6, 1.0 - 1e-6)
    e = np.abs(np.log(yc / (1.0 - yc)) - np.log(ref[m] / (1.0 - ref[m])))
    return float(e.max()), float(np.percentile(e, 99.9)), int(m.sum())


FLIP_RATE = 5e-3         # hit-kind flips allowed (visibility within NET_TOL_TCNN of the 0.5 threshold)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def gold():
    if not os.path.exists(G_SMALL):
        pytest.skip("reference goldens not generated yet (tools/make_ref_render_goldens.py on a GPU box)")
    g = dict(np.load(G_SMALL))
    g["n"] = int(g["W"]) * int(g["H"])
    cnt = np.minimum(g["iv_count"].astype(np.int64), MAX_IV)
    g["mask"] = np.arange(MAX_IV)[None, :] < cnt[:, None]
    g["first"] = np.concatenate([[0], np.cumsum(cnt)[:-1]])  # offset of each pixel's first interval in the ragged arrays
    return g


@pytest.fixture(scope="module")
def chess_cfg(chess_root, gold):
    from nbvh_b200.scene_io import CHESS_BASE
    cfg = json.load(open(os.path.join(chess_root, CHESS_BASE + "_neural.json")))
    cfg["cuda_scene"]["scene_constants"]["envmap_eval"] = False
    cfg["cuda_scene"]["fb_size"] = [int(gold["W"]), int(gold["H"])]
    return cfg


@pytest.fixture(scope="module")
def cpu_env(oracle, chess_root, chess_cfg):
    from nbvh_b200.scene_io import CHESS_BASE
    nbvh = oracle.read_nbvh(os.path.join(chess_root, CHESS_BASE + ".nbvh"))
    params = oracle.read_network_json(os.path.join(chess_root, CHESS_BASE + "_network.json"))
    scene, _ = oracle.load_chess_fixture(chess_root)
    return dict(nbvh=nbvh, const=oracle.config_constants(chess_cfg, nbvh), net=oracle.make_net(chess_cfg, params), scene=scene)


def golden_rays(gold):
    r = gold["rays"].view(np.float32)
    return np.ascontiguousarray(r[:, :3]), np.ascontiguousarray(r[:, 3:])


def check_hits(gold, kind, t, uv, prim_bits, label):
    """Final primary hit records against the reference's HybridIntersectionResult (tcnn fp16-accumulate inference)."""
    gk = gold["hit_kind"]
    same = kind == gk
    flips = int((~same).sum())
    assert flips <= max(3, int(FLIP_RATE * len(gk))), "%s: %d hit-kind flips of %d" % (label, flips, len(gk))
    gt = gold["hit_t"].view(np.float32)
    cls = same & (gk == 2)
    assert cls.sum() > 1000
    agree = bits(t[cls]) == gold["hit_t"][cls]
    assert agree.mean() > 0.995, "%s: classical t bit-exact on %.4f" % (label, agree.mean())
    sel = np.flatnonzero(cls)[agree]
    assert (bits(uv[sel, 0]) == gold["hit_a"][sel, 0]).all() and (bits(uv[sel, 1]) == gold["hit_a"][sel, 1]).all()
    assert (prim_bits[sel] == gold["hit_b"][sel, 1]).all()
    neu = same & (gk == 1)
    assert neu.sum() > 1000
    dt = np.abs(t[neu] - gt[neu])
    # depth = clamp(out1) * (len + 1e-3) + mint: an output error of NET_TOL_TCNN moves t by at most NET_TOL_TCNN * len
    assert np.median(dt) < 2e-3, "%s: median |dt| %g" % (label, np.median(dt))
    assert (dt < 5e-2).mean() > 0.99, "%s: %.4f of neural hits within 5e-2" % (label, (dt < 5e-2).mean())
    return flips, float(np.median(dt))


# ------------------------------------------------------------------------------------------------ oracle (CPU)
def test_oracle_raygen_matches_reference_kernel(gold, oracle, cpu_env):
    o, d = oracle.raygen(cpu_env["const"], np.arange(gold["n"]))
    assert (bits(np.concatenate([o, d], axis=1)) == gold["rays"]).all()


def test_oracle_interval_lists_match_reference_kernel(gold, oracle, cpu_env):
    o, d = golden_rays(gold)
    cnt, node, mint, maxt, _ = oracle.trace_intervals(cpu_env["nbvh"], cpu_env["const"], o, d, MAX_IV)
    assert (cnt == gold["iv_count"]).all()
    m = gold["mask"]
    assert m.sum() > 30000
    assert (node[m] == gold["iv_node"]).all()
    assert (bits(mint)[m] == gold["iv_mint"]).all()
    assert (bits(maxt)[m] == gold["iv_maxt"]).all()


def test_oracle_classical_hits_match_reference_kernel(gold, oracle, cpu_env):
    o, d = golden_rays(gold)
    t, u, v, prim, _ = oracle.blas_intersect(cpu_env["scene"], 1, o, d)
    hit = gold["blas_prim"] != 0xFFFFFFFF
    assert hit.sum() > 5000 and ((t < 3e38) == hit).all()
    assert (bits(t)[hit] == gold["blas_t"][hit]).all()
    assert (prim[hit] == gold["blas_prim"][hit]).all()
    assert (bits(u)[hit] == gold["blas_u"][hit]).all() and (bits(v)[hit] == gold["blas_v"][hit]).all()


def test_oracle_segment_encoding_matches_reference_kernel(gold, oracle, cpu_env):
    """First wavefront step: the reference encodes each queued ray's first interval (encode_bvh_network_input_data)."""
    o, d = golden_rays(gold)
    px = gold["s0_px"].astype(np.int64)
    assert len(px) > 5000 and (gold["iv_count"][px] > 0).all()
    first = gold["first"][px]
    mint = gold["iv_mint"].view(np.float32)[first]
    maxt = gold["iv_maxt"].view(np.float32)[first]
    x = oracle.encode_segments(cpu_env["const"], o[px], d[px], mint, maxt)
    assert (bits(x) == gold["s0_in"]).all()


def test_oracle_network_outputs_match_tcnn(gold, oracle, cpu_env):
    x = gold["s0_in"].view(np.float32)
    ref = gold["s0_out"].astype(np.float32)
    y32 = oracle.query(cpu_env["net"], x, mode=0)   # fp16 operands, fp32 accumulate (what the CUDA path computes)
    assert np.abs(y32 - ref).max() < NET_TOL_TCNN
    y16 = oracle.query(cpu_env["net"], x, mode=1)   # tcnn emulation: fp16 accumulate
    assert np.abs(y16 - ref).max() < 4e-3
    assert np.median(np.abs(y16 - ref)) < np.median(np.abs(y32 - ref)) + 1e-6
    mx, p999, n = logit_error(y32, ref)
    print("oracle (fp32 accumulate) logit-space error vs tcnn over %d unsaturated outputs: max %.4f, 99.9th percentile %.4f" % (n, mx, p999))
    assert n > 1000 and mx < LOGIT_TOL_TCNN
    assert logit_error(y16, ref)[0] < 0.03  # the fp16-accumulate emulation: within the fp16 quantisation of tcnn's own outputs


def test_oracle_hit_records_match_reference_wavefront(gold, oracle, cpu_env):
    hits, _ = oracle.render_primary(cpu_env["scene"], cpu_env["nbvh"], cpu_env["const"], cpu_env["net"], 0, gold["n"])
    check_hits(gold, hits["valid_hit"].astype(np.uint8), hits["t"], hits["normal_or_uv"], bits(hits["albedo_or_ids"][:, 1]), "oracle")


# ------------------------------------------------------------------------------------------------ CUDA path (GPU)
@pytest.fixture(scope="module")
def gpu_env(gold, chess_root, chess_cfg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import nbvh_b200
    mod = nbvh_b200.load_chess_module()
    mod.reinitialise_inference_data((int(gold["W"]), int(gold["H"])))
    cam = chess_cfg["cuda_scene"]["camera"]
    mod.set_camera(cam["matrix"], cam["yfov"])
    yield dict(mod=mod, torch=torch)
    mod.close()


@pytest.mark.gpu
def test_gpu_raygen_matches_reference_kernel(gold, gpu_env):
    o, d = gpu_env["mod"].raygen(0, gold["n"])
    got = np.concatenate([o.cpu().numpy(), d.cpu().numpy()], axis=1)
    assert (bits(got) == gold["rays"]).all()


@pytest.mark.gpu
def test_gpu_interval_lists_match_reference_kernel(gold, gpu_env):
    torch = gpu_env["torch"]
    o, d = golden_rays(gold)
    cnt, node, mint, maxt = gpu_env["mod"].trace_intervals(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), MAX_IV)
    assert (cnt.cpu().numpy().astype(np.uint32) == gold["iv_count"]).all()
    m = gold["mask"]
    assert (node.cpu().numpy().astype(np.uint32)[m] == gold["iv_node"]).all()
    assert (bits(mint.cpu().numpy())[m] == gold["iv_mint"]).all()
    assert (bits(maxt.cpu().numpy())[m] == gold["iv_maxt"]).all()


@pytest.mark.gpu
def test_gpu_classical_hits_match_reference_kernel(gold, gpu_env):
    torch = gpu_env["torch"]
    o, d = golden_rays(gold)
    t, u, v, prim = gpu_env["mod"].blas_intersect(1, torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda())
    hit = gold["blas_prim"] != 0xFFFFFFFF
    t = t.cpu().numpy()
    assert ((t < 3e38) == hit).all()
    assert (bits(t)[hit] == gold["blas_t"][hit]).all()
    assert (prim.cpu().numpy().astype(np.uint32)[hit] == gold["blas_prim"][hit]).all()
    assert (bits(u.cpu().numpy())[hit] == gold["blas_u"][hit]).all() and (bits(v.cpu().numpy())[hit] == gold["blas_v"][hit]).all()


@pytest.mark.gpu
def test_gpu_segment_encoding_matches_reference_kernel(gold, gpu_env):
    torch = gpu_env["torch"]
    o, d = golden_rays(gold)
    px = gold["s0_px"].astype(np.int64)
    first = gold["first"][px]
    mint = np.ascontiguousarray(gold["iv_mint"].view(np.float32)[first])
    maxt = np.ascontiguousarray(gold["iv_maxt"].view(np.float32)[first])
    x = gpu_env["mod"].encode_segments(torch.from_numpy(o[px]).cuda(), torch.from_numpy(d[px]).cuda(),
                                       torch.from_numpy(mint).cuda(), torch.from_numpy(maxt).cuda())
    assert (bits(x.cpu().numpy()) == gold["s0_in"]).all()


@pytest.mark.gpu
def test_gpu_network_outputs_match_tcnn(gold, gpu_env):
    torch = gpu_env["torch"]
    x = np.ascontiguousarray(gold["s0_in"].view(np.float32))
    y = gpu_env["mod"].inference(torch.from_numpy(x).cuda()).cpu().numpy()
    err = np.abs(y - gold["s0_out"].astype(np.float32))
    assert err.max() < NET_TOL_TCNN, "max-abs error vs tcnn %g" % err.max()
    mx, p999, n = logit_error(y, gold["s0_out"].astype(np.float32))
    print("logit-space error vs tcnn over %d unsaturated outputs: max %.4f, 99.9th percentile %.4f" % (n, mx, p999))
    assert n > 1000 and mx < LOGIT_TOL_TCNN, "max-abs logit error vs tcnn %g" % mx


@pytest.mark.gpu
def test_gpu_hit_records_match_reference_wavefront(gold, gpu_env):
    import nbvh_b200
    mod, torch = gpu_env["mod"], gpu_env["torch"]
    hits_dev = mod.attach_hit_buffer(gold["n"])
    mod.neural_render(0, 0)
    torch.cuda.synchronize()
    h = hits_dev.cpu().numpy().view(nbvh_b200.HIT_DTYPE)
    check_hits(gold, h["valid_hit"].astype(np.uint8), h["t"], h["normal_or_uv"], bits(h["albedo_or_ids"][:, 1]), "cuda")


@pytest.mark.gpu
def test_gpu_worldt_depth_mode_matches_reference_wavefront(gold, gpu_env):
    """DepthLearningMode::WorldT (neural_bvh_module_common.cuh:58-88) against the reference's wavefront path run with the same
    constant (tools/make_ref_worldt_golden.py): hit kinds, classical t bit-exact, neural t within the network tolerance scaled
    by the scene-box interval the depth output is decoded against."""
    import nbvh_b200
    path = os.path.join(HERE, "golden", "ref_chess_worldt_256x144.npz")
    if not os.path.exists(path):
        pytest.skip("WorldT golden not generated (tools/make_ref_worldt_golden.py on a GPU box)")
    g = np.load(path)
    assert int(g["n"]) == gold["n"]
    mod, torch = gpu_env["mod"], gpu_env["torch"]
    pt = mod.get_constants(nbvh_b200.CONST_PT_OUTPUT)
    saved = mod.get_constants(nbvh_b200.CONST_PT_OUTPUT)
    pt.depth_learning_mode = 1
    mod.set_constants(nbvh_b200.CONST_PT_OUTPUT, pt)
    try:
        hits_dev = mod.attach_hit_buffer(gold["n"])
        mod.neural_render(0, 0)
        torch.cuda.synchronize()
        h = hits_dev.cpu().numpy().view(nbvh_b200.HIT_DTYPE)
    finally:
        mod.set_constants(nbvh_b200.CONST_PT_OUTPUT, saved)
    kind, t = h["valid_hit"].astype(np.uint8), h["t"]
    gk, gt = g["hit_kind"], g["hit_t"].view(np.float32)
    assert (gk == 1).sum() > 1000 and not (gk == gold["hit_kind"]).all()  # the mode changes the frame
    same = kind == gk
    assert (~same).sum() <= int(0.02 * len(gk)), "%d hit-kind flips" % int((~same).sum())
    cls = same & (gk == 2)
    assert (bits(t[cls]) == g["hit_t"][cls]).mean() > 0.99
    neu = same & (gk == 1)
    box = mod.get_constants(nbvh_b200.CONST_SCENE_BBOX)
    diag = float(np.linalg.norm(np.array(box.max[:3]) - np.array(box.min[:3])))
    dt = np.abs(t[neu] - gt[neu])
    assert np.median(dt) < 2e-3 * diag, "median |dt| %g (scene diagonal %g)" % (np.median(dt), diag)
    assert (dt < NET_TOL_TCNN * diag).mean() > 0.99


G_UNIFORM = os.path.join(HERE, "golden", "ref_chess_uniform_256x144.npz")


def uniform_const(oracle, chess_cfg, nbvh):
    cfg = json.loads(json.dumps(chess_cfg))
    cfg["neural_bvh_renderer"]["neural_module"].setdefault("bvh_module_constants", {})["segment_distribution"] = "Uniform"
    return oracle.config_constants(cfg, nbvh)


def test_oracle_uniform_segment_points_match_reference_kernel(gold, oracle, cpu_env, chess_cfg):
    """SegmentDistribution::Uniform on the wavefront path (neural_bvh_module_kernels.cu:53-69,292-297): the reference's first-step
    network inputs (tools/make_ref_uniform_golden.py) are reproduced BIT FOR BIT from each pixel's generator state after ray
    generation -- this pins which generator the draws come from, their order and the blend expression; the final hit records pin
    the draws of the later queries of each chain."""
    if not os.path.exists(G_UNIFORM):
        pytest.skip("Uniform golden not generated (tools/make_ref_uniform_golden.py on a GPU box)")
    g = np.load(G_UNIFORM)
    const = uniform_const(oracle, chess_cfg, cpu_env["nbvh"])
    assert const.segment_distribution == 2
    px = g["s0_px"].astype(np.int64)
    assert (px == gold["s0_px"]).all()  # the same rays queue a first query
    o, d = golden_rays(gold)
    first = gold["first"][px]
    mint = gold["iv_mint"].view(np.float32)[first]
    maxt = gold["iv_maxt"].view(np.float32)[first]
    rng = oracle.raygen_rng(const, px)
    x, rng1 = oracle.encode_segments_rng(const, o[px], d[px], mint, maxt, rng)
    assert (rng1 != rng).all()
    assert (bits(x) == g["s0_in"]).all()
    assert not (g["s0_in"] == gold["s0_in"]).all()
    hits, _ = oracle.render_primary(cpu_env["scene"], cpu_env["nbvh"], const, cpu_env["net"], 0, gold["n"])
    check_uniform_hits(g, hits["valid_hit"].astype(np.uint8), hits["t"], "oracle")


def check_uniform_hits(g, kind, t, label):
    gk, gt = g["hit_kind"], g["hit_t"].view(np.float32)
    same = kind == gk
    assert (gk == 1).sum() > 1000
    assert (~same).sum() <= int(0.02 * len(gk)), "%s: %d hit-kind flips" % (label, int((~same).sum()))
    cls = same & (gk == 2)
    assert (bits(t[cls]) == g["hit_t"][cls]).mean() > 0.99
    neu = same & (gk == 1)
    dt = np.abs(t[neu] - gt[neu])
    assert np.median(dt) < 2e-3, "%s: median |dt| %g" % (label, np.median(dt))
    assert (dt < 5e-2).mean() > 0.98


@pytest.mark.gpu
def test_gpu_uniform_segment_points_match_reference_wavefront(gold, gpu_env, oracle, cpu_env, chess_cfg):
    """The CUDA path with SegmentDistribution::Uniform (k_neural<., true>) against the reference's wavefront run with the same
    constant, and -- tighter -- against the oracle, whose first-step inputs are pinned bit for bit by the CPU test above."""
    import nbvh_b200
    if not os.path.exists(G_UNIFORM):
        pytest.skip("Uniform golden not generated (tools/make_ref_uniform_golden.py on a GPU box)")
    g = np.load(G_UNIFORM)
    mod, torch = gpu_env["mod"], gpu_env["torch"]
    saved = mod.get_constants(nbvh_b200.CONST_MODULE)
    mc = mod.get_constants(nbvh_b200.CONST_MODULE)
    mc.segment_distribution = 2
    mod.set_constants(nbvh_b200.CONST_MODULE, mc)
    try:
        hits_dev = mod.attach_hit_buffer(gold["n"])
        mod.neural_render(0, 0)
        torch.cuda.synchronize()
        h = hits_dev.cpu().numpy().view(nbvh_b200.HIT_DTYPE).copy()
    finally:
        mod.set_constants(nbvh_b200.CONST_MODULE, saved)
    kind, t = h["valid_hit"].astype(np.uint8), h["t"]
    assert not (kind == gold["hit_kind"]).all() or not (bits(t) == gold["hit_t"]).all()  # the mode changes the frame
    check_uniform_hits(g, kind, t, "CUDA")
    const = uniform_const(oracle, chess_cfg, cpu_env["nbvh"])
    oh, _ = oracle.render_primary(cpu_env["scene"], cpu_env["nbvh"], const, cpu_env["net"], 0, gold["n"])
    ok = oh["valid_hit"].astype(np.uint8)
    assert (kind != ok).sum() <= int(2e-3 * len(ok)), "%d kind flips against the oracle" % int((kind != ok).sum())
    neu = (kind == ok) & (ok == 1)
    assert np.median(np.abs(t[neu] - oh["t"][neu])) < 1e-4


@pytest.mark.gpu
def test_gpu_1080p_frame_matches_reference_summaries(chess_root, chess_cfg, gpu_env):
    """Full BASELINE.json size: per-pixel interval counts, per-row checksums of the interval lists, classical-hit row sums
    and hit kinds of the whole 1920x1080 frame against the reference's run."""
    if not os.path.exists(G_FULL):
        pytest.skip("1080p reference summary not generated")
    import nbvh_b200
    from tools.make_ref_render_goldens import row_checksums
    g = np.load(G_FULL)
    W, H = int(g["W"]), int(g["H"])
    mod, torch = gpu_env["mod"], gpu_env["torch"]
    cam = chess_cfg["cuda_scene"]["camera"]
    mod.reinitialise_inference_data((W, H))
    mod.set_camera(cam["matrix"], cam["yfov"])
    try:
        n = W * H
        o, d = mod.raygen(0, n)
        cnt, node, mint, maxt = mod.trace_intervals(o, d, MAX_IV)
        cnt = cnt.cpu().numpy().astype(np.uint32)
        assert (np.minimum(cnt, 255).astype(np.uint8) == g["iv_count"]).all()
        rows = row_checksums(cnt, node.cpu().numpy().astype(np.uint32), mint.cpu().numpy(), maxt.cpu().numpy(), W)
        assert (rows == g["iv_row_checksum"]).all()
        del node, mint, maxt
        t, u, v, prim = mod.blas_intersect(1, o, d)
        t = t.cpu().numpy()
        prim = prim.cpu().numpy().astype(np.uint32)
        prim[t >= 3e38] = 0xFFFFFFFF
        t[t >= 3e38] = np.float32(3.4028234663852886e38)
        assert (prim.astype(np.uint64).reshape(H, W).sum(axis=1) == g["blas_prim_row_sum"]).all()
        assert (bits(t).astype(np.uint64).reshape(H, W).sum(axis=1) == g["blas_t_row_sum"]).all()
        hits_dev = mod.attach_hit_buffer(n)
        mod.neural_render(0, 0)
        torch.cuda.synchronize()
        kind = hits_dev.cpu().numpy().view(nbvh_b200.HIT_DTYPE)["valid_hit"].astype(np.uint8)
        flips = int((kind != g["hit_kind"]).sum())
        assert flips <= int(FLIP_RATE * n), "%d hit-kind flips of %d" % (flips, n)
    finally:
        mod.reinitialise_inference_data((256, 144))
        mod.set_camera(cam["matrix"], cam["yfov"])
