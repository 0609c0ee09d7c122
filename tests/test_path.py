"""Hybrid neural path tracing (max_depth > 1): shading, BSDF sampling, Russian roulette, environment map, re-queue
(reference: path_inference_kernels.cu:224-429, material.cuh, path_module.cu:139-245).

Parity gates
  * rendered images against the reference's OWN path tracer (tests/golden/ref_chess_path_256x144.npz: its kernels incl. the
    shading kernel + tiny-cuda-nn, run headless on a B200 by tools/make_ref_path_goldens.py, shipped Chess scene / config /
    weights / HDRI / chessboard textures): relMSE as the reference defines it (neural_bvh_renderer_kernels.cu:139-142:
    mean over pixels and channels of (ref - x)^2 / (ref + 0.1)) <= RELMSE_1SPP / RELMSE_NSPP.  Every pixel follows the same
    random path in both implementations, so 1-spp images are comparable pixel by pixel; the residual comes from tcnn's
    fp16-accumulate network: the CPU oracle run with its fp16-accumulate emulation (mlp_mode 1) reproduces the reference's
    1-spp image to relMSE 5e-6 (99.8 % of the pixels within 1e-3) -- this pins the shading restatement -- while fp32
    accumulation (the oracle's mode 0 and the CUDA kernels) perturbs the network's normals by ~1e-2, which sends 20 % of
    the glossy neural-hit paths (roughness 0.2) to a different part of the HDRI.  Both are unbiased estimates of the same
    image: the means agree to 2e-4 and the relMSE falls with the sample count.
  * CUDA path against the CPU oracle (same fp32-accumulate network): per-pixel agreement (PIXEL_MATCH of the pixels within
    PIXEL_TOL relative) on the synthetic scene (procedural materials, textures and environment map) and on Chess.
  * primary-hit records of the multi-bounce pipeline are bit-identical to the max_depth = 1 pipeline; pixel-range shards
    reproduce the full frame."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G_PATH = os.path.join(HERE, "golden", "ref_chess_path_256x144.npz")
RELMSE_ORACLE_TCNN = 1e-4  # oracle with tcnn's fp16-accumulate network vs the reference's 1-spp image (measured 4.8e-6)
RELMSE_1SPP = 0.15         # CUDA (fp32 accumulate) vs the reference's image, 1 spp  (measured 0.077: divergent glossy paths)
RELMSE_NSPP = 0.03         # ... 16 spp (measured 0.015)
RELMSE_CONVERGED = 4e-3    # ... 256 spp: Monte-Carlo noise of both images
MEAN_TOL = 0.01            # relative difference of the mean radiance at 16 / 256 spp (measured 2e-4)
PIXEL_TOL = 2e-2           # per pixel: |a - b| <= PIXEL_TOL * (|b| + 0.1) on every channel
PIXEL_MATCH = 0.97         # CUDA vs oracle, synthetic scene with random weights (outputs crowd the visibility threshold)
PIXEL_MATCH_CHESS = 0.999  # CUDA vs oracle on Chess (measured 0.9999); oracle(fp16 acc) vs reference (measured 0.9998)


def relmse(x, ref):
    return float((np.square(ref - x) / (ref + 0.1)).mean())


def pixel_match(a, b):
    ok = (np.abs(a - b) <= PIXEL_TOL * (np.abs(b) + 0.1)).all(axis=-1)
    return float(ok.mean())


def procedural_envmap(w=64, h=32, seed=5):
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(h) / h, np.arange(w) / w, indexing="ij")
    env = np.zeros((h, w, 4), np.float32)
    for ch in range(3):
        env[..., ch] = 0.6 + 0.5 * np.sin(2 * np.pi * (xx * rng.integers(1, 3) + rng.uniform())) * np.cos(np.pi * yy) + 2.0 * (yy < 0.2)
    env[..., 3] = 1.0
    return np.maximum(env, 0.02)


def synthetic_path_inputs(synthetic):
    """Synthetic scene + procedural shading inputs: the ground BLAS gets texture coordinates, two material segments
    (textured base colour / metallic-roughness texture, constant colour) and per-vertex normals."""
    import nbvh_b200
    cfg, nbvh, tlas, tlas_idx, blas = synthetic
    cfg = json.loads(json.dumps(cfg))
    blas = [dict(b) for b in blas]
    g = blas[1]
    nv = len(g["vertices"]) // 3
    v = np.asarray(g["vertices"], np.float32).reshape(nv, 3)
    lo, hi = v.min(axis=0), v.max(axis=0)
    uv = ((v - lo) / np.maximum(hi - lo, 1e-6))[:, [0, 2]] * 3.0
    g["texcoords"] = uv.astype(np.float32).reshape(-1)
    n = np.tile(np.array([0.05, 1.0, 0.03], np.float32), (nv, 1))
    g["normals"] = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32).reshape(-1)
    n_tri = len(g["indices"]) // 3
    g["material_ids"] = np.array([0, 1], np.int32)
    g["material_map"] = np.array([0, n_tri // 2], np.uint32)
    mats, textures = nbvh_b200.procedural_materials(2, tex_size=32, seed=9)
    mats["metallic"][0] = nbvh_b200.scene_io.texture_handle(0, 2)
    mats["roughness"][0] = nbvh_b200.scene_io.texture_handle(0, 1)
    mats["roughness"][1] = 0.4
    mats["specular"][1] = 0.5
    mats["clearcoat"][1] = 0.3
    mats["clearcoat_gloss"][1] = 0.5
    mats["base_emission"][1] = (0.02, 0.01, 0.0)
    sc = cfg.setdefault("cuda_scene", {}).setdefault("scene_constants", {})
    sc.update(envmap_eval=True, skylike_background=False, envmap_rotation=75.0, envmap_strength=1.5, ema_accumulate=False,
              emission_eval=True, force_constant_material=False)
    nm = cfg["neural_bvh_renderer"]["neural_module"]
    nm["max_depth"] = 3
    nm["path_module_constants"] = dict(apply_russian_roulette=True, neural_metallic=0.1, neural_roughness=0.2, neural_specular=0.2,
                                       normal_epsilon_offset=1e-4, output_mode="NeuralPathTracing", rr_start_depth=2)
    return cfg, nbvh, tlas, tlas_idx, blas, mats, textures, procedural_envmap()


def spread_params(n_params, n_mlp, seed):
    rng = np.random.default_rng(seed)
    p = np.zeros(n_params, np.float16)
    p[:n_mlp] = (rng.standard_normal(n_mlp) * 0.2).astype(np.float16)
    p[n_mlp:] = (rng.standard_normal(n_params - n_mlp) * 0.7).astype(np.float16)
    return p.view(np.uint16)


N_MLP = 64 * 96 + 3 * 64 * 64 + 16 * 64


def oracle_synthetic(oracle, inputs, params, max_depth=3, spp=1):
    cfg, nbvh, tlas, tlas_idx, blas, mats, textures, env = inputs
    const = oracle.config_constants(cfg, nbvh)
    net = oracle.make_net(cfg, params)
    scene = oracle.Scene(tlas, tlas_idx, blas)
    path = oracle.PathSetup(cfg, blas, mats, textures, env, max_depth=max_depth)
    n = const.fb_w * const.fb_h
    acc = np.zeros((n, 3), np.float32)
    hits0 = None
    for s in range(spp):
        rad, hits, cnt = oracle.render_path(scene, nbvh, const, net, path, 0, n, sample_idx=s)
        acc += rad
        hits0 = hits if hits0 is None else hits0
    return acc, hits0, cnt


# ------------------------------------------------------------------------------------------------ CPU: oracle sanity
def test_oracle_path_synthetic_is_sane(oracle, synthetic):
    inputs = synthetic_path_inputs(synthetic)
    params = spread_params(int(oracle.expected_n_params(oracle.make_net(inputs[0], np.zeros(1, np.uint16)))[0]), N_MLP, 11)
    rad3, hits, cnt3 = oracle_synthetic(oracle, inputs, params, max_depth=3)
    rad1, _, cnt1 = oracle_synthetic(oracle, inputs, params, max_depth=1)
    assert np.isfinite(rad3).all() and (rad3 >= 0).all()
    miss = hits["valid_hit"] == 0
    assert miss.any() and (~miss).any()
    # camera rays that escape see the environment map in both modes, identically
    assert (rad3[miss] == rad1[miss]).all() and (rad3[miss] > 0).any()
    # at max_depth 1 every hit pixel is black except for emission; with bounces light arrives
    assert rad1[~miss].sum() < rad3[~miss].sum()
    assert cnt3[4] > cnt1[4] == len(rad1)  # bounce rays were traced
    assert rad3.max() <= 100.0 * 1.5 * 4   # envmap clamp x strength x throughput head-room


def test_oracle_matches_reference_path_image(oracle, chess_root):
    import nbvh_b200
    from nbvh_b200.scene_io import CHESS_BASE
    if not os.path.exists(G_PATH) or not nbvh_b200.has_shading_fixture(chess_root):
        pytest.skip("path goldens / shading fixture not present")
    g = np.load(G_PATH)
    W, H = int(g["W"]), int(g["H"])
    cfg = json.load(open(os.path.join(chess_root, CHESS_BASE + "_neural.json")))
    cfg["cuda_scene"]["fb_size"] = [W, H]
    nbvh = oracle.read_nbvh(os.path.join(chess_root, CHESS_BASE + ".nbvh"))
    params = oracle.read_network_json(os.path.join(chess_root, CHESS_BASE + "_network.json"))
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(chess_root, with_attributes=True)
    mats, textures, env = nbvh_b200.load_shading_arrays(chess_root)
    const = oracle.config_constants(cfg, nbvh)
    net = oracle.make_net(cfg, params)
    scene = oracle.Scene(tl, ti, bl)
    path = oracle.PathSetup(cfg, bl, mats, textures, env, max_depth=int(g["max_depth"]))
    rad, _, _ = oracle.render_path(scene, nbvh, const, net, path, 0, W * H, mlp_mode=1)  # tcnn's fp16-accumulate arithmetic
    rad = rad.reshape(H, W, 3)
    err, m = relmse(rad, g["rgb_1spp"]), pixel_match(rad, g["rgb_1spp"])
    assert err <= RELMSE_ORACLE_TCNN and m >= PIXEL_MATCH_CHESS, "oracle vs reference path tracer, 1 spp: relMSE %.4g, match %.4f" % (err, m)
    rad32, _, _ = oracle.render_path(scene, nbvh, const, net, path, 0, W * H, mlp_mode=0)  # the B200 kernels' arithmetic
    assert relmse(rad32.reshape(H, W, 3), g["rgb_1spp"]) <= RELMSE_1SPP
    assert abs(rad32.mean() - g["rgb_1spp"].mean()) <= 0.02 * g["rgb_1spp"].mean()


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def gpu_render(mod, torch, W, H, spp=1, first=0, count=0):
    accum, ldr = mod.attach_framebuffer((W, H))
    for s in range(spp):
        mod.neural_render(s, 0, first, count)
    torch.cuda.synchronize()
    return accum.cpu().numpy()[..., :3].reshape(H, W, 3), ldr.cpu().numpy()


@pytest.fixture(scope="module")
def syn_path(torch_cuda, oracle, synthetic):
    import nbvh_b200
    inputs = synthetic_path_inputs(synthetic)
    cfg, nbvh, tlas, tlas_idx, blas, mats, textures, env = inputs
    mod = nbvh_b200.NeuralPathModule(cfg)
    mod.set_neural_bvh(nbvh)
    mod.set_scene(tlas, tlas_idx, blas, materials=mats, textures=textures)
    mod.set_envmap(env)
    params = spread_params(mod.n_params(), N_MLP, 11)
    mod.load_params(params)
    yield dict(mod=mod, inputs=inputs, params=params, W=cfg["cuda_scene"]["fb_size"][0], H=cfg["cuda_scene"]["fb_size"][1])
    mod.close()


@pytest.mark.gpu
@pytest.mark.parametrize("max_depth", [2, 3, 5])
def test_gpu_path_matches_oracle_synthetic(syn_path, oracle, torch_cuda, max_depth):
    mod, W, H = syn_path["mod"], syn_path["W"], syn_path["H"]
    mod.max_depth = max_depth
    img, _ = gpu_render(mod, torch_cuda, W, H)
    rad, _, _ = oracle_synthetic(oracle, syn_path["inputs"], syn_path["params"], max_depth=max_depth)
    rad = rad.reshape(H, W, 3)
    assert np.isfinite(img).all()
    m = pixel_match(img, rad)
    ok = (np.abs(img - rad) <= PIXEL_TOL * (np.abs(rad) + 0.1)).all(axis=-1)
    print("synthetic max_depth %d: %.4f of the pixels agree with the oracle" % (max_depth, m))
    assert m >= PIXEL_MATCH, "max_depth %d: %.4f of the pixels agree with the oracle" % (max_depth, m)
    assert relmse(img[ok], rad[ok]) <= 1e-5 and abs(img.mean() - rad.mean()) <= 0.05 * rad.mean()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["ignore_neural_blas", "force_constant_material", "dielectric", "anisotropic", "skylike_no_rr", "uniform_points"])
def test_gpu_path_shading_branches_match_oracle(oracle, synthetic, torch_cuda, variant):
    """The remaining branches of the shading kernel against the oracle: the non-neural rendering of the neural mesh
    (ignore_neural_blas: its classical BLAS, unpack_albedo + the module's constant specular parameters), force_constant_material,
    a fully transmissive material (dielectric lobe), an anisotropic + sheen material, sky-like background without roulette."""
    import nbvh_b200
    from nbvh_b200 import _capi as A
    inputs = synthetic_path_inputs(synthetic)
    cfg, nbvh, tlas, tlas_idx, blas, mats, textures, env = inputs
    cfg = json.loads(json.dumps(cfg))
    mats = mats.copy()
    sc = cfg["cuda_scene"]["scene_constants"]
    if variant == "ignore_neural_blas":
        cfg["neural_bvh_renderer"]["bvh_constants"]["ignore_neural_blas"] = True
    elif variant == "force_constant_material":
        sc.update(force_constant_material=True, constant_metallic=0.3, constant_roughness=0.6, constant_specular=0.4)
    elif variant == "dielectric":
        mats["specular_transmission"][:] = 1.0
        mats["ior"][:] = 1.45
    elif variant == "anisotropic":
        mats["anisotropy"][:] = 0.7
        mats["sheen"][:] = 0.5
        mats["sheen_tint"][:] = 0.3
        mats["specular_tint"][:] = 0.4
        mats["roughness"][1] = 0.25
    elif variant == "uniform_points":
        # SegmentDistribution::Uniform: every query draws its segment parameters from the path's generator, which the shading step
        # continues from (neural_bvh_module_kernels.cu:53-69,292-297; path_module.h:111,141)
        cfg["neural_bvh_renderer"]["neural_module"].setdefault("bvh_module_constants", {})["segment_distribution"] = "Uniform"
    elif variant == "skylike_no_rr":
        sc.update(skylike_background=True)
        cfg["neural_bvh_renderer"]["neural_module"]["path_module_constants"]["apply_russian_roulette"] = False
    inputs = (cfg, nbvh, tlas, tlas_idx, blas, mats, textures, env)
    mod = nbvh_b200.NeuralPathModule(cfg)
    mod.set_neural_bvh(nbvh)
    mod.set_scene(tlas, tlas_idx, blas, materials=mats, textures=textures)
    mod.set_envmap(env)
    params = spread_params(mod.n_params(), N_MLP, 11)
    mod.load_params(params)
    mod.max_depth = 3
    assert bool(mod.get_constants(A.CONST_RENDERER).ignore_neural_blas) == (variant == "ignore_neural_blas")
    W, H = cfg["cuda_scene"]["fb_size"]
    img, _ = gpu_render(mod, torch_cuda, W, H)
    rad, hits, _ = oracle_synthetic(oracle, inputs, params, max_depth=3)
    rad = rad.reshape(H, W, 3)
    m = pixel_match(img, rad)
    print("%s: %.4f of the pixels agree with the oracle" % (variant, m))
    assert np.isfinite(img).all() and m >= PIXEL_MATCH
    if variant == "ignore_neural_blas":
        assert (hits["valid_hit"] != 1).all() and (hits["valid_hit"] == 2).sum() > 1000
    if variant == "uniform_points":
        assert mod.get_constants(A.CONST_MODULE).segment_distribution == 2
        cfg0 = json.loads(json.dumps(cfg))
        cfg0["neural_bvh_renderer"]["neural_module"]["bvh_module_constants"]["segment_distribution"] = "EquidistantInterior"
        rad0, _, _ = oracle_synthetic(oracle, (cfg0,) + inputs[1:], params, max_depth=3)
        assert pixel_match(rad0.reshape(H, W, 3), rad) < 0.9  # the option changes the frame: the agreement above is not vacuous
    mod.close()


@pytest.mark.gpu
def test_gpu_path_primary_hits_equal_single_bounce_pipeline(syn_path, torch_cuda):
    import nbvh_b200
    mod, W, H = syn_path["mod"], syn_path["W"], syn_path["H"]
    recs = []
    for depth in (1, 3):
        mod.max_depth = depth
        buf = mod.attach_hit_buffer(W * H)
        gpu_render(mod, torch_cuda, W, H)
        recs.append(buf.cpu().numpy().view(nbvh_b200.HIT_DTYPE).copy())
    a, b = recs
    assert (a["valid_hit"] == b["valid_hit"]).all()
    assert (a["t"].view(np.uint32) == b["t"].view(np.uint32)).all()
    assert (a["normal_or_uv"].view(np.uint32) == b["normal_or_uv"].view(np.uint32)).all()
    assert (a["albedo_or_ids"].view(np.uint32) == b["albedo_or_ids"].view(np.uint32)).all()
    mod._lib.nbvh_set_hit_buffer(mod._h, None)


@pytest.mark.gpu
def test_gpu_path_shards_and_accumulation(syn_path, torch_cuda):
    mod, W, H = syn_path["mod"], syn_path["W"], syn_path["H"]
    mod.max_depth = 3
    full, ldr_full = gpu_render(mod, torch_cuda, W, H, spp=3)
    accum, ldr = mod.attach_framebuffer((W, H))
    half = (W * H // 2 // 32) * 32
    for s in range(3):
        mod.neural_render(s, 0, half, W * H - half)
        mod.neural_render(s, 0, 0, half)
    torch_cuda.cuda.synchronize()
    assert (accum.cpu().numpy()[..., :3].reshape(H, W, 3).view(np.uint32) == full.view(np.uint32)).all()
    assert (ldr.cpu().numpy() == ldr_full).all()


@pytest.mark.gpu
@pytest.mark.parametrize("max_depth", [1, 3])
def test_gpu_interleaved_tile_shards_reproduce_full_frame(syn_path, torch_cuda, max_depth):
    """Multi-GPU tile partition (nbvh_render_interleaved): the union of the shards is bit-identical to the full frame."""
    mod, W, H = syn_path["mod"], syn_path["W"], syn_path["H"]
    mod.max_depth = max_depth
    full, ldr_full = gpu_render(mod, torch_cuda, W, H, spp=2)
    accum, ldr = mod.attach_framebuffer((W, H))
    for s in range(2):
        for shard in (2, 0, 1):
            mod.neural_render_shard(s, 0, shard, 3)
    torch_cuda.cuda.synchronize()
    assert (accum.cpu().numpy()[..., :3].reshape(H, W, 3).view(np.uint32) == full.view(np.uint32)).all()
    assert (ldr.cpu().numpy() == ldr_full).all()


@pytest.mark.gpu
@pytest.mark.parametrize("max_depth", [1, 3])
def test_gpu_queue_drain_order_does_not_change_results(syn_path, torch_cuda, max_depth):
    """nbvh_set_queue_order: draining the neural ray queue by expected chain length (per-pixel history of the last frame) is
    scheduling only -- accumulation buffer, sRGB8 frame, hit records and the work counters are bit-identical to arrival order,
    for every class layout / predictor, on a cold history (first frame) and on a warm one."""
    mod, W, H = syn_path["mod"], syn_path["W"], syn_path["H"]
    mod.max_depth = max_depth
    results = []
    try:
        for limits, flags in (((0, 0, 0, 0), 0), ((4, 3, 2, 2), 1), ((2, 2, 2, 2), 0), ((6, 4, 3, 2), 3), ((5, 4, 3, 2), 5)):
            mod.set_queue_order(limits, flags)
            hits = mod.attach_hit_buffer(W * H)
            mod.reset_counters()
            acc, ldr = gpu_render(mod, torch_cuda, W, H, spp=4)
            c = mod.counters()
            results.append((acc.copy(), ldr.copy(), hits.cpu().numpy().copy(), {k: c[k] for k in ("rays", "queries", "nbvh_steps", "queued_rays")}))
    finally:
        mod.set_queue_order()
        mod.attach_hit_buffer(0)
    assert results[0][3]["queued_rays"] > 0
    for acc, ldr, hits, cnt in results[1:]:
        assert (acc.view(np.uint32) == results[0][0].view(np.uint32)).all()
        assert (ldr == results[0][1]).all()
        assert (hits == results[0][2]).all()
        assert cnt == results[0][3]
    with pytest.raises(Exception):
        mod.set_queue_order((2, 3, 2, 2), 1)  # limits must be descending


@pytest.mark.gpu
@pytest.mark.parametrize("max_depth", [1, 3])
def test_gpu_multi_sample_launch_equals_sample_by_sample(syn_path, torch_cuda, max_depth):
    """nbvh_render_samples: 5 samples per pixel in one launch sequence (and 3 + 2 over two shards) == five nbvh_render calls, up
    to the fp32 order of the per-pixel sum (16-byte reductions instead of read-modify-write); max_depth 1 exercises the
    k_primary / k_neural pair, max_depth 3 the path pipeline."""
    mod, W, H = syn_path["mod"], syn_path["W"], syn_path["H"]
    mod.max_depth = max_depth
    ref, ldr_ref = gpu_render(mod, torch_cuda, W, H, spp=5)
    accum, ldr = mod.attach_framebuffer((W, H))
    mod.neural_render_samples(0, 5)
    torch_cuda.cuda.synchronize()
    a = accum.cpu().numpy()[..., :3]
    assert np.allclose(a, ref, rtol=2e-5, atol=1e-6)
    assert np.abs(ldr.cpu().numpy().astype(int) - ldr_ref.astype(int)).max() <= 1
    accum, ldr = mod.attach_framebuffer((W, H))
    for shard in (1, 0):
        mod.neural_render_samples(0, 3, 0, shard, 2)
        mod.neural_render_samples(3, 2, 0, shard, 2)
    torch_cuda.cuda.synchronize()
    assert np.allclose(accum.cpu().numpy()[..., :3], ref, rtol=2e-5, atol=1e-6)
    assert np.abs(ldr.cpu().numpy().astype(int) - ldr_ref.astype(int)).max() <= 1


@pytest.fixture(scope="module")
def chess_path(torch_cuda, oracle, chess_root):
    import nbvh_b200
    if not os.path.exists(G_PATH) or not nbvh_b200.has_shading_fixture(chess_root):
        pytest.skip("path goldens / shading fixture not present")
    g = np.load(G_PATH)
    mod = nbvh_b200.load_chess_module(max_depth=int(g["max_depth"]), envmap_eval=True, shading=True)
    yield dict(mod=mod, g=g, W=int(g["W"]), H=int(g["H"]))
    mod.close()


@pytest.mark.gpu
def test_gpu_path_image_matches_reference_path_tracer(chess_path, torch_cuda):
    mod, g, W, H = chess_path["mod"], chess_path["g"], chess_path["W"], chess_path["H"]
    img1, _ = gpu_render(mod, torch_cuda, W, H, spp=1)
    e1 = relmse(img1, g["rgb_1spp"])
    spp = int(g["spp"])
    imgn, _ = gpu_render(mod, torch_cuda, W, H, spp=spp)
    en = relmse(imgn / spp, g["rgb_nspp"])
    print("relMSE vs the reference path tracer: 1 spp %.4g, %d spp %.4g; mean radiance %.5f vs %.5f" %
          (e1, spp, en, (imgn / spp).mean(), g["rgb_nspp"].mean()))
    # the library's own metric kernel (compute_and_display_render_error) agrees with the numpy definition
    ref4 = torch_cuda.from_numpy(np.concatenate([g["rgb_nspp"], np.ones((H, W, 1), np.float32)], axis=-1)).cuda().contiguous()
    for mode, fn in (("RelMSEError", lambda x, r: np.square(r - x) / (r + 0.1)), ("MSEError", lambda x, r: np.square(r - x)),
                     ("AbsError", lambda x, r: np.abs(r - x)), ("RelAbsError", lambda x, r: np.abs(r - x) / (r + 0.1))):
        dev_err = mod.image_error(ref4, spp, mode)
        assert abs(dev_err - float(fn(imgn / spp, g["rgb_nspp"]).mean())) <= 1e-4 * max(1e-3, dev_err), mode
    assert e1 <= RELMSE_1SPP and en <= RELMSE_NSPP
    assert abs((imgn / spp).mean() - g["rgb_nspp"].mean()) <= MEAN_TOL * g["rgb_nspp"].mean()
    if "rgb_converged" in g:
        sc = int(g["spp_converged"])
        imgc, _ = gpu_render(mod, torch_cuda, W, H, spp=sc)
        ec = relmse(imgc / sc, g["rgb_converged"])
        print("%d spp: relMSE %.4g, mean radiance %.5f vs %.5f" % (sc, ec, (imgc / sc).mean(), g["rgb_converged"].mean()))
        assert ec <= RELMSE_CONVERGED
        assert abs((imgc / sc).mean() - g["rgb_converged"].mean()) <= MEAN_TOL * g["rgb_converged"].mean()


@pytest.mark.gpu
def test_gpu_path_matches_oracle_chess(chess_path, oracle, chess_root, torch_cuda):
    import nbvh_b200
    from nbvh_b200.scene_io import CHESS_BASE
    mod, g, W, H = chess_path["mod"], chess_path["g"], chess_path["W"], chess_path["H"]
    cfg = json.load(open(os.path.join(chess_root, CHESS_BASE + "_neural.json")))
    cfg["cuda_scene"]["fb_size"] = [W, H]
    nbvh = oracle.read_nbvh(os.path.join(chess_root, CHESS_BASE + ".nbvh"))
    params = oracle.read_network_json(os.path.join(chess_root, CHESS_BASE + "_network.json"))
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(chess_root, with_attributes=True)
    mats, textures, env = nbvh_b200.load_shading_arrays(chess_root)
    path = oracle.PathSetup(cfg, bl, mats, textures, env, max_depth=int(g["max_depth"]))
    rad, _, _ = oracle.render_path(oracle.Scene(tl, ti, bl), nbvh, oracle.config_constants(cfg, nbvh), oracle.make_net(cfg, params), path,
                                   0, W * H)
    img, _ = gpu_render(mod, torch_cuda, W, H, spp=1)
    m = pixel_match(img, rad.reshape(H, W, 3))
    print("Chess 256x144 max_depth 3: %.4f of the pixels agree with the oracle, relMSE %.4g" % (m, relmse(img, rad.reshape(H, W, 3))))
    assert m >= PIXEL_MATCH_CHESS


@pytest.mark.gpu
def test_gpu_path_full_size_properties(chess_path, torch_cuda):
    """BASELINE config 2b at its full size (1920x1080, shipped max_depth 3): size-independent properties -- the union of three
    interleaved tile shards is bit-identical to the full frame, a 4-sample multi-sample launch equals four single-sample frames up
    to the fp32 sum order, every pixel is finite and non-negative, and the mean radiance of the frame matches the mean of the
    reference's 256x144 image of the same view within 3 % (one sample per pixel, 2 M pixels)."""
    mod, g = chess_path["mod"], chess_path["g"]
    W, H = 1920, 1080
    full, ldr_full = gpu_render(mod, torch_cuda, W, H, spp=1)
    assert np.isfinite(full).all() and (full >= 0).all()
    accum, ldr = mod.attach_framebuffer((W, H))
    for shard in (2, 0, 1):
        mod.neural_render_shard(0, 0, shard, 3)
    torch_cuda.cuda.synchronize()
    assert (accum.cpu().numpy()[..., :3].view(np.uint32) == full.view(np.uint32)).all()
    assert (ldr.cpu().numpy() == ldr_full).all()
    ref4, _ = gpu_render(mod, torch_cuda, W, H, spp=4)
    accum, ldr = mod.attach_framebuffer((W, H))
    mod.neural_render_samples(0, 4)
    torch_cuda.cuda.synchronize()
    assert np.allclose(accum.cpu().numpy()[..., :3], ref4, rtol=2e-5, atol=1e-6)
    assert abs(full.mean() - g["rgb_converged"].mean()) <= 0.03 * g["rgb_converged"].mean()
    mod.reinitialise_inference_data((chess_path["W"], chess_path["H"]))


# ------------------------------------------------------------------------------------------------ config 4 at its own size
RELMSE_4K_BOX = 1.5e-3    # relMSE (the reference's definition) of the 8x8-box-filtered 16-spp images; measured 5.4e-4
FLIP_4K_BOX = 0.025       # mean LDR-FLIP of the box-filtered images (nbvh_flip_images); measured 0.0132
MEAN_4K = 0.01            # relative difference of the mean radiance; measured 1e-4


@pytest.mark.gpu
def test_gpu_path_config4_matches_reference_at_full_size(chess_root, torch_cuda):
    """BASELINE config 4 at the size bench.py times (3840x2160, max_depth 4, 16 spp, shipped Chess scene + HDRI) against the image of
    the reference's OWN path tracer at that size (tests/golden/ref_chess_path_4k_16spp.npz, tools/make_ref_path4k_golden.py: the
    reference's kernels + tcnn rebuilt for sm_100, 8x8 box filter -> 480x270, every stored pixel = 64 pixels x 16 samples).  The two
    renderers draw the same random numbers but the fp32-accumulate network moves normals by ~1e-2, so the bound is
    statistical: relMSE <= RELMSE_4K_BOX, mean LDR-FLIP <= FLIP_4K_BOX, mean radiance within 1 %."""
    import nbvh_b200
    path = os.path.join(HERE, "golden", "ref_chess_path_4k_16spp.npz")
    if not os.path.exists(path):
        pytest.skip("4K golden not generated (tools/make_ref_path4k_golden.py on a GPU box)")
    g = np.load(path)
    W, H, spp, box = int(g["W"]), int(g["H"]), int(g["spp"]), int(g["box"])
    ref = g["rgb_box"]
    mod = nbvh_b200.load_chess_module(max_depth=int(g["max_depth"]), envmap_eval=True, shading=True)
    accum, _ = mod.attach_framebuffer((W, H))
    mod.neural_render_samples(0, spp)  # one launch sequence for the 16 samples, as bench.py's config 4
    torch_cuda.cuda.synchronize()
    img = accum.cpu().numpy()[..., :3] / spp
    assert np.isfinite(img).all() and (img >= 0).all()
    ours = img.reshape(H // box, box, W // box, box, 3).mean(axis=(1, 3)).astype(np.float32)
    r = relmse(ours, ref)
    to4 = lambda a: torch_cuda.from_numpy(np.concatenate([a, np.ones(a.shape[:2] + (1,), np.float32)], axis=-1)).cuda().contiguous()
    flip, _ = mod.flip_images(to4(ref), to4(ours))
    print("config 4 (4K, depth 4, 16 spp) vs the reference path tracer, box-filtered: relMSE %.3g, LDR-FLIP %.4f, mean %.5f vs %.5f" %
          (r, flip, ours.mean(), ref.mean()))
    assert r <= RELMSE_4K_BOX and flip <= FLIP_4K_BOX
    assert abs(ours.mean() - ref.mean()) <= MEAN_4K * ref.mean()
    mod.close()
