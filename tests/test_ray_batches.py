"""Ray-batch neural queries (nbvh_trace_rays) and the synthetic instanced scene of BASELINE config 5.

CPU: the level-synchronous builder yields a proper fixed-depth tree-cut (children adjacent and enclosed, balanced leaves).
GPU: nbvh_trace_rays against the oracle's per-ray restatement on the same random rays: hit kinds equal up to the visibility
threshold tolerance, t within the network tolerance for neural hits, bit-exact for classical hits; identical to the primary-hit
records nbvh_render produces for camera rays."""
import numpy as np
import pytest

FLIP_RATE = 5e-3
T_TOL = 2e-3  # relative to the leaf interval length: the network's depth output differs by <= 2e-3 (fp32 vs oracle ordering)


def small_instanced():
    import nbvh_b200
    return nbvh_b200.synthetic_instanced_scene(lattice=(4, 4, 4), subdiv=2, cut_depth=8, log2_hashmap=14, device="cpu")


def test_instanced_scene_builder_is_a_fixed_depth_cut():
    cfg, nbvh, tlas, ti, blas, info = small_instanced()
    depth = 8
    n_int = (1 << depth) - 1
    assert len(nbvh) == (1 << (depth + 1)) - 1 and info["neural_leaves"] == 1 << depth
    assert (nbvh["a"][:n_int] == 0).all() and (nbvh["b"][:n_int] == 2 * np.arange(n_int) + 1).all()
    assert (nbvh["a"][n_int:] == 8).all() and (nbvh["b"][n_int:] == 0xFFFFFFFF).all()
    for i in range(n_int):
        for c in (2 * i + 1, 2 * i + 2):
            assert (nbvh[c]["mn"] >= nbvh[i]["mn"]).all() and (nbvh[c]["mx"] <= nbvh[i]["mx"]).all()
    assert info["triangles"] == 64 * 320 and info["triangles_per_leaf_max"] - info["triangles_per_leaf_min"] <= 1
    assert (nbvh["mx"] > nbvh["mn"]).all()
    assert tlas[0]["b"] == 2 and len(ti) == 1


def test_oracle_trace_rays_matches_render_primary(oracle, synthetic):
    cfg, nbvh, tlas, tlas_idx, blas = synthetic
    const = oracle.config_constants(cfg, nbvh)
    rng = np.random.default_rng(3)
    total, _ = oracle.expected_n_params(oracle.make_net(cfg, np.zeros(1, np.uint16)))
    params = (rng.standard_normal(total) * 0.3).astype(np.float16).view(np.uint16)
    net = oracle.make_net(cfg, params)
    scene = oracle.Scene(tlas, tlas_idx, blas)
    n = 4096
    o, d = oracle.raygen(const, np.arange(n))
    a, _ = oracle.trace_rays(scene, nbvh, const, net, o, d)
    b, _ = oracle.render_primary(scene, nbvh, const, net, 0, n)
    assert a.tobytes() == b.tobytes()


@pytest.mark.gpu
def test_gpu_trace_rays_matches_oracle(oracle):
    import torch
    import nbvh_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg, nbvh, tlas, ti, blas, info = small_instanced()
    mod = nbvh_b200.NeuralPathModule(cfg)
    mod.set_neural_bvh(nbvh)
    mod.set_scene(tlas, ti, blas)
    rng = np.random.default_rng(5)
    n_mlp = 64 * 96 + 3 * 64 * 64 + 16 * 64
    p = np.zeros(mod.n_params(), np.float16)
    p[:n_mlp] = (rng.standard_normal(n_mlp) * 0.2).astype(np.float16)
    p[n_mlp:] = (rng.standard_normal(len(p) - n_mlp) * 0.7).astype(np.float16)
    params = p.view(np.uint16)
    mod.load_params(params)
    n = 20000
    o, d = nbvh_b200.synthetic_ray_batch(nbvh[0], n, seed=42, device="cuda")
    const = oracle.config_constants(cfg, nbvh)
    net = oracle.make_net(cfg, params)
    scene = oracle.Scene(tlas, ti, blas)
    for depth in (1, 2):
        hits = mod.trace_rays(o, d, ray_depth=depth).cpu().numpy().view(nbvh_b200.HIT_DTYPE)
        ref, cnt = oracle.trace_rays(scene, nbvh, const, net, o.cpu().numpy(), d.cpu().numpy(), ray_depth=depth)
        same = hits["valid_hit"] == ref["valid_hit"]
        assert (~same).sum() <= max(3, FLIP_RATE * n), "depth %d: %d hit-kind flips" % (depth, (~same).sum())
        assert (ref["valid_hit"] == 1).sum() > n // 20 and (ref["valid_hit"] == 0).sum() > n // 20
        hit = same & (ref["valid_hit"] == 1)
        dt = np.abs(hits["t"][hit] - ref["t"][hit])
        assert np.quantile(dt, 0.99) <= T_TOL, "depth %d: neural t differs by %g at the 99th percentile" % (depth, np.quantile(dt, 0.99))
        miss = same & (ref["valid_hit"] == 0)
        assert (hits["t"][miss].view(np.uint32) == ref["t"][miss].view(np.uint32)).all()
    mod.close()


@pytest.mark.gpu
def test_gpu_trace_rays_equals_render_hit_records(synthetic):
    import torch
    import nbvh_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg, nbvh, tlas, tlas_idx, blas = synthetic
    mod = nbvh_b200.NeuralPathModule(cfg)
    mod.set_neural_bvh(nbvh)
    mod.set_scene(tlas, tlas_idx, blas)
    mod.init_params(7)
    W, H = cfg["cuda_scene"]["fb_size"]
    n = W * H
    buf = mod.attach_hit_buffer(n)
    mod.neural_render(0, 0)
    torch.cuda.synchronize()
    a = buf.cpu().numpy().view(nbvh_b200.HIT_DTYPE).copy()
    o, d = mod.raygen(0, n)
    b = mod.trace_rays(o, d, ray_depth=1).cpu().numpy().view(nbvh_b200.HIT_DTYPE)
    assert a.tobytes() == b.tobytes()
    mod.close()


@pytest.mark.gpu
def test_gpu_full_size_deep_cut_matches_oracle(oracle):
    """BASELINE config 5 at the size bench.py --sweep times: the 10.38 M-triangle instanced scene with its depth-17 tree-cut (262,143
    nodes, 131,072 neural leaves).  A 2^16-ray sample of the benched batch: the ordered leaf-interval lists (counts, node ids, entry /
    exit parameters) are BIT-EXACT against the oracle's walk of the same deep cut, and the hit records of nbvh_trace_rays agree with
    the oracle's per-ray restatement like on the small scenes (kinds up to the visibility-threshold tolerance, neural t within the
    network tolerance, misses bit-exact)."""
    import torch
    import nbvh_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg, nbvh, tlas, ti, blas, info = nbvh_b200.synthetic_instanced_scene(device="cuda")
    assert info["neural_leaves"] == 1 << 17 and len(nbvh) == (1 << 18) - 1 and info["triangles"] > 10_000_000
    mod = nbvh_b200.NeuralPathModule(cfg)
    mod.set_neural_bvh(nbvh)
    mod.set_scene(tlas, ti, blas)
    rng = np.random.default_rng(1337)
    n_mlp = 64 * 96 + 3 * 64 * 64 + 16 * 64
    p = np.zeros(mod.n_params(), np.float16)
    p[:n_mlp] = (rng.standard_normal(n_mlp) * 0.2).astype(np.float16)
    p[n_mlp:] = (rng.standard_normal(len(p) - n_mlp) * 0.7).astype(np.float16)
    params = p.view(np.uint16)
    mod.load_params(params)
    n = 1 << 16
    o, d = nbvh_b200.synthetic_ray_batch(nbvh[0], n, seed=42, device="cuda")
    oc, dc = o.cpu().numpy(), d.cpu().numpy()
    const = oracle.config_constants(cfg, nbvh)
    # ---- interval lists on the deep cut: bit-exact
    MAX_IV = 32
    cnt, node, mint, maxt = mod.trace_intervals(o, d, MAX_IV)
    cnt_c, node_c, mint_c, maxt_c, _ = oracle.trace_intervals(nbvh, const, oc, dc, MAX_IV)
    assert (cnt.cpu().numpy().astype(np.uint32) == cnt_c).all()
    m = np.arange(MAX_IV)[None, :] < np.minimum(cnt_c, MAX_IV)[:, None]
    assert m.sum() > 4 * n  # several leaves per ray on this cut
    assert (node.cpu().numpy().astype(np.uint32)[m] == node_c[m]).all()
    assert (mint.cpu().numpy().view(np.uint32)[m] == mint_c.view(np.uint32)[m]).all()
    assert (maxt.cpu().numpy().view(np.uint32)[m] == maxt_c.view(np.uint32)[m]).all()
    # ---- hit records of the ray-batch query
    net = oracle.make_net(cfg, params)
    scene = oracle.Scene(tlas, ti, blas)
    hits = mod.trace_rays(o, d, ray_depth=1).cpu().numpy().view(nbvh_b200.HIT_DTYPE)
    ref, _ = oracle.trace_rays(scene, nbvh, const, net, oc, dc, ray_depth=1)
    same = hits["valid_hit"] == ref["valid_hit"]
    assert (~same).sum() <= max(3, FLIP_RATE * n), "%d hit-kind flips" % (~same).sum()
    hit = same & (ref["valid_hit"] == 1)
    assert hit.sum() > n // 20
    dt = np.abs(hits["t"][hit] - ref["t"][hit])
    assert np.quantile(dt, 0.99) <= T_TOL
    miss = same & (ref["valid_hit"] == 0)
    assert (hits["t"][miss].view(np.uint32) == ref["t"][miss].view(np.uint32)).all()
    mod.close()
