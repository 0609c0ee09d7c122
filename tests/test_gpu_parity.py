"""GPU parity tests: every C-ABI entry point of libnbvh_b200.so against the CPU oracle on the same seeded inputs.
Bit-exact (floats compared as u32) for ray generation, N-BVH interval lists and classical closest hits; max-abs
tolerance 2e-3 on the 8 post-sigmoid outputs for the fp16-in / fp32-accumulate network (stated in each test)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NET_TOL = 2e-3  # max-abs error on post-sigmoid outputs vs the oracle's fp16-in/fp32-accumulate restatement


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def spread_params(n_params, n_mlp, seed):
    """Parameters with enough dynamic range that visibility is not pinned at 0.5 (tcnn's init gives sigmoid(~0))."""
    rng = np.random.default_rng(seed)
    p = np.zeros(n_params, np.float16)
    p[:n_mlp] = (rng.standard_normal(n_mlp) * 0.2).astype(np.float16)
    p[n_mlp:] = (rng.standard_normal(n_params - n_mlp) * 0.7).astype(np.float16)
    return p.view(np.uint16)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.fixture(scope="module")
def syn(torch_cuda, oracle, synthetic):
    import nbvh_b200
    cfg, nbvh, tlas, tlas_idx, blas = synthetic
    mod = nbvh_b200.NeuralPathModule(cfg)
    mod.set_neural_bvh(nbvh)
    mod.set_scene(tlas, tlas_idx, blas)
    n_mlp = 64 * 96 + 3 * 64 * 64 + 16 * 64
    params = spread_params(mod.n_params(), n_mlp, 11)
    mod.load_params(params)
    const = oracle.config_constants(cfg, nbvh)
    net = oracle.make_net(cfg, params)
    yield dict(mod=mod, cfg=cfg, nbvh=nbvh, scene=(tlas, tlas_idx, blas), params=params, const=const, net=net)
    mod.close()


@pytest.fixture(scope="module")
def chess(torch_cuda, oracle, chess_root):
    import nbvh_b200
    from nbvh_b200.scene_io import CHESS_BASE
    mod = nbvh_b200.load_chess_module()
    cfg = json.load(open(os.path.join(chess_root, CHESS_BASE + "_neural.json")))
    cfg["cuda_scene"]["scene_constants"]["envmap_eval"] = False
    nbvh = oracle.read_nbvh(os.path.join(chess_root, CHESS_BASE + ".nbvh"))
    params = oracle.read_network_json(os.path.join(chess_root, CHESS_BASE + "_network.json"))
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(chess_root)
    const = oracle.config_constants(cfg, nbvh)
    net = oracle.make_net(cfg, params)
    yield dict(mod=mod, cfg=cfg, nbvh=nbvh, scene=(tl, ti, bl), params=params, const=const, net=net)
    mod.close()


def _rays(env, oracle, first, n):
    o_g, d_g = env["mod"].raygen(first, n)
    o_c, d_c = oracle.raygen(env["const"], np.arange(first, first + n))
    return o_g, d_g, o_c, d_c


def check_raygen(env, oracle, n):
    W, H = env["const"].fb_w, env["const"].fb_h
    first = min((W * H // 2 // 32) * 32, W * H - n)
    o_g, d_g, o_c, d_c = _rays(env, oracle, first, n)
    assert (bits(o_g.cpu().numpy()) == bits(o_c)).all()
    assert (bits(d_g.cpu().numpy()) == bits(d_c)).all()


def check_intervals(env, oracle, n, max_i=48):
    W, H = env["const"].fb_w, env["const"].fb_h
    first = min((W * H // 2 // 32) * 32, W * H - n)
    o_g, d_g, o_c, d_c = _rays(env, oracle, first, n)
    cnt_g, node_g, mint_g, maxt_g = env["mod"].trace_intervals(o_g, d_g, max_i)
    cnt_c, node_c, mint_c, maxt_c, steps = oracle.trace_intervals(env["nbvh"], env["const"], o_c, d_c, max_i)
    assert (cnt_g.cpu().numpy().astype(np.uint32) == cnt_c).all()
    m = np.arange(max_i)[None, :] < np.minimum(cnt_c, max_i)[:, None]
    assert m.sum() > 0
    assert (node_g.cpu().numpy().astype(np.uint32)[m] == node_c[m]).all()
    assert (bits(mint_g.cpu().numpy())[m] == bits(mint_c)[m]).all()
    assert (bits(maxt_g.cpu().numpy())[m] == bits(maxt_c)[m]).all()


def check_blas(env, oracle, n):
    W, H = env["const"].fb_w, env["const"].fb_h
    first = min((W * H // 2 // 32) * 32, W * H - n)
    o_g, d_g, o_c, d_c = _rays(env, oracle, first, n)
    tl, ti, bl = env["scene"]
    sc = oracle.Scene(tl, ti, bl)
    tested = 0
    for bi, b in enumerate(bl):
        if len(b["nodes"]) < 2:
            continue
        t_c, u_c, v_c, p_c, _ = oracle.blas_intersect(sc, bi, o_c, d_c)
        t_g, u_g, v_g, p_g = env["mod"].blas_intersect(bi, o_g, d_g)
        hit = t_c < 3e38
        assert (bits(t_g.cpu().numpy()) == bits(t_c)).all()
        assert (p_g.cpu().numpy().astype(np.uint32)[hit] == p_c[hit]).all()
        assert (bits(u_g.cpu().numpy())[hit] == bits(u_c)[hit]).all()
        assert (bits(v_g.cpu().numpy())[hit] == bits(v_c)[hit]).all()
        tested += 1
    assert tested > 0


def check_query(env, oracle, torch, sizes=(0, 1, 127, 128, 129, 300, 5000)):
    rng = np.random.default_rng(7)
    x = rng.uniform(0.0, 1.0, size=(max(sizes), 9)).astype(np.float32)
    x[0, :3] = 0.0  # grid corner
    x[1, :3] = 1.0
    y0 = oracle.query(env["net"], x, mode=0)
    for n in sizes:
        xt = torch.from_numpy(x[:n].copy()).cuda()
        y = env["mod"].inference(xt).cpu().numpy()
        assert y.shape == (n, 8)
        if n:
            assert np.abs(y - y0[:n]).max() < NET_TOL


def check_render(env, oracle, torch, stride):
    import nbvh_b200
    W, H = env["const"].fb_w, env["const"].fb_h
    tl, ti, bl = env["scene"]
    sc = oracle.Scene(tl, ti, bl)
    mod = env["mod"]
    hits_dev = mod.attach_hit_buffer(W * H)
    mod.reset_counters()
    mod.neural_render(0, 0)
    torch.cuda.synchronize()
    cnt = mod.counters()
    assert cnt["rays"] == W * H
    hits_g = hits_dev.cpu().numpy().view(nbvh_b200.HIT_DTYPE)
    hits_c, _ = oracle.render_primary(sc, env["nbvh"], env["const"], env["net"], 0, W * H, stride=stride)
    hg = hits_g[::stride][: len(hits_c)]
    same = hg["valid_hit"] == hits_c["valid_hit"]
    # a visibility within the network tolerance of the 0.5 threshold may flip the discrete result (hard parts)
    assert (~same).sum() <= max(3, len(hits_c) // 500), "hit-kind flips: %d of %d" % ((~same).sum(), len(hits_c))
    nrm = (hits_c["valid_hit"] == 2) & same
    assert nrm.sum() > 0
    # classical hits (chessboard / ground): t, prim id and barycentrics bit-exact
    agree_t = bits(hg["t"][nrm]) == bits(hits_c["t"][nrm])
    assert agree_t.mean() > 0.998  # a flipped neural hit in front can shorten ray.maxt for the classical BLAS
    sel = np.flatnonzero(nrm)[agree_t]
    assert (bits(hg["albedo_or_ids"][sel][:, 1]) == bits(hits_c["albedo_or_ids"][sel][:, 1])).all()
    assert (bits(hg["normal_or_uv"][sel][:, :2]) == bits(hits_c["normal_or_uv"][sel][:, :2])).all()
    neu = (hits_c["valid_hit"] == 1) & same
    if neu.any():
        dt = np.abs(hg["t"][neu] - hits_c["t"][neu])
        assert np.median(dt) < 1e-4
        assert (dt < 1e-2).mean() > 0.995
    return cnt


# ------------------------------------------------------------------------------------------------ synthetic scene
def test_synthetic_raygen_bit_exact(syn, oracle):
    check_raygen(syn, oracle, 16384)


def test_synthetic_interval_lists_bit_exact(syn, oracle):
    check_intervals(syn, oracle, 16384)


def test_synthetic_blas_closest_hit_bit_exact(syn, oracle):
    check_blas(syn, oracle, 16384)


def test_synthetic_query_within_tolerance(syn, oracle, torch_cuda):
    check_query(syn, oracle, torch_cuda)


def test_synthetic_render_matches_oracle(syn, oracle, torch_cuda):
    check_render(syn, oracle, torch_cuda, stride=1)


def test_render_shards_reproduce_full_frame(syn, torch_cuda):
    """Multi-GPU partitioning is by pixel range: the union of shards equals the single launch bit-for-bit."""
    import nbvh_b200
    mod = syn["mod"]
    W, H = syn["const"].fb_w, syn["const"].fb_h
    n = W * H
    hits_dev = mod.attach_hit_buffer(n)
    mod.neural_render(0, 0)
    torch_cuda.cuda.synchronize()
    full = hits_dev.cpu().numpy().copy()
    hits_dev.zero_()
    from nbvh_b200.parallel import shard_pixel_range
    for r in range(3):
        first, count = shard_pixel_range(r, 3, n)
        mod.neural_render(0, 0, first_pixel=first, pixel_count=count)
    torch_cuda.cuda.synchronize()
    assert (hits_dev.cpu().numpy() == full).all()


def test_error_paths(syn):
    import nbvh_b200
    mod = syn["mod"]
    with pytest.raises(nbvh_b200.NbvhError):
        mod.load_params(np.zeros(10, np.uint16))          # wrong parameter count
    with pytest.raises(nbvh_b200.NbvhError):
        mod.neural_render(0, 0, first_pixel=5, pixel_count=7)  # unaligned shard
    with pytest.raises(nbvh_b200.NbvhError):
        mod.load_neural_bvh("/nonexistent/file.nbvh")
    with pytest.raises(nbvh_b200.NbvhError):
        mod.load_config_from_json('{"neural_bvh_renderer": {"neural_module": {"bvh_module_constants": {"segment_distribution": "Nope"}}}}')
    mod.load_config_from_json(json.dumps(syn["cfg"]))
    mod.load_params(syn["params"])


def test_file_format_round_trips(syn, oracle, tmp_path):
    mod = syn["mod"]
    p = str(tmp_path / "cut.nbvh")
    mod.save_neural_bvh(p)
    assert oracle.read_nbvh(p).tobytes() == np.ascontiguousarray(syn["nbvh"]).tobytes()
    q = str(tmp_path / "net.json")
    mod.save_network_model(q)
    assert (oracle.read_network_json(q) == syn["params"]).all()
    mod.load_network_model(q)
    assert (mod.get_params() == syn["params"]).all()


def test_network_file_float_payload_and_errors(syn, tmp_path):
    """tcnn's deserialize also accepts params_type "float" (trainer.h:292-295): fp32 payloads are converted to the fp16 parameters;
    malformed files fail loudly and leave the loaded parameters untouched."""
    import nbvh_b200
    mod = syn["mod"]
    bits = syn["params"]
    f32 = bits.view(np.float16).astype(np.float32)
    doc = {"n_params": int(bits.size), "params_binary": {"bytes": f32.view(np.uint8).tolist(), "subtype": None}, "params_type": "float"}
    q = str(tmp_path / "net_float.json")
    json.dump(doc, open(q, "w"))
    mod.load_params(np.zeros_like(bits))
    mod.load_network_model(q)
    assert (mod.get_params() == bits).all()
    for bad in ({**doc, "params_type": "double"}, {**doc, "n_params": int(bits.size) + 1}, {"n_params": 4}):
        json.dump(bad, open(q, "w"))
        with pytest.raises(nbvh_b200.NbvhError):
            mod.load_network_model(q)
        assert (mod.get_params() == bits).all()


# ------------------------------------------------------------------------------------------------ shipped Chess
def test_chess_loaders_match_independent_readers(chess):
    assert chess["mod"].neural_bvh_nodes().tobytes() == chess["nbvh"].tobytes()
    assert (chess["mod"].get_params() == chess["params"]).all()
    assert chess["mod"].n_params() == 5411840
    assert len(chess["nbvh"]) == 25249


def test_chess_raygen_bit_exact(chess, oracle):
    check_raygen(chess, oracle, 65536)


def test_chess_interval_lists_bit_exact(chess, oracle):
    check_intervals(chess, oracle, 65536)


def test_chess_blas_closest_hit_bit_exact(chess, oracle):
    check_blas(chess, oracle, 65536)


def test_chess_query_within_tolerance(chess, oracle, torch_cuda):
    check_query(chess, oracle, torch_cuda)


def test_chess_render_1080p_matches_oracle(chess, oracle, torch_cuda):
    cnt = check_render(chess, oracle, torch_cuda, stride=51)
    assert cnt["queries"] > 100000


def test_hoisted_division_is_ieee(syn):
    """The slab tests divide by a per-ray hoisted reciprocal (csrc/device_core.cuh: hoisted_div).  Inside its documented
    operand range every quotient must carry the same bits as div.rn: 2^28 random + boundary-mantissa pairs on the device."""
    assert syn["mod"].selftest_division(1 << 28, seed=12345) == 0
    assert syn["mod"].selftest_division(1 << 24, seed=777) == 0


def test_pipelined_host_render_equals_synchronous(syn, torch_cuda):
    """nbvh_render_host_submit / _wait (two frames in flight) returns the same frames as the blocking nbvh_render_host."""
    mod = syn["mod"]
    W, H = syn["const"].fb_w, syn["const"].fb_h
    mod.reinitialise_inference_data((W, H))
    n = W * H
    ref = []
    for s in range(4):
        buf = np.zeros(n * 4, np.uint8)
        mod.neural_render_host(0, s, ldr_host=buf)
        ref.append(buf.copy())
    bufs = [torch_cuda.empty(n * 4, dtype=torch_cuda.uint8, pin_memory=True).numpy() for _ in range(2)]
    got = []
    for s in range(4):
        if s >= 2:
            mod.neural_render_host_wait(s & 1)
            got.append(bufs[s & 1].copy())
        mod.neural_render_host_submit(0, s, bufs[s & 1], s & 1)
    for s in (0, 1):
        mod.neural_render_host_wait(s)
        got.append(bufs[s].copy())
    for a, b in zip(ref, got):
        assert (a == b).all()
    # up to NBVH_MAX_FRAMES_IN_FLIGHT = 4 slots, each with its own queue and stream; a fifth slot is refused
    import nbvh_b200
    for depth in (3, 4):
        bufs = [torch_cuda.empty(n * 4, dtype=torch_cuda.uint8, pin_memory=True).numpy() for _ in range(depth)]
        got = {}
        for s in range(9):
            k = s % depth
            if s >= depth:
                mod.neural_render_host_wait(k)
                got[s - depth] = bufs[k].copy()
            mod.neural_render_host_submit(0, s % 4, bufs[k], k)
        for s in range(9 - depth, 9):
            mod.neural_render_host_wait(s % depth)
            got[s] = bufs[s % depth].copy()
        assert all((got[s] == ref[s % 4]).all() for s in range(9))
    with pytest.raises(nbvh_b200.NbvhError):
        mod.neural_render_host_submit(0, 0, bufs[0], 4)
