"""Edge cases of the hot-path entry points: empty and ragged batches (sizes that are not multiples of the 128-ray tile, the 32-ray
warp chunk or the queue chunk), rays that miss everything, truncated interval lists, single-chunk frames.

Bit-exact checks: a ragged call must reproduce the matching rows of one large call (the result of a ray / query may not depend on
its tile mates), interval lists must equal the oracle's, truncated lists must be prefixes of the full ones."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(oracle, synthetic):
    import torch
    import nbvh_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg, nbvh, tlas, tlas_idx, blas = synthetic
    mod = nbvh_b200.NeuralPathModule(cfg)
    mod.set_neural_bvh(nbvh)
    mod.set_scene(tlas, tlas_idx, blas)
    rng = np.random.default_rng(21)
    n_mlp = 64 * 96 + 3 * 64 * 64 + 16 * 64
    p = np.zeros(mod.n_params(), np.float16)
    p[:n_mlp] = (rng.standard_normal(n_mlp) * 0.2).astype(np.float16)
    p[n_mlp:] = (rng.standard_normal(len(p) - n_mlp) * 0.7).astype(np.float16)
    mod.load_params(p.view(np.uint16))
    const = oracle.config_constants(cfg, nbvh)
    yield dict(torch=torch, mod=mod, cfg=cfg, nbvh=nbvh, const=const, net=oracle.make_net(cfg, p.view(np.uint16)), O=oracle, nb=nbvh_b200)
    mod.close()


def test_query_ragged_and_empty_batches(env):
    torch, mod, O = env["torch"], env["mod"], env["O"]
    x = np.random.default_rng(1).uniform(0.01, 0.99, (1000, 9)).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    full = mod.inference(xt).cpu().numpy()
    assert np.abs(full - O.query(env["net"], x, mode=0)).max() < 2e-3
    for n in (1, 2, 31, 32, 33, 127, 128, 129, 255, 257, 999):
        part = mod.inference(xt[:n].contiguous()).cpu().numpy()
        assert (part.view(np.uint32) == full[:n].view(np.uint32)).all(), "query result depends on the batch size (n = %d)" % n
    # a window that starts in the middle of a tile of the large call
    part = mod.inference(xt[77:77 + 300].contiguous()).cpu().numpy()
    assert (part.view(np.uint32) == full[77:377].view(np.uint32)).all()
    empty = mod.inference(torch.empty((0, 9), dtype=torch.float32, device="cuda"))
    assert empty.shape == (0, 8)
    # the output buffer of a ragged call is not written past row n
    out = torch.full((200, 8), -7.0, dtype=torch.float32, device="cuda")
    mod.inference(xt[:129].contiguous(), out)
    assert float(out[129:].min()) == -7.0 and float(out[129:].max()) == -7.0


def _rays(env, n, seed=3):
    o, d = env["nb"].synthetic_ray_batch(env["nbvh"][0], n, seed=seed, device="cuda")
    return o.contiguous(), d.contiguous()


def test_trace_rays_ragged_empty_and_missing(env):
    torch, mod, nb = env["torch"], env["mod"], env["nb"]
    o, d = _rays(env, 3000)
    full = mod.trace_rays(o, d).cpu().numpy().view(nb.HIT_DTYPE)
    assert (full["valid_hit"] == 1).sum() > 50 and (full["valid_hit"] == 0).sum() > 50
    for n in (1, 31, 33, 127, 129, 1025):
        part = mod.trace_rays(o[:n].contiguous(), d[:n].contiguous()).cpu().numpy().view(nb.HIT_DTYPE)
        assert part.tobytes() == full[:n].tobytes(), "hit records depend on the batch size (n = %d)" % n
    assert mod.trace_rays(o[:0].contiguous(), d[:0].contiguous()).numel() == 0
    # rays that start outside the scene box and point away from it: no hit, t stays at the ray's far bound
    lo, hi = env["nbvh"][0]["mn"], env["nbvh"][0]["mx"]
    oo = torch.from_numpy(np.tile((hi + (hi - lo)).astype(np.float32), (100, 1))).cuda()
    dd = torch.from_numpy(np.tile(np.float32([0.0, 1.0, 0.0]), (100, 1))).cuda()
    miss = mod.trace_rays(oo, dd).cpu().numpy().view(nb.HIT_DTYPE)
    assert (miss["valid_hit"] == 0).all() and (miss["t"] > 3e38).all()


def test_interval_lists_truncated_and_ragged(env):
    torch, mod, O = env["torch"], env["mod"], env["O"]
    o, d = _rays(env, 777, seed=9)
    oc, dc = o.cpu().numpy(), d.cpu().numpy()
    cnt, node, mint, maxt = [t.cpu().numpy() for t in mod.trace_intervals(o, d, 48)]
    cnt_c, node_c, mint_c, maxt_c, _ = O.trace_intervals(env["nbvh"], env["const"], oc, dc, 48)
    assert (cnt.astype(np.uint32) == cnt_c).all() and cnt.max() >= 3 and (cnt == 0).any()
    m = np.arange(48)[None, :] < np.minimum(cnt, 48)[:, None]
    assert (node.astype(np.uint32)[m] == node_c[m]).all() and (mint.view(np.uint32)[m] == mint_c.view(np.uint32)[m]).all()
    for cap in (1, 2):  # truncated lists are prefixes of the full ones; the count still reports every pierced leaf
        c1, n1, a1, b1 = [t.cpu().numpy() for t in mod.trace_intervals(o, d, cap)]
        assert (c1 == cnt).all()
        m1 = np.arange(cap)[None, :] < np.minimum(cnt, cap)[:, None]
        assert (n1[m1] == node[:, :cap][m1]).all() and (a1.view(np.uint32)[m1] == mint[:, :cap].view(np.uint32)[m1]).all()
        assert (b1.view(np.uint32)[m1] == maxt[:, :cap].view(np.uint32)[m1]).all()
    c0, _, _, _ = mod.trace_intervals(o[:0].contiguous(), d[:0].contiguous(), 4)
    assert c0.numel() == 0


def test_single_chunk_and_tail_shards_equal_full_frame(env):
    torch, mod, nb = env["torch"], env["mod"], env["nb"]
    w, h = env["cfg"]["cuda_scene"]["fb_size"]
    n = w * h
    accum, _ = mod.attach_framebuffer((w, h))
    hits = mod.attach_hit_buffer(n)
    mod.neural_render(0, 0)
    torch.cuda.synchronize()
    full_a, full_h = accum.cpu().numpy().reshape(n, 4).copy(), hits.cpu().numpy().view(nb.HIT_DTYPE).copy()
    accum.zero_()
    hits.zero_()
    # one 32-pixel chunk, a ragged middle range, and the last chunk of the frame
    for first, count in ((0, 32), (1024, 32 * 37), (n - 32, 32)):
        mod.neural_render(0, 0, first_pixel=first, pixel_count=count)
    torch.cuda.synchronize()
    got_a, got_h = accum.cpu().numpy().reshape(n, 4), hits.cpu().numpy().view(nb.HIT_DTYPE)
    touched = np.zeros(n, bool)
    for first, count in ((0, 32), (1024, 32 * 37), (n - 32, 32)):
        touched[first:first + count] = True
    # pixels are swizzled inside 8x4 blocks, so compare as sets: every rendered pixel equals the full frame, all others stay zero
    rendered = got_h["t"] != 0
    assert rendered.sum() == touched.sum()
    assert (got_a[rendered].view(np.uint32) == full_a[rendered].view(np.uint32)).all()
    assert got_h[rendered].tobytes() == full_h[rendered].tobytes()
    assert not got_a[~rendered].any()


def test_malformed_files_fail_loudly_and_leave_state_untouched(env, tmp_path):
    import struct
    mod, nb = env["mod"], env["nb"]
    good = str(tmp_path / "good.nbvh")
    mod.save_neural_bvh(good)
    raw = open(good, "rb").read()
    nodes_before, params_before = mod.neural_bvh_nodes().tobytes(), mod.get_params().tobytes()
    n_nodes, n_prims = struct.unpack_from("<QQ", raw)
    cases = {"truncated": raw[:len(raw) - 13], "short": raw[:7], "empty": b"", "zero_nodes": struct.pack("<QQ", 0, 0),
             "huge_counts": struct.pack("<QQ", n_nodes, (1 << 61) + n_prims) + raw[16:],      # n_prims * 8 wraps around 2^64
             "trailing": raw + b"\0" * 8}
    for name, blob in cases.items():
        p = tmp_path / (name + ".nbvh")
        p.write_bytes(blob)
        with pytest.raises(nb.NbvhError):
            mod.load_neural_bvh(str(p))
        assert mod.neural_bvh_nodes().tobytes() == nodes_before, name
    for name, text in {"garbage": "{not json", "no_bytes": '{"n_params": 3, "params_type": "__half"}',
                       "short_bytes": '{"n_params": 4, "params_binary": {"bytes": [1, 2]}, "params_type": "__half"}'}.items():
        p = tmp_path / (name + ".json")
        p.write_text(text)
        with pytest.raises(nb.NbvhError):
            mod.load_network_model(str(p))
        assert mod.get_params().tobytes() == params_before, name
    for text in ("{", '{"neural_bvh_renderer": {"neural_module": {"pt_output_constants": {"depth_loss": "Nope"}}}}',
                 '{"cuda_scene": {"camera": {"matrix": [1, 2, 3]}}}'):
        with pytest.raises(nb.NbvhError):
            nb.NeuralPathModule(text)
    # a config that is refused half-way (after the scene / renderer blocks parsed fine) leaves every constant block, the network
    # description and the loaded parameters as they were; the same for a constant block the kernels cannot run
    import copy
    import json
    from nbvh_b200 import _capi as A
    blocks = (A.CONST_RENDERER, A.CONST_MODULE, A.CONST_PT_OUTPUT, A.CONST_PATH, A.CONST_SCENE, A.CONST_NETWORK, A.CONST_OPTIMIZER)
    before = [bytes(mod.get_constants(w)) for w in blocks]
    bad = copy.deepcopy(env["cfg"])
    bad.setdefault("cuda_scene", {}).setdefault("scene_constants", {})["envmap_strength"] = 7.5
    bad["neural_bvh_renderer"].setdefault("bvh_constants", {})["rendered_lod"] = 3
    for breakage in ({"pt_output_constants": {"depth_loss": "Nope"}}, {"network": {"n_neurons": 32}},
                     {"bvh_module_constants": {"input_segment_points": 5}}):
        cfg_bad = copy.deepcopy(bad)
        for k, v in breakage.items():
            cfg_bad["neural_bvh_renderer"]["neural_module"].setdefault(k, {}).update(v)
        with pytest.raises(nb.NbvhError):
            mod.load_config_from_json(json.dumps(cfg_bad))
        assert [bytes(mod.get_constants(w)) for w in blocks] == before, breakage
        assert mod.get_params().tobytes() == params_before, breakage
    net = mod.get_constants(A.CONST_NETWORK)
    net.n_neurons = 128
    with pytest.raises(nb.NbvhError):
        mod.set_constants(A.CONST_NETWORK, net)
    pt = mod.get_constants(A.CONST_PT_OUTPUT)
    pt.output_network_dims = 12
    with pytest.raises(nb.NbvhError):
        mod.set_constants(A.CONST_PT_OUTPUT, pt)
    assert [bytes(mod.get_constants(w)) for w in blocks] == before and mod.get_params().tobytes() == params_before
    mod.load_neural_bvh(good)
    assert mod.neural_bvh_nodes().tobytes() == nodes_before
