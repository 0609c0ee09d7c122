#!/usr/bin/env python3
"""GPU self-check with verbose diagnostics (development aid; the asserting versions live in tests/).
Compares every C-ABI entry point against the CPU oracle on the synthetic scene and, when present, the Chess fixture."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

import nbvh_b200  # noqa: E402
from oracle import oracle as O  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def report(name, ok, extra=""):
    print("[%s] %s %s" % ("PASS" if ok else "FAIL", name, extra), flush=True)
    return ok


def check_module(tag, mod, cfg, nbvh_nodes, scene_arrays, params_bits, n_rays=20000, full_render=True):
    dev = torch.device("cuda", 0)
    ok_all = True
    tlas_nodes, tlas_idx, blas = scene_arrays
    const = O.config_constants(cfg, nbvh_nodes)
    net = O.make_net(cfg, params_bits)
    W, H = const.fb_w, const.fb_h

    # ---- raygen
    n = min(n_rays, W * H)
    first = (W * H // 2 // 32) * 32
    first = min(first, W * H - n)
    o_g, d_g = mod.raygen(first, n)
    o_c, d_c = O.raygen(const, np.arange(first, first + n))
    ok = (bits(o_g.cpu().numpy()) == bits(o_c)).all() and (bits(d_g.cpu().numpy()) == bits(d_c)).all()
    ok_all &= report(tag + " raygen bit-exact", ok, "max|dd|=%g" % np.abs(d_g.cpu().numpy() - d_c).max())

    # ---- unconditional interval lists
    MAXI = 48
    cnt_g, node_g, mint_g, maxt_g = mod.trace_intervals(o_g, d_g, MAXI)
    cnt_c, node_c, mint_c, maxt_c, steps_c = O.trace_intervals(nbvh_nodes, const, o_c, d_c, MAXI)
    cnt_g = cnt_g.cpu().numpy().astype(np.uint32)
    same_cnt = (cnt_g == cnt_c).all()
    m = np.arange(MAXI)[None, :] < np.minimum(cnt_c, MAXI)[:, None]
    same = same_cnt and (node_g.cpu().numpy().astype(np.uint32)[m] == node_c[m]).all() and \
        (bits(mint_g.cpu().numpy())[m] == bits(mint_c)[m]).all() and (bits(maxt_g.cpu().numpy())[m] == bits(maxt_c)[m]).all()
    ok_all &= report(tag + " interval lists bit-exact", same,
                     "rays=%d intervals=%d max/ray=%d mismatching counts=%d" % (n, cnt_c.sum(), cnt_c.max(), (cnt_g != cnt_c).sum()))

    # ---- classical BLAS closest hit
    for bi, b in enumerate(blas):
        if len(b["nodes"]) < 2:
            continue
        sc = O.Scene(tlas_nodes, tlas_idx, blas)
        t_c, u_c, v_c, p_c, _ = O.blas_intersect(sc, bi, o_c, d_c)
        t_g, u_g, v_g, p_g = mod.blas_intersect(bi, o_g, d_g)
        hit = t_c < 3e38
        ok = (bits(t_g.cpu().numpy()) == bits(t_c)).all() and (p_g.cpu().numpy().astype(np.uint32)[hit] == p_c[hit]).all() and \
            (bits(u_g.cpu().numpy())[hit] == bits(u_c)[hit]).all() and (bits(v_g.cpu().numpy())[hit] == bits(v_c)[hit]).all()
        ok_all &= report(tag + " BLAS %d closest hit bit-exact" % bi, ok, "hits=%d/%d" % (hit.sum(), n))

    # ---- network queries
    rng = np.random.default_rng(7)
    nq = 5000
    x = rng.uniform(0.02, 0.98, size=(nq, 9)).astype(np.float32)
    y_g = mod.inference(torch.from_numpy(x).to(dev)).cpu().numpy()
    y0 = O.query(net, x, mode=0)
    y1 = O.query(net, x, mode=1)
    y2 = O.query(net, x, mode=2)
    e0 = np.abs(y_g - y0).max()
    ok_all &= report(tag + " query vs oracle(fp16 in/fp32 acc) tol 2e-3", e0 < 2e-3,
                     "max|err|=%.3g  vs tcnn-emulation=%.3g  vs fp32=%.3g  (emu vs fp32=%.3g)" %
                     (e0, np.abs(y_g - y1).max(), np.abs(y_g - y2).max(), np.abs(y1 - y2).max()))
    if e0 >= 2e-3:
        bad = np.argwhere(np.abs(y_g - y0) > 2e-3)
        print("   first bad rows:", bad[:8].tolist())
        print("   gpu :", y_g[:3])
        print("   ref :", y0[:3])
        print("   rows with error:", np.unique(bad[:, 0])[:32].tolist(), " cols:", np.unique(bad[:, 1]).tolist())
    # a partial tile + tiny batch
    for nn in (1, 127, 129, 300):
        yy = mod.inference(torch.from_numpy(x[:nn].copy()).to(dev)).cpu().numpy()
        ok_all &= report(tag + " query n=%d" % nn, np.abs(yy - y0[:nn]).max() < 2e-3, "max|err|=%.3g" % np.abs(yy - y0[:nn]).max())

    # ---- whole frame
    if full_render:
        sc = O.Scene(tlas_nodes, tlas_idx, blas)
        hits_dev = mod.attach_hit_buffer(W * H)
        mod.reset_counters()
        torch.cuda.synchronize()
        t0 = time.time()
        mod.neural_render(0, 0)
        torch.cuda.synchronize()
        t1 = time.time()
        for _ in range(3):
            mod.neural_render(0, 0)
        torch.cuda.synchronize()
        t2 = time.time()
        print("   render: first %.3f ms, steady %.3f ms/frame -> %.1f Mrays/s" % ((t1 - t0) * 1e3, (t2 - t1) / 3 * 1e3,
                                                                                 W * H / ((t2 - t1) / 3) / 1e6))
        cnt = mod.counters()
        print("   counters (4 frames):", cnt)
        hits_g = hits_dev.cpu().numpy().view(nbvh_b200.HIT_DTYPE)
        stride = max(1, (W * H) // 40000)
        t0 = time.time()
        hits_c, cc = O.render_primary(sc, nbvh_nodes, const, net, 0, W * H, stride=stride)
        print("   oracle: %d rays in %.2f s, counters %s" % (len(hits_c), time.time() - t0, cc))
        hg = hits_g[::stride][: len(hits_c)]
        same_kind = hg["valid_hit"] == hits_c["valid_hit"]
        print("   hit kinds gpu:", np.bincount(hg["valid_hit"], minlength=3), " oracle:", np.bincount(hits_c["valid_hit"], minlength=3))
        nrm = hits_c["valid_hit"] == 2
        ok = same_kind[nrm].all() and (bits(hg["t"][nrm]) == bits(hits_c["t"][nrm])).all() and \
            (bits(hg["albedo_or_ids"][nrm][:, 1]) == bits(hits_c["albedo_or_ids"][nrm][:, 1])).all()
        ok_all &= report(tag + " render: classical hits bit-exact (t, prim)", ok, "n=%d" % nrm.sum())
        miss = hits_c["valid_hit"] == 0
        ok_all &= report(tag + " render: misses agree", same_kind[miss].all(), "n=%d disagree=%d" % (miss.sum(), (~same_kind[miss]).sum()))
        neu = (hits_c["valid_hit"] == 1) & same_kind
        dt = np.abs(hg["t"][neu] - hits_c["t"][neu])
        flips = (~same_kind).sum()
        ok_all &= report(tag + " render: neural hits within 1e-3, kind flips < 0.5%",
                         (dt.max() if neu.any() else 0) < 1e-3 and flips <= 0.005 * len(hits_c),
                         "n=%d max|dt|=%.3g mean|dt|=%.3g kind flips=%d" % (neu.sum(), dt.max() if neu.any() else 0,
                                                                           dt.mean() if neu.any() else 0, flips))
    return ok_all


def main():
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    ok = True
    # ---------------- synthetic scene
    cfg, nbvh, tlas, tlas_idx, blas = nbvh_b200.synthetic_scene()
    mod = nbvh_b200.NeuralPathModule(cfg)
    mod.set_neural_bvh(nbvh)
    mod.set_scene(tlas, tlas_idx, blas)
    mod.init_params(1337)
    params = mod.get_params()
    print("synthetic: n_params", mod.n_params(), "nbvh nodes", len(nbvh))
    ok &= check_module("synthetic", mod, cfg, nbvh, (tlas, tlas_idx, blas), params, n_rays=16384)
    mod.close()
    # ---------------- Chess
    root = nbvh_b200.chess_fixture_dir()
    if root:
        t0 = time.time()
        mod = nbvh_b200.load_chess_module()
        print("chess loaded through the C ABI in %.2f s" % (time.time() - t0))
        cfg = json.load(open(os.path.join(root, nbvh_b200.scene_io.CHESS_BASE + "_neural.json")))
        cfg["cuda_scene"]["scene_constants"]["envmap_eval"] = False
        nbvh_c = O.read_nbvh(os.path.join(root, nbvh_b200.scene_io.CHESS_BASE + ".nbvh"))
        ok &= report("chess .nbvh loader == numpy reader", (mod.neural_bvh_nodes().tobytes() == nbvh_c.tobytes()))
        params_c = O.read_network_json(os.path.join(root, nbvh_b200.scene_io.CHESS_BASE + "_network.json"))
        ok &= report("chess weights loader == python json", (mod.get_params() == params_c).all())
        tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root)
        ok &= check_module("chess", mod, cfg, nbvh_c, (tl, ti, bl), params_c, n_rays=65536)
        mod.close()
    else:
        print("chess fixture not present")
    print("SELFCHECK", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
