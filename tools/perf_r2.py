#!/usr/bin/env python3
"""Development timing harness of round 2 (GPU box, CUDA events): headline frame with its per-kernel split and k_neural timeline,
path configs, the two-pass sample generator at several batch sizes, and the optimisation-step kernels.
usage: perf_r2.py [out.json] [--quick]"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import nbvh_b200  # noqa: E402

M = (1 << 64) - 1


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    quick = "--quick" in sys.argv
    out = {"lib": os.environ.get("NBVH_B200_LIB", "default")}
    root = nbvh_b200.chess_fixture_dir()
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # ---- headline frame
    mod = nbvh_b200.load_chess_module(max_depth=1, envmap_eval=False)
    mod.reinitialise_inference_data((1920, 1080))
    for i in range(5):
        mod.neural_render(i, 0)
    torch.cuda.synchronize()
    n = 30
    mod.reset_counters()
    mod.enable_stage_timing(n)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i in range(n):
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        mod.neural_render(5 + i, 0)
        ev[i][1].record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / n
    pm, nm, fr = mod.stage_times()
    mod.enable_stage_timing(0)
    c = mod.counters()
    out["frame_1080p"] = {"ms": ms, "Mrays_s": 1920 * 1080 / ms / 1e3, "k_primary_ms": pm / max(fr, 1), "k_neural_ms": nm / max(fr, 1),
                          "queries": c["queries"] / n, "queued": c["queued_rays"] / n, "tile_iterations": c["tile_iterations"] / n,
                          "fill": c["queries"] / (128.0 * max(c["tile_iterations"], 1))}
    mod.reset_counters()
    mod.neural_render(99, 0)
    torch.cuda.synchronize()
    c = mod.counters()
    t0, tx, t1 = M - c["t_start_inv"], M - c["t_exhaust_inv"], c["t_end"]
    out["k_neural_timeline_us"] = {"span": (t1 - t0) / 1e3, "queue_dry_at": (tx - t0) / 1e3, "mean_cta_lifetime": c["t_cta_sum"] / 296 / 1e3}
    if c.get("ph_mlp", 0) or c.get("ph_mlp_d", 0):
        it_s, it_d = max(c["tile_iterations"] - c["iters_drain"], 1), max(c["iters_drain"], 1)
        out["k_neural_phase_cycles_per_iteration"] = {
            "steady": {k: c["ph_" + k] / it_s for k in ("walk", "sync", "encode", "mlp")}, "steady_iterations": it_s,
            "drain": {k: c["ph_" + k + "_d"] / it_d for k in ("walk", "sync", "encode", "mlp")}, "drain_iterations": it_d}
    ms8 = timeit(lambda: mod.neural_render_samples(0, 8), iters=5, warmup=2)
    out["frame_8spl"] = {"ms_per_launch": ms8, "ms_per_sample_frame": ms8 / 8}
    print(json.dumps(out), flush=True)
    if "--frame-only" in sys.argv:
        mod.close()
        return

    # ---- optimisation: sampler + step kernels
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root, with_neural_mesh=True, with_attributes=True)
    mats, texs = nbvh_b200.procedural_materials(int(max(b["material_ids"].max() for b in bl)) + 1)
    mod.set_scene(tl, ti, bl, materials=mats, textures=texs)
    cut = nbvh_b200.procedural_tree_cut(len(mod.neural_bvh_nodes()), seed=11)
    cut_dev = torch.from_numpy(cut.view(np.uint8).reshape(-1).copy()).to(dev)
    out["sampler"] = {}
    for ahead in (4, 16):  # walk-ahead: one walk launch per `ahead` steps
        mod.sampler_walk_ahead(ahead, 1)
        for log2b in (18, 15):
            B = 1 << log2b
            nodes = torch.zeros(B, dtype=torch.int32, device=dev)
            tx_ = torch.zeros((B, 9), dtype=torch.float32, device=dev)
            tt_ = torch.zeros((B, 8), dtype=torch.float32, device=dev)
            step = [0]

            def gen():
                mod.optimisation_data_generation(step[0], B, 0, nodes, 1.0, cut_dev, tx_, tt_)
                step[0] += 1
            ms = timeit(gen, iters=2 * ahead, warmup=ahead)
            out["sampler"]["2^%d_walk_ahead_%d" % (log2b, ahead)] = {"ms_per_step": ms, "Msamples_s": B / ms / 1e3}
    mod.sampler_walk_ahead(1, 1)
    for log2b in (18, 16, 15):
        B = 1 << log2b
        nodes = torch.zeros(B, dtype=torch.int32, device=dev)
        tx_ = torch.zeros((B, 9), dtype=torch.float32, device=dev)
        tt_ = torch.zeros((B, 8), dtype=torch.float32, device=dev)
        step = [0]

        def gen():
            mod.optimisation_data_generation(step[0], B, 0, nodes, 1.0, cut_dev, tx_, tt_)
            step[0] += 1
        mod.reset_counters()
        ms = timeit(gen, iters=20, warmup=3)
        c = mod.counters()
        out["sampler"]["2^%d" % log2b] = {"ms": ms, "Msamples_s": B / ms / 1e3, "nbvh_steps_per_slot": c["nbvh_steps"] / 23.0 / B,
                                          "blas_steps_per_slot": c["blas_steps"] / 23.0 / B}
        if log2b == 18:
            ms_fb = timeit(lambda: mod.training_forward_backward(tx_, tt_), iters=20)
            ms_opt = timeit(lambda: mod.optimizer_step(), iters=20)

            def full():
                gen()
                mod.training_forward_backward(tx_, tt_)
                mod.optimizer_step()
            ms_full = timeit(full, iters=20)
            out["train_2p18"] = {"forward_backward_ms": ms_fb, "optimizer_ms": ms_opt, "generation_plus_step_ms": ms_full,
                                 "Msamples_s": B / ms_full / 1e3}
        if log2b == 15:
            ms_fb = timeit(lambda: mod.training_forward_backward(tx_, tt_, global_batch=1 << 18), iters=20)
            out["train_2p15"] = {"forward_backward_ms": ms_fb}
    print(json.dumps(out), flush=True)
    mod.close()

    # ---- path configs
    if not quick and nbvh_b200.has_shading_fixture(root):
        for name, (W, H, depth, it) in {"config2b_1080p_depth3": (1920, 1080, 3, 20), "config4_2160p_depth4": (3840, 2160, 4, 8)}.items():
            pmod = nbvh_b200.load_chess_module(max_depth=depth, envmap_eval=True, shading=True)
            pmod.attach_framebuffer((W, H))
            s = [0]

            def fr_():
                pmod.neural_render(s[0], 0)
                s[0] += 1
            ms = timeit(fr_, iters=it)
            out[name] = {"ms_per_frame": ms, "Mpixels_s": W * H / ms / 1e3}
            pmod.close()
    txt = json.dumps(out, indent=1)
    print(txt)
    if len(sys.argv) > 1 and not sys.argv[1].startswith("--"):
        open(os.path.join(REPO, "gpurun_out", sys.argv[1]), "w").write(txt)


if __name__ == "__main__":
    main()
