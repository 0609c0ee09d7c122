#!/usr/bin/env python3
"""bench.py -- headline benchmark of the N-BVH neural ray-query hot path on B200 ("Measurement").

A step = one 1-spp pass of the hot path over one frame of primary rays of the shipped Chess configuration
(1920x1080, shipped camera / .nbvh / weights): ray generation, TLAS + N-BVH walk, hash-grid encoding, 64-wide MLP,
decode, closest-hit bookkeeping, classical BLAS closest hit, splat -- two kernels per frame: k_primary (camera ray, TLAS,
classical BLAS, N-BVH walk to the first leaf interval, ballot-compacted neural ray queue) and the persistent tcgen05
k_neural (encode + MLP + decode + walk resume per queued ray).
  value  = rays / device time of the kernel region (inputs resident in HBM, CUDA events, max over ranks)
  e2e    = same metric through the C-ABI host call nbvh_render_host: camera constants H2D, 8-bit frame D2H
  N > 1  = every rank renders its own sample of the frame (spp sharding, no data-path collective): weak scaling
--impl reference times the reference's CPU path (ext/bvh closest-hit traversal of the same primary rays, the neural
mesh traced as triangles) on the host cores through oracle/_ref/cpu_bvh_trace (or the oracle port when absent).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
# stdout carries exactly one JSON line: everything native libraries print (NCCL's version banner goes to fd 1) is sent to stderr
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

METRIC = "neural ray queries/sec (Mrays/s, 1080p Chess)"
UNIT = "Mrays/s"
CPU_TOOL = os.path.join(REPO, "oracle", "_ref", "cpu_bvh_trace")


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:  # NVML: sub-millisecond samples
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self.stop_flag:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for nme, b in bits.items():
                    if r & b:
                        self.reasons.add(nme)
                time.sleep(0.002)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nme, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(nme)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_reference_run(repeats, root, w, h):
    """The reference's CPU path on the host cores.  Returns (Mrays/s, cores, kind, sample description)."""
    import numpy as np
    if os.path.exists(CPU_TOOL) and root is not None:
        out = subprocess.run([CPU_TOOL, root, str(w), str(h), "0", "0", str(repeats)], capture_output=True, text=True,
                             check=True).stdout.strip().splitlines()[-1]
        j = json.loads(out)
        return j["mrays_s"], j["threads"], "reference", ("full %dx%d frame of primary rays, closest hit through ext/bvh "
                                                         "Bvh::intersect (all meshes as triangles), best of %d, BVH build %.2f s "
                                                         "excluded" % (w, h, repeats, j["build_s"])), j["best_s"] * 1e3, j.get("rep_s")
    # oracle port: neural primary pipeline of the oracle on a pixel sample (OpenMP)
    from oracle import oracle as O
    import nbvh_b200
    cfg, nbvh, tlas, tlas_idx, blas = nbvh_b200.synthetic_scene(fb=(w, h))
    const = O.config_constants(cfg, nbvh)
    rng = np.random.default_rng(1337)
    net0 = O.make_net(cfg, np.zeros(16, np.uint16))
    total, _ = O.expected_n_params(net0)
    net = O.make_net(cfg, (rng.standard_normal(total) * 0.05).astype(np.float16).view(np.uint16))
    sc = O.Scene(tlas, tlas_idx, blas)
    stride = max(1, (w * h) // 200000)
    best = 1e30
    for _ in range(max(1, repeats)):
        t0 = time.time()
        hits, _ = O.render_primary(sc, nbvh, const, net, 0, w * h, stride=stride)
        best = min(best, time.time() - t0)
    return len(hits) / best / 1e6, os.cpu_count(), "port", "every %d-th pixel of a %dx%d synthetic frame, oracle neural pipeline" % (
        stride, w, h), best * 1e3, None


def sweep(args, rank, local_rank, world, dev):
    """BASELINE configs[4]: synthetic procedurally instanced ~10M-triangle scene, fixed-depth-17 neural BLAS, random weights
    (throughput only), rays from the 1.5x box surface toward interior points; Mrays/s per batch size."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import nbvh_b200
    t0 = time.time()
    cfg, nbvh, tlas, tlas_idx, blas, info = nbvh_b200.synthetic_instanced_scene(device=dev, return_mesh=(rank == 0))
    build_s = time.time() - t0
    mod = nbvh_b200.NeuralPathModule(cfg, device=local_rank)
    blas_build = None
    if rank == 0:  # device LBVH build of the full 10 M-triangle classical BLAS (nbvh_build_blas): the host stage it replaces takes seconds
        try:
            mv, mi = info.pop("mesh_vertices"), info.pop("mesh_indices")
            mod.build_blas(mv, mi)
            torch.cuda.synchronize()
            tb = time.time()
            _, _, depth = mod.build_blas(mv, mi)
            torch.cuda.synchronize()
            blas_build = {"triangles": int(mi.shape[0]), "ms": (time.time() - tb) * 1e3, "tree_depth": depth}
            del mv, mi
        except Exception as ex:  # a tree deeper than the traversal stack is refused
            blas_build = {"error": str(ex)}
    mod.set_neural_bvh(nbvh)
    mod.set_scene(tlas, tlas_idx, blas)
    rng = np.random.default_rng(1337)
    n_mlp = 64 * 96 + 3 * 64 * 64 + 16 * 64
    p = np.zeros(mod.n_params(), np.float16)  # spread weights: visibility is not pinned at the 0.5 threshold (tcnn's init gives sigmoid(~0))
    p[:n_mlp] = (rng.standard_normal(n_mlp) * 0.2).astype(np.float16)
    p[n_mlp:] = (rng.standard_normal(len(p) - n_mlp) * 0.7).astype(np.float16)
    mod.load_params(p.view(np.uint16))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    results = []
    reps = max(3, min(args.steps, 10))
    for log2n in (16, 18, 20, 22, 24):
        n_total = 1 << log2n
        n_local = n_total // world
        o, d = nbvh_b200.synthetic_ray_batch(nbvh[0], n_total, seed=42, device=dev)
        o, d = o[rank * n_local:(rank + 1) * n_local].contiguous(), d[rank * n_local:(rank + 1) * n_local].contiguous()
        hits = torch.empty(n_local * nbvh_b200.HIT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        for _ in range(max(args.warmup, 3)):
            mod.trace_rays(o, d, 1, hits, stream=stream.cuda_stream)
        barrier()
        mod.reset_counters()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for i in range(reps):
            flush.fill_(i & 0xFF)
            ev[i][0].record(stream)
            mod.trace_rays(o, d, 1, hits, stream=stream.cuda_stream)
            ev[i][1].record(stream)
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev) / reps
        c = mod.counters()
        kinds = np.bincount(hits.cpu().numpy().view(nbvh_b200.HIT_DTYPE)["valid_hit"], minlength=3)
        t = torch.tensor([ms, c["queries"] / reps, c["nbvh_steps"] / reps, float(kinds[1])], dtype=torch.float64, device=dev)
        if world > 1:
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms = float(tm[0])
        results.append({"log2_batch": log2n, "rays": n_total, "ms": ms, "Mrays_s": n_total / ms / 1e3, "queries_per_ray": float(t[1]) / n_total,
                        "nbvh_steps_per_ray": float(t[2]) / n_total, "neural_hit_fraction": float(t[3]) / n_total})
        del o, d, hits
    if rank == 0:
        best = max(results, key=lambda r: r["Mrays_s"])
        line = {"metric": "neural ray queries/sec (Mrays/s, synthetic ~10M-triangle scene, ray-batch sweep)", "value": best["Mrays_s"], "unit": UNIT,
                "n_gpus": world, "steps": reps, "warmup": max(args.warmup, 3), "ms_per_step": best["ms"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32 traversal + f16 hash grid/MLP operands with f32 accumulate",
                "data": "synthetic (procedural scene, random weights: throughput only)",
                "config": {"workload": "BASELINE configs[4]: %d instances x %d triangles = %.2f M triangles on a jittered lattice, flattened into one "
                                       "neural mesh; N-BVH = fixed-depth-17 cut (%d nodes, %d neural leaves, %d-%d triangles per leaf); hash grid 8x4 2^18 + "
                                       "64-wide MLP; rays from the 1.5x box surface toward uniform interior points (seed 42); value = best batch size"
                                       % (info["instances"], info["triangles_per_instance"], info["triangles"] / 1e6, info["nbvh_nodes"],
                                          info["neural_leaves"], info["triangles_per_leaf_min"], info["triangles_per_leaf_max"]),
                           "sharding": "each batch split evenly across the ranks, scene and weights replicated, no collective",
                           "l2": "256 MiB buffer written between timed repetitions", "scene_build_s": build_s},
                "sweep": results, "blas_build": blas_build, "api": "nbvh_trace_rays (C ABI)", "gpu_launches": 4 * reps}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    mod.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", action="store_true",
                    help="BASELINE configs[4] instead of the headline: synthetic ~10M-triangle instanced scene, deep neural BLAS "
                         "(depth-17 tree-cut, 131k neural leaves), ray-batch sweep 2^16..2^24, batches split evenly across the ranks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import nbvh_b200
    root = nbvh_b200.chess_fixture_dir()
    W, H = 1920, 1080
    workload = ("Chess N-BVH inference, shipped config nodes_32940_hash_18_lvl_8 (hash grid 8x4 2^18 + 64-wide MLP), "
                "visibility/distance/normal/albedo queries, 1920x1080 1spp primary rays (max_depth 1)")
    if root is None:
        workload = "SYNTHETIC stand-in (Chess fixture absent): procedural scene, shipped network shape, 1920x1080 1spp primary rays"

    if args.impl == "reference":
        if rank != 0:
            return 0
        # one process: BVH build once (excluded), then warmup + steps full-frame traversals, each timed
        info = cpu_reference_run(args.warmup + args.steps, root, W, H)
        if info[5]:
            vals = [x * 1e3 for x in info[5][args.warmup:]]
        else:
            vals = [info[4]]
        ms = sum(vals) / max(1, len(vals))
        v = W * H / (ms * 1e-3) / 1e6 if info[2] == "reference" else info[0]
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "shipped Chess scene" if root else "synthetic",
                "config": {"workload": workload + " -- reference arm: CPU closest-hit traversal of the non-neural BLAS (BASELINE config 1)"},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": info[1], "kind": info[2], "sample": info[3]},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: nbvh_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.sweep:
        return sweep(args, rank, local_rank, world, dev)

    if root is not None:
        mod = nbvh_b200.load_chess_module(device=local_rank, max_depth=1, envmap_eval=False)
    else:
        cfg, nbvh, tlas, tlas_idx, blas = nbvh_b200.synthetic_scene(fb=(W, H), cut_depth=9)
        mod = nbvh_b200.NeuralPathModule(cfg, device=local_rank)
        mod.set_neural_bvh(nbvh)
        mod.set_scene(tlas, tlas_idx, blas)
        mod.init_params(1337)
    n_rays = W * H

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing: K launches of k_render, CUDA events on the launching stream
    stream = torch.cuda.current_stream()
    for i in range(args.warmup):
        flush.fill_(i & 0xFF)
        mod.neural_render(i * world + rank, 0, stream=stream.cuda_stream)
    barrier()
    mod.reset_counters()
    mod.enable_stage_timing(args.steps)  # event triples around k_primary / k_neural inside the timed region, no host sync
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.time()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # L2 flush between timed iterations (outside the event interval)
        evs[i][0].record(stream)
        mod.neural_render((args.warmup + i) * world + rank, 0, stream=stream.cuda_stream)
        evs[i][1].record(stream)
    barrier()
    t_wall = time.time() - t_wall0
    sampler.stop_flag = True
    sampler.join()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(step_ms))
    counters = mod.counters()
    prim_ms_total, neur_ms_total, stage_frames = mod.stage_times()
    mod.enable_stage_timing(0)

    # ---------------- end to end through the C-ABI host call (pinned host buffers, copies inside the timed region)
    ldr_host = torch.empty(n_rays * 4, dtype=torch.uint8, pin_memory=True).numpy()
    cam = mod.get_constants(nbvh_b200.CONST_CAMERA)
    for i in range(2):
        mod.neural_render_host(0, i, ldr_host=ldr_host, camera=cam)
    barrier()
    t0 = time.time()
    for i in range(args.steps):
        mod.neural_render_host(0, args.warmup + i, ldr_host=ldr_host, camera=cam)
    torch.cuda.synchronize()
    e2e_sync_s = time.time() - t0
    # the same call with frames in flight (nbvh_render_host_submit / _wait): the 8.3 MB device-to-host copy of frame i overlaps
    # the kernels of the next frames; every step still uploads its camera block and reads its frame back into pinned host memory.
    # Independent frames (sample 0 of a new frame each step, RNG offset = frame number): every slot renders on its own stream with
    # its own queue, so k_primary of the following frames overlaps the drain of frame i's k_neural
    # (tools/frames_in_flight.py: 1,915 / 2,697 / 2,835 / 2,849 Mrays/s at 1 / 2 / 3 / 4 frames in flight; identical frames)
    n_fif = 4
    ldr_hosts = [ldr_host] + [torch.empty(n_rays * 4, dtype=torch.uint8, pin_memory=True).numpy() for _ in range(n_fif - 1)]
    for i in range(n_fif):
        mod.neural_render_host_submit(0, i, ldr_hosts[i], i, camera=cam)
    for i in range(n_fif):
        mod.neural_render_host_wait(i)
    barrier()
    t0 = time.time()
    for i in range(args.steps):
        if i >= n_fif:
            mod.neural_render_host_wait(i % n_fif)
        mod.neural_render_host_submit(0, args.warmup + i, ldr_hosts[i % n_fif], i % n_fif, camera=cam)
    for i in range(min(n_fif, args.steps)):
        mod.neural_render_host_wait(i)
    torch.cuda.synchronize()
    e2e_s = time.time() - t0

    # ---------------- the same frame with 8 samples per launch pair (nbvh_render_samples): k_neural's drain of in-flight rays (60 % of
    # the 1-spp launch) is paid once per launch instead of once per sample; same images up to the fp32 sum order
    ms_n = 8
    for i in range(2):
        mod.neural_render_samples(i * ms_n, ms_n, stream=stream.cuda_stream)
    barrier()
    ms_reps = max(3, min(args.steps // ms_n, 20))
    ms_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ms_reps)]
    for i in range(ms_reps):
        flush.fill_(i & 0xFF)
        ms_ev[i][0].record(stream)
        mod.neural_render_samples(((2 + i) * world + rank) * ms_n, ms_n, stream=stream.cuda_stream)
        ms_ev[i][1].record(stream)
    barrier()
    multi_ms = float(sum(a.elapsed_time(b) for a, b in ms_ev)) / ms_reps

    # ---------------- hybrid neural path tracing (BASELINE configs[1] at the shipped max_depth = "2b", and configs[3]):
    # shading + BSDF sampling + Russian roulette + HDRI per bounce, 3 launches per depth (path.cu)
    path_info = None
    if root is not None and nbvh_b200.has_shading_fixture(root):
        path_info = {}
        pmod = nbvh_b200.load_chess_module(device=local_rank, max_depth=3, envmap_eval=True, shading=True)

        def time_frames(fn, n):
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
            for i in range(n):
                flush.fill_(i & 0xFF)
                ev[i][0].record(stream)
                fn(i)
                ev[i][1].record(stream)
            barrier()
            return float(sum(a.elapsed_time(b) for a, b in ev)) / n

        # config 2b: 1920x1080, shipped max_depth 3, one sample per rank per step (spp sharding like the headline)
        pmod.attach_framebuffer((W, H))
        for i in range(3):
            pmod.neural_render(i * world + rank, 0, stream=stream.cuda_stream)
        barrier()
        pmod.reset_counters()
        n2 = max(5, min(args.steps, 30))
        ms2 = time_frames(lambda i: pmod.neural_render((3 + i) * world + rank, 0, stream=stream.cuda_stream), n2)
        c2 = {k: v / n2 for k, v in pmod.counters().items()}
        # the same with 4 samples of the frame per launch sequence (nbvh_render_samples)
        for i in range(2):
            pmod.neural_render_samples(4 * i, 4, stream=stream.cuda_stream)
        barrier()
        ms2_multi = time_frames(lambda i: pmod.neural_render_samples(4 * ((2 + i) * world + rank), 4, stream=stream.cuda_stream), max(3, n2 // 4)) / 4.0
        # config 4: 3840x2160, max_depth 4, 16 spp; image tiles (8x4-pixel blocks) are dealt round-robin to the ranks
        # (nbvh_render_interleaved: balances the board-vs-pieces cost skew), each rank accumulates its tiles over 16 samples,
        # one NCCL reduce of the RGBA32F image to rank 0 at the end (the tiles are disjoint, the other ranks contribute zeros)
        W4, H4, spp4 = 3840, 2160, 16
        pmod.max_depth = 4
        accum4, _ = pmod.attach_framebuffer((W4, H4))
        n_pix4 = W4 * H4
        img4 = torch.empty_like(accum4) if world > 1 else accum4

        # samples per launch sequence: as many as keep pixels x samples near one 4K frame of path slots per rank
        spl4 = min(spp4, max(4, world))

        def image4(_i):
            for s_ in range(0, spp4, spl4):
                pmod.neural_render_samples(s_, spl4, 0, rank, world, stream=stream.cuda_stream)
            if world > 1:
                img4.copy_(accum4)
                dist.reduce(img4, dst=0)

        image4(0)
        barrier()
        pmod.reset_counters()
        n4 = 3
        ms4 = time_frames(image4, n4)
        c4 = {k: v / n4 for k, v in pmod.counters().items()}
        pmod.close()
        tt_ = torch.tensor([ms2, ms4, c4["rays"], c4["queries"], ms2_multi], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = tt_.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt_, op=dist.ReduceOp.SUM)
            ms2, ms4, ms2_multi = float(tmax[0]), float(tmax[1]), float(tmax[4])
            c4["rays"], c4["queries"] = float(tt_[2]), float(tt_[3])
        path_info["config2b"] = {
            "workload": "Chess hybrid neural path tracing, shipped settings (max_depth 3, Russian roulette from depth 2, HDRI, chessboard glTF "
                        "material), 1920x1080, 1 spp per rank per step", "ms_per_frame": ms2, "frames": n2,
            "Mpixels_s": world * n_rays / (ms2 * 1e-3) / 1e6, "rays_per_frame": c2["rays"],
            "Mrays_s_all_rays": world * c2["rays"] / (ms2 * 1e-3) / 1e6, "queries_per_frame": c2["queries"],
            "neural_queries_per_s": world * c2["queries"] / (ms2 * 1e-3), "launches_per_frame": 9,
            "multi_sample": {"samples_per_launch": 4, "ms_per_sample_frame": ms2_multi, "Mpixels_s": world * n_rays / (ms2_multi * 1e-3) / 1e6}}
        path_info["config4"] = {
            "workload": "Chess + shipped HDRI, max_depth 4, 3840x2160, 16 spp; 8x4-pixel tiles dealt round-robin to %d rank(s), "
                        "RGBA32F image reduced to rank 0; %d sample(s) per launch sequence (nbvh_render_samples)" % (world, spl4), "scaling": "strong",
            "ms_per_16spp_image": ms4, "images": n4, "Mpixel_samples_s": n_pix4 * spp4 / (ms4 * 1e-3) / 1e6,
            "rays_per_image": c4["rays"], "Mrays_s_all_rays": c4["rays"] / (ms4 * 1e-3) / 1e6,
            "neural_queries_per_s": c4["queries"] / (ms4 * 1e-3)}
        try:
            rp_ = json.load(open(os.path.join(REPO, "tests", "golden", "ref_chess_path_timing.json")))
            a, b = rp_["config2b_1080p_depth3"], rp_["config4_2160p_depth4"]
            path_info["config2b"]["reference_b200"] = {"ms_per_frame": a["ms_best"], "Mpixels_s": n_rays / a["ms_best"] / 1e3, "queries": a["queries"],
                                                       "launches": a["launches"], "host_syncs": a["host_syncs"], "what": a["what"]}
            path_info["config4"]["reference_b200"] = {"ms_per_1spp_frame": b["ms_best"], "ms_per_16spp_image": 16 * b["ms_best"],
                                                      "Mpixel_samples_s": n_pix4 / b["ms_best"] / 1e3, "launches_per_frame": b["launches"],
                                                      "host_syncs_per_frame": b["host_syncs"], "what": b["what"] + " (single GPU: the reference has no multi-GPU mode)"}
        except Exception:
            pass

    # ---------------- online optimisation (second half of BASELINE's metric: train samples/s), data parallel:
    # every rank owns 2^18 batch slots; per step: sample generation (K10: tree-cut walk + neural-mesh BLAS trace +
    # material lookups), hash-grid encode + fused MLP forward/loss/backward + grid backward, fp32 gradient allreduce
    # (21.6 MB, NCCL) and the replicated Adam step
    train_B = 1 << 18
    if root is not None:
        tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root, with_neural_mesh=True, with_attributes=True)
        mats, texs = nbvh_b200.procedural_materials(int(max(b["material_ids"].max() for b in bl)) + 1)
        mod.set_scene(tl, ti, bl, materials=mats, textures=texs)
    cut = nbvh_b200.procedural_tree_cut(len(mod.neural_bvh_nodes()), seed=11 + rank)
    cut_dev = torch.from_numpy(cut.view(np.uint8).reshape(-1).copy()).to(dev)
    nodes_dev = torch.zeros(train_B, dtype=torch.int32, device=dev)
    tx = torch.zeros((train_B, 9), dtype=torch.float32, device=dev)
    tt = torch.zeros((train_B, 8), dtype=torch.float32, device=dev)
    grads = mod.gradient_buffer()
    train_steps = max(5, min(args.steps, 50))
    step_no = [0]

    def generate():
        # rank r samples the slots [r*B, (r+1)*B) of the global batch: the RNG is keyed by step * global batch + slot
        mod.optimisation_data_generation(step_no[0] * world + rank, train_B, 0, nodes_dev, 1.0, cut_dev, tx, tt, stream=stream.cuda_stream)
        step_no[0] += 1

    def train_step():
        generate()
        mod.training_forward_backward(tx, tt, global_batch=train_B * world, stream=stream.cuda_stream)
        if world > 1:
            dist.all_reduce(grads)
        mod.optimizer_step(stream=stream.cuda_stream)

    for _ in range(3):
        train_step()
    barrier()
    te0, te1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    te0.record(stream)
    for _ in range(train_steps):
        train_step()
    te1.record(stream)
    barrier()
    train_ms = te0.elapsed_time(te1) / train_steps
    te0.record(stream)
    for _ in range(train_steps):
        generate()
    te1.record(stream)
    barrier()
    gen_ms = te0.elapsed_time(te1) / train_steps

    # strong scaling of the same step: the GLOBAL batch stays 2^18, every rank owns 2^18 / world slots
    strong_ms = train_ms
    if world > 1:
        sB = train_B // world
        sx, st_, sn = tx[:sB], tt[:sB], nodes_dev[:sB]

        def strong_step():
            mod.optimisation_data_generation(step_no[0] * world + rank, sB, 0, sn, 1.0, cut_dev, sx, st_, stream=stream.cuda_stream)
            step_no[0] += 1
            mod.training_forward_backward(sx, st_, global_batch=train_B, stream=stream.cuda_stream)
            dist.all_reduce(grads)
            mod.optimizer_step(stream=stream.cuda_stream)

        for _ in range(3):
            strong_step()
        barrier()
        te0.record(stream)
        for _ in range(train_steps):
            strong_step()
        te1.record(stream)
        barrier()
        strong_ms = te0.elapsed_time(te1) / train_steps

    t = torch.tensor([total_ms, e2e_s * 1e3, train_ms, strong_ms, e2e_sync_s * 1e3, multi_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, train_ms, strong_ms, e2e_sync_ms, multi_ms = (float(t[k]) for k in range(6))

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * n_rays * args.steps / (total_ms * 1e-3) / 1e6
        e2e_value = world * n_rays * args.steps / (e2e_ms * 1e-3) / 1e6
        peak, peak_src = measured_peaks()
        K = args.steps
        per = {k: v / K for k, v in counters.items()}
        # algorithmic bytes per launch: 64 B per interior step (child pair), 56 B per triangle test,
        # 48 B per ray (ray in + hit out), 1600 B per network query (1536 B table gathers + 64 B state)
        alg_bytes = 64.0 * (per["nbvh_steps"] + per["blas_steps"]) + 56.0 * per["tri_tests"] + 48.0 * per["rays"] + 1600.0 * per["queries"]
        # per kernel: k_neural (dominant) = queries + its share of the N-BVH walk + 48 B per queued ray;
        # k_primary = everything else
        neural_bytes = 1600.0 * per["queries"] + 64.0 * per["nbvh_steps_neural"] + 48.0 * per["queued_rays"]
        primary_bytes = alg_bytes - neural_bytes
        neural_ms = neur_ms_total / max(1, stage_frames)
        primary_ms = prim_ms_total / max(1, stage_frames)
        achieved = neural_bytes / (neural_ms * 1e-3) / 1e9
        tensor_peak = None
        pk = os.path.join(REPO, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            tensor_peak = json.load(open(pk)).get("bf16_tflops_sustained")
        tensor_tflops = 37888.0 * per["queries"] / (neural_ms * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(REPO, "profiles", "k_neural_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        ref_gpu, ref_train = None, None
        gp = os.path.join(REPO, "tests", "golden")
        try:  # the reference's own kernels + tcnn (sm_100 build) timed on a B200 by tools/make_ref_*_goldens.py
            rg = json.load(open(os.path.join(gp, "ref_chess_timing.json")))
            ref_gpu = {"ms_per_frame": rg["ms_best"], "Mrays_s": rg["mrays_s_best"], "queries": rg["queries"], "launches": rg["launches"],
                       "host_syncs": rg["host_syncs"], "gpu": rg.get("gpu"), "what": rg.get("what")}
            rs = json.load(open(os.path.join(gp, "ref_chess_sampler_timing.json")))
            rt = json.load(open(os.path.join(gp, "tcnn_timings.json")))
            tstep = rt["chess_1M_queries_2p18_batch"]["training_step_ms"]
            ref_train = {"sample_generation_ms": rs["ms_per_step"], "training_step_ms": tstep,
                         "samples_s": train_B / ((rs["ms_per_step"] + tstep) * 1e-3),
                         "what": "reference granular_sample_bvh_network_input_data + tcnn training_step, batch 2^18, one B200"}
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 traversal + f16 hash grid/MLP operands with f32 accumulate",
            "data": "shipped Chess scene + shipped pre-trained weights" if root else "synthetic",
            "config": {"workload": workload, "rays_per_step_per_gpu": n_rays,
                       "sharding": "one full-frame sample per rank per step (spp sharding), scene and weights replicated",
                       "l2": "256 MiB buffer written between timed steps (outside the event intervals)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 48, "d2h_bytes_per_step": n_rays * 4,
                    "api": "nbvh_render_host_submit / _wait (C ABI), four independent frames in flight, one stream and queue per slot: camera constants in, sRGB8 frame out, pinned host memory", "frames_in_flight": 4,
                    "note": "back-to-back frames (no L2 flush between them, unlike `value`): the copy is hidden behind the next frame's kernels",
                    "synchronous_value": world * n_rays * args.steps / (e2e_sync_ms * 1e-3) / 1e6,
                    "synchronous_api": "nbvh_render_host (one blocking call per frame, the reference's calling pattern)"},
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "nbvh::k_neural", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": neural_bytes, "kernel_ms": neural_ms,
                         "note": "hash table (10.8 MB), tree-cut and queue are L2-resident by design; the algorithmic bytes are "
                                 "L2-level gathers, compared with the measured HBM copy bandwidth as SURVEY 8(d) prescribes"},
            "roofline_tensor": {"bound": "tensor", "achieved": tensor_tflops, "peak": tensor_peak, "unit": "TFLOP/s",
                                "frac": (tensor_tflops / tensor_peak) if tensor_peak else None, "kernel": "nbvh::k_neural",
                                "flop_per_query": 37888},
            "roofline_primary": {"bound": "hbm", "kernel": "nbvh::k_primary", "kernel_ms": primary_ms,
                                 "algorithmic_bytes_per_launch": primary_bytes,
                                 "achieved": primary_bytes / (primary_ms * 1e-3) / 1e9 if primary_ms else None, "peak": peak,
                                 "unit": "GB/s", "frac": primary_bytes / (primary_ms * 1e-3) / 1e9 / peak if primary_ms else None},
            "frame_alg_bytes": alg_bytes,
            "reference_gpu_b200": ref_gpu,
            "train": {"metric": "train samples/s", "value": world * train_B / (train_ms * 1e-3), "unit": "samples/s",
                      "ms_per_step": train_ms, "batch_per_gpu": train_B, "steps": train_steps,
                      "sample_generation_ms": gen_ms,
                      "strong_scaling": {"global_batch": train_B, "ms_per_step": strong_ms, "samples_s": train_B / (strong_ms * 1e-3)},
                      "what": "per step: sample generation on the Chess tree-cut + neural-mesh BLAS (K10, procedural materials), "
                              "hash-grid encode + fused tcgen05 MLP forward/loss/backward + grid backward + Adam" +
                              ("; fp32 gradient allreduce (NCCL) per step" if world > 1 else ""),
                      "reference_b200": ref_train},
            "multi_sample": {"samples_per_launch": ms_n, "ms_per_launch": multi_ms, "ms_per_sample_frame": multi_ms / ms_n,
                             "value": world * n_rays * ms_n / (multi_ms * 1e-3) / 1e6, "unit": UNIT, "launches": ms_reps,
                             "what": "the headline workload with 8 samples of the frame per launch pair (nbvh_render_samples): k_primary over "
                                     "pixels x samples, one k_neural drain per launch, 16-byte reductions into the accumulation; L2 flushed between launches"},
            "path": path_info,
            "counters_per_step": per,
            "neural_queries_per_s": per["queries"] / (ms_per_step * 1e-3),
            "clocks": sampler.summary(),
            "wall_s_timed_region": t_wall,
        }
        if not args.no_cpu_baseline and world == 1:
            v, cores, kind, sample, _, _ = cpu_reference_run(5, root, W, H)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    mod.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
