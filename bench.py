#!/usr/bin/env python3
"""bench.py -- headline benchmark of the N-BVH neural ray-query hot path on B200 ("Measurement").

A step = one 1-spp pass of the hot path over one frame of primary rays of the shipped Chess configuration
(1920x1080, shipped camera / .nbvh / weights): ray generation, TLAS + N-BVH walk, hash-grid encoding, 64-wide MLP,
decode, closest-hit bookkeeping, classical BLAS closest hit, splat -- two kernels per frame: k_primary (camera ray, TLAS,
classical BLAS, N-BVH walk to the first leaf interval, ballot-compacted neural ray queue) and the persistent tcgen05
k_neural (encode + MLP + decode + walk resume per queued ray).
  value  = rays / device time of the kernel region (inputs resident in HBM, CUDA events, max over ranks)
  e2e    = same metric through the C-ABI host call nbvh_render_host: camera constants H2D, 8-bit frame D2H
  N > 1  = the partition the north star names: the frame's 8x4-pixel tiles are dealt round-robin to the ranks, every rank renders
           N samples of ITS tiles (pixels / N x N samples = one frame's worth of rays per rank: weak scaling) and the 8-bit tiles are
           summed into the full frame by NCCL inside the timed region (the library's own communicator, on a second stream)
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


def flat(prefix, d, out):
    """Flattens the numeric leaves of a nested dict into out['prefix_key'] (the driver keeps the scalar members of `config`)."""
    for k, v in d.items():
        if isinstance(v, dict):
            flat(prefix + "_" + k, v, out)
        elif isinstance(v, (int, float)) and not isinstance(v, bool):
            out[prefix + "_" + k] = float(v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the live run of the reference's own GPU path (oracle/_ref harnesses)")
    ap.add_argument("--no-path", action="store_true", help="skip the hybrid path tracing blocks (configs 2b / 4)")
    ap.add_argument("--no-train", action="store_true", help="skip the online-optimisation block (config 3)")
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
                "visibility/distance/normal/albedo queries, 1920x1080 primary rays (max_depth 1)")
    if root is None:
        workload = "SYNTHETIC stand-in (Chess fixture absent): procedural scene, shipped network shape, 1920x1080 primary rays"

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
    if world > 1:
        mod.setup_data_parallel(rank, world)  # the library's own NCCL communicator: tile gathers and gradient exchange
    accum, ldr = mod.attach_framebuffer((W, H))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    comm_stream = torch.cuda.Stream(device=dev)
    staging = [torch.zeros_like(ldr) for _ in range(2)] if world > 1 else None

    ready_ev = [torch.cuda.Event(), torch.cuda.Event()]
    done_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def frame_step(i):
        """One step of the headline stream.  N = 1: one 1-spp frame.  N > 1: the N-sample frame i, this rank's interleaved 8x4-pixel
        tiles (pixels / N x N samples), then the 8-bit tiles of all ranks are summed into the full frame (disjoint tiles: a sum of
        32-bit words) by NCCL on the communication stream -- the gather of frame i overlaps the kernels of frame i + 1 (two staging
        frames; a staging frame is reused only after its allreduce has finished)."""
        if world == 1:
            mod.neural_render(i, 0, stream=stream.cuda_stream)
            return
        mod.neural_render_samples(0, world, i, rank, world, stream=stream.cuda_stream)
        b = i & 1
        stream.wait_event(done_ev[b])
        staging[b].copy_(ldr, non_blocking=True)
        ready_ev[b].record(stream)
        comm_stream.wait_event(ready_ev[b])
        mod.allreduce(staging[b], stream=comm_stream.cuda_stream)
        done_ev[b].record(comm_stream)

    # ---------------- device-resident timing: K steps, CUDA events on the launching stream
    for i in range(args.warmup):
        flush.fill_(i & 0xFF)
        frame_step(i)
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
        frame_step(args.warmup + i)
        evs[i][1].record(stream)
    stream.wait_stream(comm_stream)
    ev_end = torch.cuda.Event(enable_timing=True)
    ev_end.record(stream)
    barrier()
    t_wall = time.time() - t_wall0
    sampler.stop_flag = True
    sampler.join()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    # the last gather ends after the last step's event: it is charged to the timed total
    total_ms = float(sum(step_ms)) + max(0.0, float(evs[-1][1].elapsed_time(ev_end)))
    counters = mod.counters()
    prim_ms_total, neur_ms_total, stage_frames = mod.stage_times()
    mod.enable_stage_timing(0)

    # ---------------- strong scaling of ONE frame (N > 1): a single 1-spp frame cut into interleaved tiles, blocking per frame
    strong = None
    if world > 1:
        def strong_frame(i):
            mod.neural_render_shard(i, 0, rank, world, stream=stream.cuda_stream)
            staging[0].copy_(ldr, non_blocking=True)  # ldr keeps this rank's tiles only (zeros elsewhere); the sum lands in the staging frame
            mod.allreduce(staging[0], stream=stream.cuda_stream)
        stream.wait_stream(comm_stream)
        for i in range(3):
            strong_frame(i)
        barrier()
        n_s = max(5, min(args.steps, 50))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(n_s):
            strong_frame(3 + i)
        e1.record(stream)
        barrier()
        strong = e0.elapsed_time(e1) / n_s

    # ---------------- end to end through the C-ABI host call (pinned host buffers, copies inside the timed region)
    ldr_host = torch.empty(n_rays * 4, dtype=torch.uint8, pin_memory=True).numpy()
    cam = mod.get_constants(nbvh_b200.CONST_CAMERA)
    e2e_sync_s = None
    if world == 1:
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
        e2e_h2d, e2e_d2h = 48, n_rays * 4
        e2e_api = ("nbvh_render_host_submit / _wait (C ABI), four independent frames in flight, one stream and queue per slot: camera "
                   "constants in, sRGB8 frame out, pinned host memory")
    else:
        # N > 1: per step the camera block goes to every rank's device, the ranks render their tiles, NCCL sums the tiles, and rank 0
        # reads the full 8-bit frame back into pinned host memory (the other ranks return nothing)
        cam_host = torch.frombuffer(bytearray(bytes(cam)), dtype=torch.uint8).pin_memory()
        cam_dev = torch.empty_like(cam_host, device=dev)
        ldr_pinned = torch.empty(n_rays * 4, dtype=torch.uint8, pin_memory=True)

        def e2e_frame(i):
            cam_dev.copy_(cam_host, non_blocking=True)
            mod.set_constants(nbvh_b200.CONST_CAMERA, cam)
            mod.neural_render_samples(0, world, i, rank, world, stream=stream.cuda_stream)
            staging[0].copy_(ldr, non_blocking=True)
            mod.allreduce(staging[0], stream=stream.cuda_stream)
            if rank == 0:
                ldr_pinned.copy_(staging[0].view(-1), non_blocking=True)
        for i in range(2):
            e2e_frame(i)
        barrier()
        t0 = time.time()
        for i in range(args.steps):
            e2e_frame(args.warmup + i)
        torch.cuda.synchronize()
        e2e_s = time.time() - t0
        e2e_h2d, e2e_d2h = len(bytes(cam)), n_rays * 4
        e2e_api = ("NeuralPathModule.neural_render_samples on the rank's tiles + nbvh_allreduce_buffer (C ABI, NCCL) + rank 0 reads the "
                   "full sRGB8 frame into pinned host memory; camera constants uploaded every step")

    # ---------------- the same frame with 8 samples per launch pair (nbvh_render_samples): k_neural's drain of in-flight rays
    # is paid once per launch instead of once per sample; same images up to the fp32 sum order
    ms_n, multi_ms, ms_reps = 8, None, 0
    if world == 1:
        for i in range(2):
            mod.neural_render_samples(i * ms_n, ms_n, stream=stream.cuda_stream)
        barrier()
        ms_reps = max(3, min(args.steps // ms_n, 20))
        ms_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ms_reps)]
        for i in range(ms_reps):
            flush.fill_(i & 0xFF)
            ms_ev[i][0].record(stream)
            mod.neural_render_samples((2 + i) * ms_n, ms_n, stream=stream.cuda_stream)
            ms_ev[i][1].record(stream)
        barrier()
        multi_ms = float(sum(a.elapsed_time(b) for a, b in ms_ev)) / ms_reps

    # ---------------- online optimisation (second half of BASELINE's metric: train samples/s), data parallel inside the library:
    # every rank owns 2^18 batch slots; per step: sample generation (K10: tree-cut walk + neural-mesh BLAS trace + material lookups),
    # hash-grid encode + fused MLP forward/loss/backward + grid backward, the gradient exchange (comm.cu: two NCCL buckets on the
    # library's communication stream, the MLP bucket overlapping the grid backward) and the replicated Adam step
    train = None
    if not args.no_train:
        train_B = 1 << 18
        if root is not None:
            tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root, with_neural_mesh=True, with_attributes=True)
            mats, texs = nbvh_b200.procedural_materials(int(max(b["material_ids"].max() for b in bl)) + 1)
            mod.set_scene(tl, ti, bl, materials=mats, textures=texs)
        cut = nbvh_b200.procedural_tree_cut(len(mod.neural_bvh_nodes()), seed=11)
        cut_dev = torch.from_numpy(cut.view(np.uint8).reshape(-1).copy()).to(dev)
        in_dims = int(mod.get_constants(nbvh_b200.CONST_MODULE).input_network_dims)
        out_dims = int(mod.network_output_dims())
        nodes_dev = torch.zeros(train_B, dtype=torch.int32, device=dev)
        tx = torch.zeros((train_B, in_dims), dtype=torch.float32, device=dev)
        tt = torch.zeros((train_B, out_dims), dtype=torch.float32, device=dev)
        walk_ahead = 16
        train_steps = max(walk_ahead, min(args.steps, 48) // walk_ahead * walk_ahead)  # whole walk-ahead groups inside the timed region
        step_no = [0]
        mod.sampler_walk_ahead(walk_ahead, world)  # one walk launch serves 16 steps of this rank (training_step = step * world + rank)

        # second set of sample buffers + a generation stream: the samples of a step depend on the tree-cut and the RNG only (not on the
        # parameters), so step k + 1 is generated while step k runs forward / backward / gradient exchange / Adam -- the same batches
        # in the same order as the sequential loop (tools/train_overlap_probe.py); at N > 1 this hides the wait for the NCCL buckets
        nodes_dev2 = torch.zeros(train_B, dtype=torch.int32, device=dev)
        tx2 = torch.zeros((train_B, in_dims), dtype=torch.float32, device=dev)
        tt2 = torch.zeros((train_B, out_dims), dtype=torch.float32, device=dev)
        gen_stream = torch.cuda.Stream(device=dev)
        buf_ready = [torch.cuda.Event(), torch.cuda.Event()]
        buf_free = [torch.cuda.Event(), torch.cuda.Event()]

        def run_steps(B_local, B_global, n, what="step"):
            bufs = [(tx[:B_local], tt[:B_local], nodes_dev[:B_local]), (tx2[:B_local], tt2[:B_local], nodes_dev2[:B_local])]
            if what != "step":  # "generate": sample generation alone; "sequential": one stream (per-kernel stage timing, kernels alone)
                sx, st_, sn = bufs[0]
                for _ in range(n):
                    mod.optimisation_data_generation(step_no[0] * world + rank, B_local, 0, sn, 1.0, cut_dev, sx, st_, stream=stream.cuda_stream)
                    step_no[0] += 1
                    if what == "sequential":
                        mod.training_forward_backward(sx, st_, global_batch=B_global, stream=stream.cuda_stream)
                        mod.optimizer_step(stream=stream.cuda_stream)
                return
            gen_stream.wait_stream(stream)
            for b in range(2):
                buf_free[b].record(stream)

            def generate(k):
                # rank r samples its own slots of the global batch: the RNG is keyed by training_step * batch + slot
                sx, st_, sn = bufs[k & 1]
                gen_stream.wait_event(buf_free[k & 1])
                mod.optimisation_data_generation((step_no[0] + k) * world + rank, B_local, 0, sn, 1.0, cut_dev, sx, st_,
                                                 stream=gen_stream.cuda_stream)
                buf_ready[k & 1].record(gen_stream)

            generate(0)
            for k in range(n):
                if k + 1 < n:
                    generate(k + 1)
                sx, st_, sn = bufs[k & 1]
                stream.wait_event(buf_ready[k & 1])
                mod.training_forward_backward(sx, st_, global_batch=B_global, stream=stream.cuda_stream)
                mod.optimizer_step(stream=stream.cuda_stream)
                buf_free[k & 1].record(stream)
            step_no[0] += n

        def time_steps(B_local, B_global, what="step"):
            run_steps(B_local, B_global, (walk_ahead - step_no[0] % walk_ahead) % walk_ahead + walk_ahead, what)  # warm-up to a group boundary
            barrier()
            te0, te1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            te0.record(stream)
            run_steps(B_local, B_global, train_steps, what)
            te1.record(stream)
            barrier()
            return te0.elapsed_time(te1) / train_steps

        train_ms = time_steps(train_B, train_B * world)
        seq_ms = time_steps(train_B, train_B * world, what="sequential")
        mod.enable_training_stage_timing(train_steps)
        run_steps(train_B, train_B * world, train_steps, what="sequential")
        torch.cuda.synchronize()
        stage, stage_n = mod.training_stage_times()
        mod.enable_training_stage_timing(0)
        gen_ms = time_steps(train_B, train_B * world, what="generate")
        # strong scaling of the same step: the GLOBAL batch stays 2^18, every rank owns 2^18 / world slots
        strong_ms, strong_stage = train_ms, None
        if world > 1:
            strong_ms = time_steps(train_B // world, train_B)
            mod.enable_training_stage_timing(train_steps)
            run_steps(train_B // world, train_B, train_steps, what="sequential")
            torch.cuda.synchronize()
            strong_stage, sn_ = mod.training_stage_times()
            mod.enable_training_stage_timing(0)
            strong_stage = {k: v / max(sn_, 1) for k, v in strong_stage.items()}
            strong_gen = time_steps(train_B // world, train_B, what="generate")
            strong_stage["sample_generation"] = strong_gen
        train = dict(train_B=train_B, train_ms=train_ms, seq_ms=seq_ms, gen_ms=gen_ms, strong_ms=strong_ms, strong_stage=strong_stage,
                     stage={k: v / max(stage_n, 1) for k, v in stage.items()}, steps=train_steps)

    # ---------------- hybrid neural path tracing (BASELINE configs[1] at the shipped max_depth = "2b", and configs[3]):
    # shading + BSDF sampling + Russian roulette + HDRI per bounce, 3 launches per depth (path.cu)
    path_info = None
    if not args.no_path and root is not None and nbvh_b200.has_shading_fixture(root):
        path_info = {}
        pmod = nbvh_b200.load_chess_module(device=local_rank, max_depth=3, envmap_eval=True, shading=True)
        if world > 1:
            pmod.share_data_parallel(mod)  # the same NCCL communicator as the headline module

        def time_frames(fn, n):
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
            for i in range(n):
                flush.fill_(i & 0xFF)
                ev[i][0].record(stream)
                fn(i)
                ev[i][1].record(stream)
            barrier()
            return float(sum(a.elapsed_time(b) for a, b in ev)) / n

        # config 2b: 1920x1080, shipped max_depth 3.  N = 1: one sample per step; N > 1: the N-sample frame, tile-partitioned like the
        # headline (one frame's worth of camera paths per rank), 8-bit tiles summed by NCCL
        acc2, ldr2 = pmod.attach_framebuffer((W, H))

        def frame2(i):
            if world == 1:
                pmod.neural_render(i, 0, stream=stream.cuda_stream)
            else:
                pmod.neural_render_samples(0, world, i, rank, world, stream=stream.cuda_stream)
                staging[0].copy_(ldr2, non_blocking=True)
                pmod.allreduce(staging[0], stream=stream.cuda_stream)
        for i in range(3):
            frame2(i)
        barrier()
        pmod.reset_counters()
        n2 = max(5, min(args.steps, 30))
        ms2 = time_frames(lambda i: frame2(3 + i), n2)
        c2 = {k: v / n2 for k, v in pmod.counters().items()}
        # config 4: 3840x2160, max_depth 4, 16 spp; image tiles (8x4-pixel blocks) are dealt round-robin to the ranks
        # (nbvh_render_interleaved: balances the board-vs-pieces cost skew), all 16 samples of a rank's tiles in ONE launch sequence at
        # every N (pixels x 16 path slots keep the persistent kernels full), the RGBA32F tiles summed by NCCL to the full image
        W4, H4, spp4 = 3840, 2160, 16
        pmod.max_depth = 4
        accum4, _ = pmod.attach_framebuffer((W4, H4))
        n_pix4 = W4 * H4
        spl4 = spp4

        def image4(_i):
            if world > 1:
                accum4.zero_()  # the in-place sum of the previous image left every rank's tiles here: start from this rank's tiles only
            pmod.neural_render_samples(0, spp4, 0, rank, world, stream=stream.cuda_stream)
            if world > 1:
                pmod.allreduce(accum4, stream=stream.cuda_stream)

        image4(0)
        barrier()
        pmod.reset_counters()
        n4 = 3
        ms4 = time_frames(image4, n4)
        c4 = {k: v / n4 for k, v in pmod.counters().items()}
        pmod.close()
        tt_ = torch.tensor([ms2, ms4, c4["rays"], c4["queries"], c2["rays"], c2["queries"]], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = tt_.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt_, op=dist.ReduceOp.SUM)
            ms2, ms4 = float(tmax[0]), float(tmax[1])
            c4["rays"], c4["queries"], c2["rays"], c2["queries"] = float(tt_[2]), float(tt_[3]), float(tt_[4]), float(tt_[5])
        path_info["config2b"] = {
            "workload": "Chess hybrid neural path tracing, shipped settings (max_depth 3, Russian roulette from depth 2, HDRI, chessboard glTF "
                        "material), 1920x1080; %s" % ("1 spp per step" if world == 1 else
                                                      "%d spp per step, interleaved 8x4-pixel tiles over %d ranks + NCCL tile sum" % (world, world)),
            "ms_per_step": ms2, "frames": n2, "scaling": "weak",
            "Mpixel_samples_s": world * n_rays / (ms2 * 1e-3) / 1e6, "rays_per_step": c2["rays"],
            "Mrays_s_all_rays": c2["rays"] / (ms2 * 1e-3) / 1e6, "queries_per_step": c2["queries"],
            "neural_queries_per_s": c2["queries"] / (ms2 * 1e-3), "launches_per_frame": 9}
        path_info["config4"] = {
            "workload": "Chess + shipped HDRI, max_depth 4, 3840x2160, 16 spp; 8x4-pixel tiles dealt round-robin to %d rank(s), all 16 samples "
                        "of a rank's tiles in one launch sequence (the same at every N), RGBA32F tiles summed by NCCL" % world,
            "scaling": "strong", "samples_per_launch": spl4,
            "ms_per_16spp_image": ms4, "images": n4, "Mpixel_samples_s": n_pix4 * spp4 / (ms4 * 1e-3) / 1e6,
            "rays_per_image": c4["rays"], "Mrays_s_all_rays": c4["rays"] / (ms4 * 1e-3) / 1e6,
            "neural_queries_per_s": c4["queries"] / (ms4 * 1e-3)}

    vals = [total_ms, e2e_s * 1e3, train["train_ms"] if train else 0.0, train["strong_ms"] if train else 0.0,
            (e2e_sync_s or 0.0) * 1e3, multi_ms or 0.0, strong or 0.0]
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, train_ms, strong_train_ms, e2e_sync_ms, multi_ms_r, strong = (float(t[k]) for k in range(7))
    # counters over all ranks (queries, rays ... of the whole job)
    ckeys = ["rays", "queries", "nbvh_steps", "blas_steps", "tri_tests", "queued_rays", "nbvh_steps_neural", "tile_iterations"]
    ct = torch.tensor([float(counters[k]) for k in ckeys], dtype=torch.float64, device=dev)
    st_t = torch.tensor([prim_ms_total / max(1, stage_frames), neur_ms_total / max(1, stage_frames)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ct, op=dist.ReduceOp.SUM)
        dist.all_reduce(st_t, op=dist.ReduceOp.MAX)

    if rank == 0:
        K = args.steps
        ms_per_step = total_ms / K
        value = world * n_rays * K / (total_ms * 1e-3) / 1e6
        e2e_value = world * n_rays * K / (e2e_ms * 1e-3) / 1e6
        peak, peak_src = measured_peaks()
        per = {k: float(ct[i]) / K / world for i, k in enumerate(ckeys)}  # per launch of ONE rank
        # algorithmic bytes per launch: 64 B per interior step (child pair), 56 B per triangle test,
        # 48 B per ray (ray in + hit out), 1600 B per network query (1536 B table gathers + 64 B state)
        alg_bytes = 64.0 * (per["nbvh_steps"] + per["blas_steps"]) + 56.0 * per["tri_tests"] + 48.0 * per["rays"] + 1600.0 * per["queries"]
        neural_bytes = 1600.0 * per["queries"] + 64.0 * per["nbvh_steps_neural"] + 48.0 * per["queued_rays"]
        primary_bytes = alg_bytes - neural_bytes
        primary_ms, neural_ms = float(st_t[0]), float(st_t[1])
        achieved = neural_bytes / (neural_ms * 1e-3) / 1e9
        tensor_peak = None
        pk = os.path.join(REPO, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            tensor_peak = json.load(open(pk)).get("bf16_tflops_sustained")
        tensor_peak = tensor_peak or 1387.0
        tensor_tflops = 37888.0 * per["queries"] / (neural_ms * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(REPO, "profiles", "k_neural_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        config = {"workload": workload + (", 1 spp per step" if world == 1 else
                                          ", %d spp per step: 8x4-pixel tiles dealt round-robin to %d ranks, every rank renders %d samples of its "
                                          "tiles, 8-bit tiles summed by NCCL inside the timed region" % (world, world, world)),
                  "rays_per_step_per_gpu": n_rays,
                  "sharding": ("single GPU" if world == 1 else "image tiles (interleaved 8x4-pixel blocks) x samples; scene and weights replicated; "
                               "one NCCL allreduce of the 8.3 MB sRGB8 frame per step on the library's communicator, overlapped with the next step"),
                  "l2": "256 MiB buffer written between timed steps (outside the event intervals)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 traversal + f16 hash grid/MLP operands with f32 accumulate",
            "data": "shipped Chess scene + shipped pre-trained weights" if root else "synthetic",
            "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": e2e_d2h, "api": e2e_api},
            "gpu_launches": (2 if world == 1 else 4) * K,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "nbvh::k_neural", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": neural_bytes, "kernel_ms": neural_ms,
                         "tensor_tflops": tensor_tflops, "tensor_peak_tflops": tensor_peak, "tensor_frac": tensor_tflops / tensor_peak,
                         "k_primary_ms": primary_ms, "k_primary_alg_bytes": primary_bytes,
                         "k_primary_gbs": primary_bytes / (primary_ms * 1e-3) / 1e9 if primary_ms else None,
                         "k_primary_frac": primary_bytes / (primary_ms * 1e-3) / 1e9 / peak if primary_ms else None,
                         "note": "hash table (10.8 MB), tree-cut and queue are L2-resident by design; the algorithmic bytes are "
                                 "L2-level gathers, compared with the measured HBM copy bandwidth as SURVEY 8(d) prescribes"},
            "frame_alg_bytes": alg_bytes,
            "counters_per_step_per_gpu": per,
            "neural_queries_per_s": world * per["queries"] / (ms_per_step * 1e-3),
            "clocks": sampler.summary(),
            "wall_s_timed_region": t_wall,
        }
        if world == 1:
            line["e2e"].update({"frames_in_flight": 4,
                                "note": "back-to-back frames (no L2 flush between them, unlike `value`): the copy is hidden behind the next frame's kernels",
                                "synchronous_value": n_rays * K / (e2e_sync_ms * 1e-3) / 1e6,
                                "synchronous_api": "nbvh_render_host (one blocking call per frame, the reference's calling pattern)"})
            line["multi_sample"] = {"samples_per_launch": ms_n, "ms_per_launch": multi_ms_r, "ms_per_sample_frame": multi_ms_r / ms_n,
                                    "value": n_rays * ms_n / (multi_ms_r * 1e-3) / 1e6, "unit": UNIT, "launches": ms_reps}
            config["multi_sample_8spl_Mrays_s"] = n_rays * ms_n / (multi_ms_r * 1e-3) / 1e6
            config["e2e_blocking_call_Mrays_s"] = n_rays * K / (e2e_sync_ms * 1e-3) / 1e6
        else:
            line["strong_scaling_one_frame"] = {
                "what": "ONE 1-spp 1080p frame cut into interleaved 8x4-pixel tiles over the ranks (nbvh_render_interleaved) + NCCL tile sum, "
                        "blocking per frame: per-frame latency", "ms_per_frame": strong, "Mrays_s": n_rays / (strong * 1e-3) / 1e6}
            config["strong_one_frame_ms"] = strong
            config["strong_one_frame_Mrays_s"] = n_rays / (strong * 1e-3) / 1e6
        if path_info:
            line["path"] = path_info
            config["config2b_ms_per_step"] = path_info["config2b"]["ms_per_step"]
            config["config2b_Mpixel_samples_s"] = path_info["config2b"]["Mpixel_samples_s"]
            config["config4_16spp_image_ms"] = path_info["config4"]["ms_per_16spp_image"]
        if train:
            B = train["train_B"]
            stg = train["stage"]
            n_par = float(mod.n_params())
            kern = {
                "sample_generation": {"ms": train["gen_ms"], "what": "K10 per step: k_node_keep + k_sample_targets, plus 1/16 of a k_sample_walk "
                                                                      "launch that walks the slots of 16 steps (nbvh_sampler_walk_ahead)"},
                "k_encode_features": {"ms": stg["encode"], "bound": "hbm", "achieved": 1536.0 * B / (stg["encode"] * 1e-3) / 1e9, "peak": peak,
                                      "unit": "GB/s", "algorithmic": "1536 B of table gathers per sample"},
                "k_train_mlp": {"ms": stg["mlp"], "bound": "tensor", "achieved": 113664.0 * B / (stg["mlp"] * 1e-3) / 1e12, "peak": tensor_peak,
                                "unit": "TFLOP/s", "algorithmic": "113,664 FLOP per sample (forward + dgrad + wgrad)"},
                "k_grid_backward": {"ms": stg["grid_backward"], "bound": "hbm", "achieved": 1536.0 * B / (stg["grid_backward"] * 1e-3) / 1e9,
                                    "peak": peak, "unit": "GB/s", "algorithmic": "1536 B of scatter-adds per sample (shared-memory + L2 reductions)"},
                "k_adam": {"ms": stg["adam"], "bound": "hbm", "achieved": 36.0 * n_par / (stg["adam"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                           "algorithmic": "36 B per parameter (dense upper bound; untouched grid entries are skipped)"},
                "gradient_exchange_wait": {"ms": stg["exchange_wait"], "what": "time the compute stream waits for the two NCCL buckets"},
            }
            for k_, v_ in kern.items():
                if "achieved" in v_:
                    v_["frac"] = v_["achieved"] / v_["peak"]
            line["train"] = {"metric": "train samples/s", "value": world * B / (train_ms * 1e-3), "unit": "samples/s",
                             "ms_per_step": train_ms, "batch_per_gpu": B, "steps": train["steps"], "scaling": "weak",
                             "sample_generation_ms": train["gen_ms"], "ms_per_step_one_stream": train["seq_ms"], "roofline": kern,
                             "strong_scaling": {"global_batch": B, "ms_per_step": strong_train_ms, "samples_s": B / (strong_train_ms * 1e-3),
                                                "per_kernel_ms": train["strong_stage"]},
                             "what": "per step: sample generation on the Chess tree-cut + neural-mesh BLAS (K10, procedural materials; generated for step k + 1 on a second stream while step k trains: same batches, same order), "
                                     "hash-grid encode + fused tcgen05 MLP forward/loss/backward + grid backward + Adam" +
                                     ("; gradient exchange inside the library (NCCL, two buckets)" if world > 1 else "")}
            config["train_samples_s"] = world * B / (train_ms * 1e-3)
            config["train_ms_per_step"] = train_ms
            config["train_strong_2p18_ms_per_step"] = strong_train_ms
            line["roofline"].update({"k_train_mlp_frac": kern["k_train_mlp"]["frac"], "k_train_mlp_ms": stg["mlp"],
                                     "k_encode_features_frac": kern["k_encode_features"]["frac"], "k_encode_features_ms": stg["encode"],
                                     "k_grid_backward_frac": kern["k_grid_backward"]["frac"], "k_grid_backward_ms": stg["grid_backward"],
                                     "k_adam_frac": kern["k_adam"]["frac"], "k_adam_ms": stg["adam"], "sample_generation_ms": train["gen_ms"]})
        # ---------------- the reference's OWN GPU path, live on this box (rank 0, outside every timed region)
        if world == 1 and not args.no_reference_gpu and root is not None:
            sys.path.insert(0, os.path.join(REPO, "tools"))
            try:
                import ref_gpu_arm
                mod.close()
                mod = None
                torch.cuda.empty_cache()
                ref = ref_gpu_arm.measure(root, sampler_cls=ClockSampler, device_index=local_rank, frames=max(5, min(K, 20)),
                                          do_path=path_info is not None, do_train=train is not None)
                line["reference_gpu_b200"] = ref
                a = ref.get("config2a")
                if a:
                    config["reference_gpu_live_ms_per_frame"] = a["ms_mean"]
                    config["reference_gpu_live_Mrays_s"] = a["Mrays_s_mean"]
                    config["speedup_vs_reference_gpu"] = value / a["Mrays_s_mean"]
                    line["e2e"]["reference_gpu_live_Mrays_s"] = a["Mrays_s_mean"]
                    line["e2e"]["speedup_vs_reference_gpu_blocking_call"] = (n_rays * K / (e2e_sync_ms * 1e-3) / 1e6) / a["Mrays_s_mean"]
                if path_info and "config2b" in ref:
                    path_info["config2b"]["reference_gpu_live_ms_per_frame"] = ref["config2b"]["ms_mean"]
                    config["config2b_reference_gpu_live_ms"] = ref["config2b"]["ms_mean"]
                if path_info and "config4_1spp" in ref:
                    path_info["config4"]["reference_gpu_live_ms_per_16spp_image"] = 16.0 * ref["config4_1spp"]["ms_mean"]
                    config["config4_reference_gpu_live_16spp_ms"] = 16.0 * ref["config4_1spp"]["ms_mean"]
                if train and "train_2p18" in ref:
                    line["train"]["reference_gpu_live"] = ref["train_2p18"]
                    config["train_reference_gpu_live_samples_s"] = ref["train_2p18"]["samples_s"]
            except Exception as ex:  # the harness is absent (built only where /root/reference is mounted) or failed: say so, never fake it
                line["reference_gpu_b200"] = {"live": False, "error": str(ex)[:300]}
        if not args.no_cpu_baseline and world == 1:
            v, cores, kind, sample, _, _ = cpu_reference_run(5, root, W, H)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if mod is not None:
        mod.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
