#!/usr/bin/env python3
"""Golden vectors from the reference's own tcnn CUDA path (TEST INFRASTRUCTURE; run ON THE GPU BOX under gpurun).

Drives oracle/_ref/tcnn_harness (the reference's tiny-cuda-nn fork + PathTracingComposite loss rebuilt for sm_100 from
/root/reference by oracle/Makefile.tcnn) on seeded inputs and writes compact goldens to gpurun_out/tcnn_golden/;
copy them to tests/golden/ and commit.  The inputs are regenerated from the seeds by the tests (golden_inputs())."""
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
HARNESS = os.path.join(REPO, "oracle", "_ref", "tcnn_harness")
N_MLP = 64 * 96 + 3 * 64 * 64 + 16 * 64


def grid_entries(log2_hashmap):
    tot = 0
    for l in range(8):
        res = int(np.ceil(2.0 ** l * 8 - 1)) + 1
        n = (res ** 3 + 7) // 8 * 8
        tot += min(n, 1 << log2_hashmap)
    return tot


def golden_inputs(log2_hashmap, n_query, batch, n_batches, seed):
    """Deterministic parameters / queries / training batches shared by the generator and the tests."""
    rng = np.random.default_rng(seed)
    n_params = N_MLP + 4 * grid_entries(log2_hashmap)
    p = np.zeros(n_params, np.float16)
    p[:N_MLP] = (rng.standard_normal(N_MLP) * 0.12).astype(np.float16)
    p[N_MLP:] = (rng.standard_normal(n_params - N_MLP) * 0.5).astype(np.float16)
    q = rng.uniform(0, 1, size=(n_query, 9)).astype(np.float32)
    tin = rng.uniform(0, 1, size=(n_batches, batch, 9)).astype(np.float32)
    tgt = rng.uniform(0, 1, size=(n_batches, batch, 8)).astype(np.float32)
    vis = (rng.uniform(0, 1, size=(n_batches, batch)) < 0.5)
    tgt[..., 0] = vis
    tgt[~vis, 1:] = 0.0
    return p.view(np.uint16), q, tin, tgt


def run_case(name, log2_hashmap, n_query, batch, n_batches, steps, seed, params=None, keep=True, curve_only=False):
    wd = os.path.join(REPO, "gpurun_out", "tcnn_work_" + name)
    os.makedirs(wd, exist_ok=True)
    p, q, tin, tgt = golden_inputs(log2_hashmap, n_query, batch, n_batches, seed)
    if params is not None:
        p = params
    p.tofile(os.path.join(wd, "params_init.u16"))
    q.tofile(os.path.join(wd, "query_in.f32"))
    tin.tofile(os.path.join(wd, "train_in.f32"))
    tgt.tofile(os.path.join(wd, "train_tgt.f32"))
    out = subprocess.run([HARNESS, wd, str(log2_hashmap), str(n_query), str(batch), str(steps)], capture_output=True, text=True)
    print(out.stdout[-600:], out.stderr[-2000:])
    if out.returncode != 0:
        raise SystemExit("tcnn_harness failed for " + name)
    timing = json.load(open(os.path.join(wd, "timing.json")))
    if not keep:
        return timing
    gold = dict(
        meta=np.array([log2_hashmap, n_query, batch, n_batches, steps, seed], dtype=np.int64),
        query_out=np.fromfile(os.path.join(wd, "query_out.f32"), dtype="<f4").reshape(n_query, 8),
    )
    if steps and curve_only:  # long runs: the loss curve is the golden
        gold = dict(meta=gold["meta"], train_loss=np.fromfile(os.path.join(wd, "train_loss.f32"), dtype="<f4"))
    elif steps:
        gold.update(train_loss=np.fromfile(os.path.join(wd, "train_loss.f32"), dtype="<f4"),
                    train_L0=np.fromfile(os.path.join(wd, "train_L0.f32"), dtype="<f4").reshape(batch, 16),
                    train_out0=np.fromfile(os.path.join(wd, "train_out0.u16"), dtype="<u2").reshape(batch, 16),
                    grads0=np.fromfile(os.path.join(wd, "grads0.u16"), dtype="<u2"),
                    params_after1=np.fromfile(os.path.join(wd, "params_after1.u16"), dtype="<u2"),
                    params_after=np.fromfile(os.path.join(wd, "params_after.u16"), dtype="<u2"))
    gd = os.path.join(REPO, "gpurun_out", "tcnn_golden")
    os.makedirs(gd, exist_ok=True)
    np.savez_compressed(os.path.join(gd, "tcnn_%s.npz" % name), **gold)
    return timing


def main():
    timings = {}
    if "--long-only" in sys.argv:
        # four runs on identical inputs: tcnn's fp16 gradient atomics are unordered, so its own curve varies from run to run --
        # the spread of these runs is the band the B200 path is compared with
        curves = []
        for r in range(4):
            timings["long"] = run_case("long", 12, 256, 2048, 8, 1000, 20241018, curve_only=True)
            curves.append(np.load(os.path.join(REPO, "gpurun_out", "tcnn_golden", "tcnn_long.npz"))["train_loss"])
        g = dict(np.load(os.path.join(REPO, "gpurun_out", "tcnn_golden", "tcnn_long.npz")))
        g["train_loss_runs"] = np.stack(curves)
        np.savez_compressed(os.path.join(REPO, "gpurun_out", "tcnn_golden", "tcnn_long.npz"), **g)
        print(json.dumps(timings))
        return
    # small network (2^12 table): inference + 30 optimisation steps over 4 rotating batches of 2048 samples
    timings["small"] = run_case("small", 12, 2048, 2048, 4, 30, 20241017)
    # the same network for 1,000 steps (8 rotating batches): long-horizon loss curve of tcnn's Trainer (fp16 gradients, Adam, decay)
    timings["long"] = run_case("long", 12, 256, 2048, 8, 1000, 20241018, curve_only=True)
    # shipped Chess weights: inference only
    import nbvh_b200
    root = nbvh_b200.chess_fixture_dir()
    if root:
        from oracle import oracle as O
        params = O.read_network_json(os.path.join(root, nbvh_b200.scene_io.CHESS_BASE + "_network.json"))
        timings["chess"] = run_case("chess", 18, 4096, 256, 1, 0, 7, params=params)
        # throughput of the reference network path at the benchmark sizes (not stored as goldens)
        timings["chess_1M_queries_2p18_batch"] = run_case("perf", 18, 1 << 20, 1 << 18, 1, 20, 9, params=params, keep=False)
    json.dump(timings, open(os.path.join(REPO, "gpurun_out", "tcnn_golden", "tcnn_timings.json"), "w"), indent=1)
    print(json.dumps(timings))


if __name__ == "__main__":
    main()
