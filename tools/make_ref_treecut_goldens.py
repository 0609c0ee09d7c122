#!/usr/bin/env python3
"""Goldens of the reference's OWN tree-cut learning kernels (TEST INFRASTRUCTURE; run ON THE GPU BOX): update_neural_tree_cut_loss and
update_neural_tree_cut_decisions (neural_bvh_renderer_kernels.cu:202-287), driven as neural_bvh_renderer.cu:199-236 drives them, on the
depth-13 cut of the Chess neural mesh's BLAS (nbvh_tree_cut_init(13): 16 k nodes; the shipped .nbvh is a cut of the AUTHORS' ext/bvh
build, which differs from the BLAS built here in ~10 % of the boxes, so it cannot be split further against this BLAS) with seeded per-node statistics, per-element losses and sampled node indices (treecut_inputs below).
Output: gpurun_out/ref_golden/ref_chess_treecut.npz -> copy to tests/golden/."""
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
OUT = os.path.join(REPO, "gpurun_out", "ref_golden")
BATCH, K, SEED, CUT_DEPTH = 1 << 16, 187, 424242, 13
CUT_DTYPE = np.dtype([("loss", "<f4"), ("iter_alive", "<u4"), ("depth", "<u4"), ("samples", "<u4")])


def treecut_inputs(nodes, batch=BATCH, seed=SEED):
    """Seeded inputs shared by the generator and the tests: statistics for every node, losses (batch x 16, 8 used), sampled leaves."""
    rng = np.random.default_rng(seed)
    n = len(nodes)
    cut = np.zeros(n, dtype=CUT_DTYPE)
    cut["loss"] = rng.uniform(0.0, 3.0, n).astype(np.float32)
    cut["iter_alive"] = rng.integers(1, 200, n)
    cut["depth"] = rng.integers(0, 20, n)
    cut["samples"] = rng.integers(1, 4000, n) * (rng.uniform(size=n) > 0.05)
    leaves = np.flatnonzero(nodes["b"] == 0xFFFFFFFF)
    sampled = rng.choice(leaves, batch).astype(np.uint32)
    L = np.zeros((batch, 16), np.float32)
    L[:, :8] = rng.uniform(0.0, 0.01, (batch, 8)).astype(np.float32)
    return cut, L, sampled


def main():
    import nbvh_b200
    import make_ref_render_goldens as R
    from nbvh_b200.scene_io import CHESS_BASE, chess_fixture_dir
    root = chess_fixture_dir()
    cfg = json.load(open(os.path.join(root, CHESS_BASE + "_neural.json")))
    mod = nbvh_b200.load_chess_module(max_depth=1, envmap_eval=False, with_neural_mesh=True)
    mod.tree_cut_init(CUT_DEPTH)
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root)
    wd = os.path.join(OUT, "work_treecut")
    R.prepare(mod, cfg, wd, 64, 32, (tl, ti, bl))
    nodes = mod.neural_bvh_nodes()
    cut, L, sampled = treecut_inputs(nodes)
    cut.tofile(os.path.join(wd, "tree_cut.bin"))
    L.tofile(os.path.join(wd, "tc_L.f32"))
    sampled.tofile(os.path.join(wd, "tc_sampled.u32"))
    out = subprocess.run([R.HARNESS, wd, "64", "32", "1", "18", "0", "treecut", str(BATCH), str(K)], capture_output=True, text=True)
    print(out.stdout[-800:], out.stderr[-2000:])
    if out.returncode != 0:
        raise SystemExit("ref_render_harness treecut failed")
    cut_out = np.fromfile(os.path.join(wd, "tc_cut_out.bin"), dtype=CUT_DTYPE)
    np.savez_compressed(os.path.join(OUT, "ref_chess_treecut.npz"), batch=BATCH, k=K, seed=SEED, cut_depth=CUT_DEPTH, loss=cut_out["loss"], samples=cut_out["samples"],
                        iter_alive=cut_out["iter_alive"], max_losses=np.fromfile(os.path.join(wd, "tc_max_losses.f32"), dtype="<f4"),
                        split=np.fromfile(os.path.join(wd, "tc_split.u32"), dtype="<u4"))
    for fn in os.listdir(wd):
        os.remove(os.path.join(wd, fn))
    mod.close()


if __name__ == "__main__":
    main()
