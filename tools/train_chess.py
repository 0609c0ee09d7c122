#!/usr/bin/env python3
"""From-scratch online optimisation of the Chess N-BVH WITH tree-cut learning (BASELINE configs[2], tree_cut_learning: true):
root-only tree-cut, tcnn-style initial weights, the shipped split schedule (40 / 3000 / x2.5 / x2), batch 2^18, shipped Adam.
Per step: sample generation (K10) -> training step (encode, fused MLP fwd/loss/bwd, grid backward, Adam) -> tree-cut statistics,
top-k decisions and split (all on the device).  Prints a JSON summary: loss curve, tree growth, ms per step.
usage (GPU box): python tools/train_chess.py [steps] [out.json]"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import nbvh_b200  # noqa: E402
from nbvh_b200.scene_io import CHESS_BASE  # noqa: E402
from nbvh_b200.treecut import BvhSplitScheduler, TreeCutLearner  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(REPO, "gpurun_out", "train_chess.json")
    root = nbvh_b200.chess_fixture_dir()
    cfg = json.load(open(os.path.join(root, CHESS_BASE + "_neural.json")))
    mod = nbvh_b200.NeuralPathModule(json.dumps(cfg))
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root, with_neural_mesh=True, with_attributes=True)
    mats, texs = nbvh_b200.procedural_materials(int(max(b["material_ids"].max() for b in bl)) + 1)
    mod.set_scene(tl, ti, bl, materials=mats, textures=texs)
    mod.init_params(1337)
    mod.tree_cut_init(cfg["neural_bvh_renderer"].get("neural_bvh_min_depth", 0))
    sched = BvhSplitScheduler.from_config(cfg)
    sched.update_stats_and_reset(1, cfg["neural_bvh_renderer"]["max_training_steps"])
    B = cfg["neural_bvh_renderer"]["batch_size"]
    learner = TreeCutLearner(mod, sched, batch_size=B, num_lods_learned=cfg["neural_bvh_renderer"]["bvh_constants"]["num_lods_learned"])
    log = []
    torch.cuda.synchronize()
    t0 = time.time()
    t_last, loss_acc = t0, []
    for s in range(steps):
        loss, ns = learner.step()
        loss_acc.append(loss)
        if (s + 1) % 100 == 0 or s + 1 == steps:
            torch.cuda.synchronize()
            now = time.time()
            n_nodes = len(mod.neural_bvh_nodes())
            log.append({"step": s + 1, "mean_loss": float(np.mean(loss_acc)), "tree_cut_nodes": n_nodes,
                        "max_node_loss": learner.max_node_loss, "ms_per_step": (now - t_last) * 1e3 / len(loss_acc)})
            print(json.dumps(log[-1]), flush=True)
            t_last, loss_acc = now, []
    total = time.time() - t0
    # expected size of the cut after `steps` steps of the schedule
    chk = BvhSplitScheduler.from_config(cfg)
    expect, leaves_ok = 1, True
    for i in range(steps):
        k = chk.get_bvh_splits(i)
        expect += 2 * min(k, (expect + 1) // 2)
    res = {"steps": steps, "batch": B, "wall_s": total, "samples_per_s": steps * B / total, "log": log,
           "tree_cut_nodes": len(mod.neural_bvh_nodes()), "tree_cut_nodes_if_every_split_succeeds": expect,
           "tree_cut_max_depth": int(mod._lib.nbvh_tree_cut_max_depth(mod._h)),
           "what": "Chess from-scratch optimisation with tree-cut learning, shipped schedule + Adam, procedural materials, one B200; "
                   "wall clock includes the per-step host read-back of (loss, max node loss, splits)"}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "log"}))
    mod.close()


if __name__ == "__main__":
    main()
