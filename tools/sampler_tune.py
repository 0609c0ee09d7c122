#!/usr/bin/env python3
"""Sweep of the sample generator's launch tunables (GPU box): grid fractions, hard-slot threshold, lanes per hard slot."""
import itertools
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import nbvh_b200  # noqa: E402

root = nbvh_b200.chess_fixture_dir()
dev = torch.device("cuda", 0)
mod = nbvh_b200.load_chess_module(max_depth=1, envmap_eval=False)
tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root, with_neural_mesh=True, with_attributes=True)
mats, texs = nbvh_b200.procedural_materials(int(max(b["material_ids"].max() for b in bl)) + 1)
mod.set_scene(tl, ti, bl, materials=mats, textures=texs)
cut = nbvh_b200.procedural_tree_cut(len(mod.neural_bvh_nodes()), seed=11)
cut_dev = torch.from_numpy(cut.view(np.uint8).reshape(-1).copy()).to(dev)


def time_gen(B, iters=12):
    nodes = torch.zeros(B, dtype=torch.int32, device=dev)
    tx = torch.zeros((B, 9), dtype=torch.float32, device=dev)
    tt = torch.zeros((B, 8), dtype=torch.float32, device=dev)
    for s in range(3):
        mod.optimisation_data_generation(s, B, 0, nodes, 1.0, cut_dev, tx, tt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mod.reset_counters()
    e0.record()
    for s in range(iters):
        mod.optimisation_data_generation(3 + s, B, 0, nodes, 1.0, cut_dev, tx, tt)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, mod.counters()["nbvh_steps"] / iters / B


out = []
combos = [dict(NBVH_SAMPLER_HARD=h, NBVH_SAMPLER_GROUP=g, NBVH_SAMPLER_GRID_DIV=d, NBVH_SAMPLER_HGRID_DIV=hd)
          for h, g, d, hd in [(8, 8, 1, 1), (8, 8, 2, 1), (8, 8, 3, 1), (8, 8, 4, 1), (8, 8, 2, 2), (16, 8, 2, 2), (16, 16, 2, 2), (24, 16, 2, 2),
                              (32, 32, 2, 2), (32, 16, 3, 2), (16, 4, 2, 2), (12, 4, 3, 2), (1000000, 8, 2, 1), (1000000, 8, 4, 1)]]
for c in combos:
    for k, v in c.items():
        os.environ[k] = str(v)
    r = dict(c)
    for log2b in (18, 15):
        ms, steps = time_gen(1 << log2b)
        r["ms_2p%d" % log2b] = round(ms, 4)
        r["steps_per_slot_2p%d" % log2b] = round(steps, 1)
    out.append(r)
    print(json.dumps(r), flush=True)
mod.close()
