#!/usr/bin/env python3
"""Kernel sequence for ncu captures (GPU box): warm-up, then 2 headline frames and 2 optimisation steps with sample generation.
usage under ncu: ncu --set full --import-source on -k regex:'k_' --launch-skip <n> ... python tools/ncu_driver.py"""
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
mod.reinitialise_inference_data((1920, 1080))
tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root, with_neural_mesh=True, with_attributes=True)
mats, texs = nbvh_b200.procedural_materials(int(max(b["material_ids"].max() for b in bl)) + 1)
mod.set_scene(tl, ti, bl, materials=mats, textures=texs)
cut = nbvh_b200.procedural_tree_cut(len(mod.neural_bvh_nodes()), seed=11)
cut_dev = torch.from_numpy(cut.view(np.uint8).reshape(-1).copy()).to(dev)
B = 1 << 18
nodes = torch.zeros(B, dtype=torch.int32, device=dev)
tx = torch.zeros((B, 9), dtype=torch.float32, device=dev)
tt = torch.zeros((B, 8), dtype=torch.float32, device=dev)
WARM = int(os.environ.get("NCU_WARM", "2"))
for i in range(WARM + 2):
    mod.neural_render(i, 0)
torch.cuda.synchronize()
for s in range(WARM + 2):
    mod.optimisation_data_generation(s, B, 0, nodes, 1.0, cut_dev, tx, tt)
    mod.training_forward_backward(tx, tt)
    mod.optimizer_step()
torch.cuda.synchronize()
mod.close()
