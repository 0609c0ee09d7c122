#!/usr/bin/env python3
"""Renders the Chess path-traced image of the goldens' view (256x144, shipped max_depth 3) at the goldens' converged sample count on
the GPU box and stores it under gpurun_out/ for tools/flip_report.py (FLIP needs the reference tree, which only the authoring
container has)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import nbvh_b200  # noqa: E402

g = np.load(os.path.join(REPO, "tests", "golden", "ref_chess_path_256x144.npz"))
W, H, spp = int(g["W"]), int(g["H"]), int(g["spp_converged"])
mod = nbvh_b200.load_chess_module(max_depth=int(g["max_depth"]), envmap_eval=True, shading=True)
accum, _ = mod.attach_framebuffer((W, H))
for s in range(spp):
    mod.neural_render(s, 0)
torch.cuda.synchronize()
img = (accum.cpu().numpy()[..., :3] / spp).astype(np.float32)
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
np.save(os.path.join(REPO, "gpurun_out", "b200_chess_path_256x144_%dspp.npy" % spp), img)
print("saved", img.shape, img.mean())
mod.close()
