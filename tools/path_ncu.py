#!/usr/bin/env python3
"""One warm multi-bounce frame under ncu (launch list): python tools/path_ncu.py W H depth"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import nbvh_b200  # noqa: E402

W, H, depth = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mod = nbvh_b200.load_chess_module(max_depth=depth, envmap_eval=True, shading=True)
mod.attach_framebuffer((W, H))
for s in range(3):
    mod.neural_render(s, 0)
torch.cuda.synchronize()
mod.close()
