#!/usr/bin/env python3
"""LDR-FLIP of a 1080p Chess frame under ncu: python tools/flip_ncu.py   (three calls; capture the last with -k regex:k_flip -s 6 -c 3)"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import nbvh_b200  # noqa: E402

mod = nbvh_b200.load_chess_module(max_depth=1, envmap_eval=True)
accum, _ = mod.attach_framebuffer((1920, 1080))
mod.neural_render(0, 0)
torch.cuda.synchronize()
ref = (torch.roll(accum, 2, dims=1) * 0.9 + 0.01).contiguous()
for _ in range(3):
    print(mod.flip_error(ref, 1))
mod.close()
