#!/usr/bin/env python3
"""Config 4 on one GPU (3840x2160, max_depth 4, 16 spp): samples per launch sequence sweep."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import nbvh_b200
mod = nbvh_b200.load_chess_module(max_depth=4, envmap_eval=True, shading=True)
mod.attach_framebuffer((3840, 2160))
for spl in (1, 2, 4, 8, 16):
    def image():
        for s in range(0, 16, spl):
            mod.neural_render_samples(s, spl)
    image()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        image()
    e1.record()
    torch.cuda.synchronize()
    print("samples per launch %2d: %.1f ms per 16-spp image" % (spl, e0.elapsed_time(e1) / 3), flush=True)
