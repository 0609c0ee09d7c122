#!/usr/bin/env python3
"""Development probe (GPU box, NBVH_CHAIN_STATS=1 variant library): renders a few 1080p Chess frames with consecutive sample indices;
k_neural prints the histogram of (chain length predicted from the last frame's 3x3 neighbourhood, actual chain length) per launch."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import nbvh_b200  # noqa: E402

mod = nbvh_b200.load_chess_module(max_depth=1, envmap_eval=False)
mod.reinitialise_inference_data((1920, 1080))
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    print("[frame %d]" % i, flush=True)
    mod.neural_render(i, 0)
    torch.cuda.synchronize()
mod.close()
