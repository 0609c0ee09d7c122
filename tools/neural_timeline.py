#!/usr/bin/env python3
"""Timeline of ONE k_neural launch (1080p Chess primary frame): when the queue runs dry, when the last CTA ends, mean CTA lifetime."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import nbvh_b200
M = (1 << 64) - 1
for W, H, depth in ((1920, 1080, 1), (3840, 2160, 1)):
    mod = nbvh_b200.load_chess_module(max_depth=depth, envmap_eval=False)
    mod.reinitialise_inference_data((W, H))
    for i in range(3):
        mod.neural_render(i, 0)
    torch.cuda.synchronize()
    for rep in range(3):
        mod.reset_counters()
        mod.neural_render(5 + rep, 0)
        torch.cuda.synchronize()
        c = mod.counters()
        t0, tx, t1 = M - c["t_start_inv"], M - c["t_exhaust_inv"], c["t_end"]
        n_cta = 296
        print("%dx%d: k_neural span %.1f us | queue dry at %.1f us (%.0f %%) | mean CTA lifetime %.1f us | tile iterations %d, queries %d, fill %.2f" %
              (W, H, (t1 - t0) / 1e3, (tx - t0) / 1e3, 100.0 * (tx - t0) / (t1 - t0), c["t_cta_sum"] / n_cta / 1e3, c["tile_iterations"], c["queries"],
               c["queries"] / (128.0 * c["tile_iterations"])), flush=True)
    mod.close()
