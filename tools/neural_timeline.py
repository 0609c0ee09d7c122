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

# multi-sample launches of the primary pipeline: the drain is paid once per launch, not once per sample
mod = nbvh_b200.load_chess_module(max_depth=1, envmap_eval=False)
mod.reinitialise_inference_data((1920, 1080))
for n_spp in (1, 2, 4, 8):
    for _ in range(2):
        mod.neural_render_samples(0, n_spp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for r in range(reps):
        mod.neural_render_samples(r * n_spp, n_spp)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("1080p primary, %d sample(s) per launch pair: %.3f ms per launch, %.3f ms per sample-frame, %.0f Mrays/s" %
          (n_spp, ms, ms / n_spp, 1920 * 1080 * n_spp / ms / 1e3), flush=True)
