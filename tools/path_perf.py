#!/usr/bin/env python3
"""Timing of the multi-bounce path pipeline on the shipped Chess scene (GPU box; CUDA events, warm frames):
BASELINE config 2b (1920x1080, shipped max_depth 3, 1 spp) and config 4's per-GPU unit (3840x2160, max_depth 4, 1 spp)."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import nbvh_b200  # noqa: E402


def main():
    out = {}
    for name, (W, H, depth) in {"config2a_1080p_depth1": (1920, 1080, 1), "config2b_1080p_depth3": (1920, 1080, 3),
                                "config4_2160p_depth4": (3840, 2160, 4)}.items():
        mod = nbvh_b200.load_chess_module(max_depth=depth, envmap_eval=True, shading=True)
        mod.attach_framebuffer((W, H))
        for s in range(3):
            mod.neural_render(s, 0)
        torch.cuda.synchronize()
        mod.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for s in range(iters):
            mod.neural_render(s, 0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        c = mod.counters()
        rays = c["rays"] / iters
        out[name] = {"ms_per_frame": ms, "pixels": W * H, "rays_per_frame": rays, "queries_per_frame": c["queries"] / iters,
                     "Mrays_s_all_rays": rays / ms / 1e3, "Mpixels_s": W * H / ms / 1e3}
        print(name, json.dumps(out[name]), flush=True)
        mod.close()
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(REPO, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "path_perf.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
