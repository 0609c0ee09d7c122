#!/usr/bin/env python3
"""e2e throughput of the headline frame (1080p Chess, primary neural queries, host buffers) against the number of frames in flight
(nbvh_render_host_submit / _wait, slots 1 .. NBVH_MAX_FRAMES_IN_FLIGHT).  usage (GPU box): python tools/frames_in_flight.py [frames]"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import nbvh_b200  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    mod = nbvh_b200.load_chess_module()
    n = 1920 * 1080
    mod.reinitialise_inference_data((1920, 1080))
    cam = mod.get_constants(nbvh_b200.CONST_CAMERA)
    bufs = [torch.empty(n * 4, dtype=torch.uint8, pin_memory=True).numpy() for _ in range(4)]
    out = {}
    first = None
    for F in (1, 2, 3, 4):
        for rep in range(2):  # first repetition warms up
            t0 = time.time()
            for i in range(frames):
                s = i % F
                if i >= F:
                    mod.neural_render_host_wait(s)
                mod.neural_render_host_submit(0, i, bufs[s], s, camera=cam)
            for s in range(min(F, frames)):
                mod.neural_render_host_wait(s)
            torch.cuda.synchronize()
            dt = time.time() - t0
        out["frames_in_flight_%d" % F] = {"ms_per_frame": dt * 1e3 / frames, "Mrays_s": n * frames / dt / 1e6}
        # the last frame (RNG offset frames - 1) must not depend on the pipelining depth
        last = bufs[(frames - 1) % F].copy()
        if first is None:
            first = last
        out["frames_in_flight_%d" % F]["last_frame_identical_to_F1"] = bool((last == first).all())
    print(json.dumps(out, indent=1))
    json.dump(out, open(os.path.join(REPO, "gpurun_out", "frames_in_flight.json"), "w"), indent=1)
    mod.close()


if __name__ == "__main__":
    main()
