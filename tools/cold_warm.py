#!/usr/bin/env python3
"""Timing-method check for the 1080p Chess primary frame: per-frame event pairs vs one event pair around N frames."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import nbvh_b200
mod = nbvh_b200.load_chess_module(max_depth=1, envmap_eval=False)
st = torch.cuda.current_stream()
n = 50
for rep in range(2):
    for vary in (0, 1):
        for i in range(3):
            mod.neural_render(i if vary else 0, 0)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for i in range(n):
            ev[i][0].record(st)
            mod.neural_render((3 + i) if vary else 0, 0)
            ev[i][1].record(st)
        torch.cuda.synchronize()
        per_frame = sum(a.elapsed_time(b) for a, b in ev) / n
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.time()
        e0.record(st)
        for i in range(n):
            mod.neural_render((3 + i) if vary else 0, 0)
        e1.record(st)
        t_submit = time.time() - t0
        torch.cuda.synchronize()
        print("vary_sample=%d : per-frame events %.4f ms | one pair around %d frames %.4f ms | CPU submit %.4f ms/frame" %
              (vary, per_frame, n, e0.elapsed_time(e1) / n, t_submit * 1e3 / n), flush=True)
mod.reset_counters()
for i in range(10):
    mod.neural_render(0, 0)
torch.cuda.synchronize()
print({k: v / 10 for k, v in mod.counters().items()})
mod.reset_counters()
for i in range(10):
    mod.neural_render(5 + i, 0)
torch.cuda.synchronize()
print({k: v / 10 for k, v in mod.counters().items()})
