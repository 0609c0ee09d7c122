#!/usr/bin/env python3
"""compute-sanitizer target (GPU box): small synthetic frames through every consumer of the drain-order class lists --
primary pipeline (cold and warm history), multi-sample launch, interleaved shards, arrival order."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import nbvh_b200  # noqa: E402

cfg, nbvh, tlas, tlas_idx, blas = nbvh_b200.synthetic_scene(fb=(128, 96))
mod = nbvh_b200.NeuralPathModule(cfg)
mod.set_neural_bvh(nbvh)
mod.set_scene(tlas, tlas_idx, blas)
mod.init_params(1337)
mod.attach_framebuffer((128, 96))
for depth in (1,):
    mod.max_depth = depth
    for s in range(3):
        mod.neural_render(s, 0)
    mod.neural_render_samples(3, 4)
    for shard in range(3):
        mod.neural_render_shard(7, 0, shard, 3)
    torch.cuda.synchronize()
mod.set_queue_order((0, 0, 0, 0), 0)
mod.neural_render(8, 0)
torch.cuda.synchronize()
print("memcheck target done", mod.counters()["queued_rays"])
mod.close()
