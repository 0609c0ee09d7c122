#!/usr/bin/env python3
"""Development timing harness (GPU box): network queries, the training step, and the render kernel, CUDA-event timed."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import nbvh_b200  # noqa: E402


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    out = {}
    root = nbvh_b200.chess_fixture_dir()
    mod = nbvh_b200.load_chess_module() if root else None
    if mod is None:
        cfg, nbvh, tlas, tlas_idx, blas = nbvh_b200.synthetic_scene(fb=(1920, 1080))
        mod = nbvh_b200.NeuralPathModule(cfg)
        mod.set_neural_bvh(nbvh)
        mod.set_scene(tlas, tlas_idx, blas)
        mod.init_params(1)
    # the frame first: the training steps below change the weights (and with them the number of queries per frame)
    ms = timeit(lambda: mod.neural_render(0, 0), iters=20)
    out["render_1080p"] = {"ms": ms, "Mrays_s": 1920 * 1080 / ms / 1e3}
    # validation metrics on the 1080p frame (mean read back per call, like the reference's validation mode)
    accum, _ = mod.attach_framebuffer((1920, 1080))
    mod.neural_render(0, 0)
    ref_img = (torch.roll(accum, 2, dims=1) * 0.9 + 0.01).contiguous()
    ms = timeit(lambda: mod.flip_error(ref_img, 1), iters=20)
    out["flip_1080p"] = {"ms": ms, "Mpixels_s": 1920 * 1080 / ms / 1e3, "what": "nbvh_flip_error incl. the mean read-back"}
    ms = timeit(lambda: mod.image_error(ref_img, 1), iters=20)
    out["relmse_1080p"] = {"ms": ms}
    rng = np.random.default_rng(0)
    for n in (1 << 16, 1 << 20, 1 << 22):
        x = torch.from_numpy(rng.uniform(0, 1, size=(n, 9)).astype(np.float32)).cuda()
        y = torch.empty((n, 8), dtype=torch.float32, device="cuda")
        ms = timeit(lambda: mod.inference(x, y))
        out["query_%d" % n] = {"ms": ms, "Mqueries_s": n / ms / 1e3}
    B = 1 << 18
    x = torch.from_numpy(rng.uniform(0, 1, size=(B, 9)).astype(np.float32)).cuda()
    tg = rng.uniform(0, 1, size=(B, 8)).astype(np.float32)
    vis = rng.uniform(0, 1, size=B) < 0.5
    tg[:, 0] = vis
    tg[~vis, 1:] = 0
    t = torch.from_numpy(tg).cuda()
    ms = timeit(lambda: mod.training_step(x, t, want_loss=False), iters=30)
    out["train_step_2p18"] = {"ms": ms, "Msamples_s": B / ms / 1e3}
    ms = timeit(lambda: mod.training_forward_backward(x, t), iters=30)
    out["train_fwd_bwd_2p18"] = {"ms": ms}
    ms = timeit(lambda: mod.optimizer_step(), iters=30)
    out["adam_step"] = {"ms": ms}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
