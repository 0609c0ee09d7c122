#!/usr/bin/env python3
"""End-to-end quality of the online optimisation on the B200 path, measured the way the paper's comparisons are: the path-traced image of
the NEURAL scene against the path-traced image of the same scene with the neural mesh's real triangles (ignore_neural_blas = true).
Three images of the shipped Chess view (shipped max_depth 3, HDRI, real glTF materials / textures, `spp` samples each):
  ground truth       classical BLAS of the neural mesh
  shipped weights    the authors' pre-trained tree-cut + network (trained with the reference implementation)
  trained here       from-scratch optimisation with tree-cut learning through this library (shipped schedule, Adam, batch 2^18)
and LDR-FLIP (nbvh_flip_error) / relMSE (nbvh_image_error) of both neural images against the ground truth.
usage (GPU box): python tools/train_quality.py [steps] [spp] [W H]"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import nbvh_b200  # noqa: E402
from nbvh_b200.scene_io import CHESS_BASE  # noqa: E402
from nbvh_b200.treecut import BvhSplitScheduler, TreeCutLearner  # noqa: E402


def _trimmed_relmse(img, gt, keep=0.999):
    """The reference's relMSE (d^2 / (ref + 0.1), mean over pixels and channels) without the worst 0.1 % of the pixels (fireflies of
    glossy paths whose neural normal deviates slightly dominate the plain mean)."""
    e = ((img[..., :3] - gt[..., :3]) ** 2 / (gt[..., :3] + 0.1)).mean(-1).flatten()
    k = int(e.numel() * keep)
    return float(torch.sort(e).values[:k].mean())


def render(mod, accum, spp):
    for s in range(spp):
        mod.neural_render(s, 0)
    torch.cuda.synchronize()
    return (accum / float(spp)).contiguous()


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
    root = nbvh_b200.chess_fixture_dir()
    cfg = json.load(open(os.path.join(root, CHESS_BASE + "_neural.json")))
    cfg["cuda_scene"]["fb_size"] = [W, H]
    rend = cfg["neural_bvh_renderer"]
    mod = nbvh_b200.NeuralPathModule(json.dumps(cfg))
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root, with_neural_mesh=True, with_attributes=True)
    mats, textures, env = nbvh_b200.load_shading_arrays(root, with_envmap=True)
    mod.set_scene(tl, ti, bl, materials=mats, textures=textures)
    mod.set_envmap(env)
    mod.max_depth = rend["neural_module"].get("max_depth", 3)
    accum, _ = mod.attach_framebuffer((W, H))
    out = {"steps": steps, "spp": spp, "fb": [W, H], "max_depth": mod.max_depth}

    def set_ignore(flag):
        c = mod.get_constants(nbvh_b200.CONST_RENDERER)
        c.ignore_neural_blas = flag
        mod.set_constants(nbvh_b200.CONST_RENDERER, c)

    # ground truth: the neural mesh through its own triangles
    mod.load_neural_bvh(os.path.join(root, CHESS_BASE + ".nbvh"))
    mod.load_network_model(os.path.join(root, CHESS_BASE + "_network.json"))
    set_ignore(True)
    gt = render(mod, accum, spp)
    set_ignore(False)

    gt_ldr = gt.clamp(0.0, 1.0).contiguous()

    def score(tag, extra=None):
        img = render(mod, accum, spp)
        # LDR-FLIP is defined on [0, 1] images (HDR highlights make the feature exponent exceed 1 -- in flip-cuda as well as here --
        # and the error infinite), so both images are clamped first; relMSE is computed on the unclamped radiance
        flip, _ = mod.flip_images(gt_ldr, img.clamp(0.0, 1.0).contiguous())
        d = (img[..., :3] - gt[..., :3]).abs()
        out[tag] = {"flip_ldr": flip, "relmse": mod.image_error(gt, spp), "mean_abs_diff": float(d.mean()),
                    "relmse_without_top_0.1pct_pixels": _trimmed_relmse(img, gt),
                    "tree_cut_nodes": len(mod.neural_bvh_nodes())}
        if extra:
            out[tag].update(extra)

    score("shipped_weights")
    # from scratch through this library
    mod.init_params(1337)
    mod.tree_cut_init(rend.get("neural_bvh_min_depth", 0))
    sched = BvhSplitScheduler.from_config(cfg)
    sched.update_stats_and_reset(1, rend["max_training_steps"])
    learner = TreeCutLearner(mod, sched, batch_size=rend["batch_size"], num_lods_learned=rend["bvh_constants"]["num_lods_learned"])
    torch.cuda.synchronize()
    t0 = time.time()
    losses = []
    for s in range(steps):
        loss, _ = learner.step()
        losses.append(loss)
        if (s + 1) % 1000 == 0:
            print("step %d loss %.5g nodes %d" % (s + 1, float(np.mean(losses[-100:])), len(mod.neural_bvh_nodes())), flush=True)
    torch.cuda.synchronize()
    train_s = time.time() - t0
    score("trained_here", {"training_seconds": train_s, "ms_per_step": train_s * 1e3 / max(1, steps), "loss_first": losses[0],
                           "loss_last_mean_100": float(np.mean(losses[-100:]))})
    out["what"] = ("Chess, shipped view and settings; LDR-FLIP / relMSE of the neural path-traced image against the path-traced image of the "
                   "real triangles, for the authors' shipped weights and for a from-scratch optimisation through libnbvh_b200 on one B200")
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(REPO, "gpurun_out", "train_quality.json"), "w"), indent=1)
    mod.close()


if __name__ == "__main__":
    main()
