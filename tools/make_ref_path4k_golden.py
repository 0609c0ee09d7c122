#!/usr/bin/env python3
"""BASELINE config 4 at its OWN size from the reference's OWN hybrid path tracer (TEST INFRASTRUCTURE; run ON THE GPU BOX).

Drives oracle/_ref/ref_render_harness in `path` mode at 3840x2160, max_depth 4, 16 spp on the shipped Chess scene / config /
tree-cut / weights / HDRI / chessboard material -- the configuration bench.py times as config 4 -- and stores the image
8x8-box-filtered to 480x270 (each stored pixel averages 64 pixels x 16 samples of the reference) plus its mean radiance:
tests/golden/ref_chess_path_4k_16spp.npz (~1.5 MB).  Also renders the same image with nbvh_b200 and prints the metrics the test
asserts (relMSE, mean radiance, LDR-FLIP of the box-filtered images), so the stated bounds come from a measurement.

Outputs land in gpurun_out/ref_golden/; copy the .npz to tests/golden/ and commit."""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
OUT = os.path.join(REPO, "gpurun_out", "ref_golden")
W, H, DEPTH, SPP, BOX = 3840, 2160, 4, 16, 8


def box_filter(img):
    return img.reshape(H // BOX, BOX, W // BOX, BOX, 3).mean(axis=(1, 3)).astype(np.float32)


def main():
    import torch
    import nbvh_b200
    import make_ref_render_goldens as R
    import make_ref_path_goldens as P
    from nbvh_b200.scene_io import CHESS_BASE, chess_fixture_dir
    root = chess_fixture_dir()
    cfg = json.load(open(os.path.join(root, CHESS_BASE + "_neural.json")))
    log2_hashmap = int(CHESS_BASE.split("hash_")[1].split("_")[0])
    mod = nbvh_b200.load_chess_module(max_depth=DEPTH, envmap_eval=True)  # constants only: the harness renders
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root, with_attributes=True)
    os.makedirs(OUT, exist_ok=True)
    wd = os.path.join(OUT, "work_path_4k")
    R.prepare(mod, cfg, wd, W, H, (tl, ti, bl))
    P.stage_shading(root, wd, bl)
    mod.close()
    info = P.run(R.HARNESS, wd, W, H, DEPTH, log2_hashmap, 1, SPP)
    acc = np.fromfile(os.path.join(wd, "path_accum.f32"), dtype="<f4").reshape(H, W, 4)
    ref = box_filter(acc[..., :3] / SPP)
    np.savez_compressed(os.path.join(OUT, "ref_chess_path_4k_16spp.npz"), W=W, H=H, max_depth=DEPTH, spp=SPP, box=BOX, rgb_box=ref,
                        mean_radiance=np.float64(acc[..., :3].mean() / SPP), queries_1spp=info["queries"])
    for fn in os.listdir(wd):
        if not fn.endswith(".json"):
            os.remove(os.path.join(wd, fn))
    # ---- the same image with nbvh_b200 (one launch sequence for the 16 samples, as bench.py's config 4)
    pmod = nbvh_b200.load_chess_module(max_depth=DEPTH, envmap_eval=True, shading=True)
    accum, _ = pmod.attach_framebuffer((W, H))
    pmod.neural_render_samples(0, SPP)
    torch.cuda.synchronize()
    ours = box_filter(accum.cpu().numpy()[..., :3] / SPP)
    relmse = float((np.square(ref - ours) / (ref + 0.1)).mean())
    to4 = lambda a: torch.from_numpy(np.concatenate([a, np.ones(a.shape[:2] + (1,), np.float32)], axis=-1)).cuda().contiguous()
    flip, _ = pmod.flip_images(to4(ref), to4(ours))
    print(json.dumps({"relMSE_box_filtered": relmse, "mean_radiance_ref": float(ref.mean()), "mean_radiance_ours": float(ours.mean()),
                      "LDR_FLIP_box_filtered": float(flip), "max_abs": float(np.abs(ref - ours).max())}))
    pmod.close()


if __name__ == "__main__":
    main()
