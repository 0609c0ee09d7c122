#!/usr/bin/env python3
"""Golden of the reference's OWN wavefront path with SegmentDistribution::Uniform (TEST INFRASTRUCTURE, run ON THE GPU BOX).

Same harness, scene, tree-cut and weights as tools/make_ref_render_goldens.py (small frame), with
bvh_module_const.segment_distribution = Uniform (/root/reference/src/cuda/neural/neural_bvh_module_kernels.cu:53-69): every query draws
one interpolation parameter per segment point from its ray's generator (encode_bvh_network_input_data :292-297).  The shipped weights
were trained with equidistant points, so the frame is not meaningful as an image -- as a parity vector it is as good as any: the first
wavefront step's network inputs pin the generator state and the draw order bit for bit, the hit records the rest of the chain.

Writes gpurun_out/ref_golden/ref_chess_uniform_256x144.npz; copy to tests/golden/ and commit."""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tools.make_ref_render_goldens import OUT, load_outputs, prepare, run  # noqa: E402


def main():
    import nbvh_b200
    from nbvh_b200 import _capi as A
    from nbvh_b200.scene_io import CHESS_BASE, chess_fixture_dir
    root = chess_fixture_dir()
    cfg = json.load(open(os.path.join(root, CHESS_BASE + "_neural.json")))
    log2_hashmap = int(CHESS_BASE.split("hash_")[1].split("_")[0])
    mod = nbvh_b200.load_chess_module(max_depth=1, envmap_eval=False)
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root)
    params = mod.get_params().copy()
    mc = mod.get_constants(A.CONST_MODULE)
    mc.segment_distribution = 2
    mod.set_constants(A.CONST_MODULE, mc)
    mod.load_params(params)  # (a module block that leaves the layout unchanged keeps the parameters; reloading is harmless)
    W, H = 256, 144
    wd = os.path.join(OUT, "work_uniform")
    prepare(mod, cfg, wd, W, H, (tl, ti, bl))
    run(wd, W, H, 1, log2_hashmap, 1)
    o = load_outputs(wd, W * H)
    order = np.argsort(o["s0_px"], kind="stable")
    np.savez_compressed(os.path.join(OUT, "ref_chess_uniform_256x144.npz"), W=W, H=H, n=W * H,
                        hit_kind=o["hits"]["valid_hit"].astype(np.uint8), hit_t=o["hits"]["t"].view(np.uint32),
                        s0_px=o["s0_px"][order], s0_in=o["s0_in"][order].view(np.uint32), s0_out=o["s0_out"][order].astype(np.float16))
    print("[uniform golden] first-step visibility > 0.5: %d of %d; NaN outputs %d" %
          ((o["s0_out"][:, 0] > 0.5).sum(), len(order), np.isnan(o["s0_out"]).sum()))
    k = o["hits"]["valid_hit"]
    print("[uniform golden] kinds: miss %d neural %d classical %d; %d first-step queries" %
          ((k == 0).sum(), (k == 1).sum(), (k == 2).sum(), len(order)))
    for fn in os.listdir(wd):
        if not fn.endswith(".json"):
            os.remove(os.path.join(wd, fn))
    mod.close()


if __name__ == "__main__":
    main()
