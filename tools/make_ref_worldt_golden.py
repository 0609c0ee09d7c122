#!/usr/bin/env python3
"""Golden hit records of the reference's OWN wavefront path with DepthLearningMode::WorldT (TEST INFRASTRUCTURE, run ON THE GPU BOX).

Same harness, scene, tree-cut and weights as tools/make_ref_render_goldens.py (small frame), with
pt_target_const.depth_learning_mode = WorldT (/root/reference/include/cuda/neural/neural_bvh_module_common.cuh:58-88): the depth
output is decoded against the ray's interval of the scene box instead of the leaf's.  The shipped weights were trained for LocalT,
so the decoded distances are not meaningful as geometry -- as a parity vector they are as good as any.

Writes gpurun_out/ref_golden/ref_chess_worldt_256x144.npz; copy to tests/golden/ and commit."""
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
    pt = mod.get_constants(A.CONST_PT_OUTPUT)
    pt.depth_learning_mode = 1
    mod.set_constants(A.CONST_PT_OUTPUT, pt)
    W, H = 256, 144
    wd = os.path.join(OUT, "work_worldt")
    prepare(mod, cfg, wd, W, H, (tl, ti, bl))
    run(wd, W, H, 1, log2_hashmap, 1)
    o = load_outputs(wd, W * H)
    np.savez_compressed(os.path.join(OUT, "ref_chess_worldt_256x144.npz"), W=W, H=H, n=W * H,
                        hit_kind=o["hits"]["valid_hit"].astype(np.uint8), hit_t=o["hits"]["t"].view(np.uint32))
    k = o["hits"]["valid_hit"]
    print("[worldt golden] kinds: miss %d neural %d classical %d" % ((k == 0).sum(), (k == 1).sum(), (k == 2).sum()))
    for fn in os.listdir(wd):
        if not fn.endswith(".json"):
            os.remove(os.path.join(wd, fn))
    mod.close()


if __name__ == "__main__":
    main()
