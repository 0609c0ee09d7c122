#!/usr/bin/env python3
"""Validation-mode run on the Chess fixture (ValidationMode::Full of the reference, headless): from-scratch optimisation with tree-cut
learning, then `spp` frames, leaving results/ nbvhs/ network_models/ configs/ like the reference does (nbvh_b200/validation.py).
Then reloads what it wrote into a fresh module and checks that the first frame is reproduced bit for bit.
usage (GPU box): python tools/validation_run.py [steps] [spp] [W H] [out_dir]"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import nbvh_b200  # noqa: E402
from nbvh_b200 import validation  # noqa: E402
from nbvh_b200.scene_io import CHESS_BASE  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
    out = sys.argv[5] if len(sys.argv) > 5 else "/tmp/nbvh_validation"   # ~70 MB of EXR + weights: not under gpurun_out/
    root = nbvh_b200.chess_fixture_dir()
    cfg = json.load(open(os.path.join(root, CHESS_BASE + "_neural.json")))
    cfg["cuda_scene"]["fb_size"] = [W, H]
    mod = nbvh_b200.NeuralPathModule(json.dumps(cfg))
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root, with_neural_mesh=True, with_attributes=True)
    mats, texs = nbvh_b200.procedural_materials(int(max(b["material_ids"].max() for b in bl)) + 1)
    mod.set_scene(tl, ti, bl, materials=mats, textures=texs)
    mod.max_depth = 1
    accum, _ = mod.attach_framebuffer((W, H))
    res = validation.run_full_validation(mod, cfg, out, max_training_steps=steps, max_accumulated_spp=spp, base=CHESS_BASE, variant="b200",
                                         log=lambda s: print(s, flush=True))
    first = None
    mod.neural_render(0, 0)
    torch.cuda.synchronize()
    first = accum.cpu().numpy().copy()
    mod.close()
    # reload: saved config + saved tree-cut + saved weights -> the same frame
    mod2 = nbvh_b200.NeuralPathModule(open(res["paths"]["config"]).read())
    mod2.load_neural_bvh(res["paths"]["nbvh"])
    mod2.load_network_model(res["paths"]["network"])
    mod2.set_scene(tl, ti, bl, materials=mats, textures=texs)
    accum2, _ = mod2.attach_framebuffer((W, H))
    mod2.neural_render(0, 0)
    torch.cuda.synchronize()
    same = bool((accum2.cpu().numpy().view(np.uint32) == first.view(np.uint32)).all())
    mod2.close()
    summary = {"steps": res["steps"], "spp": res["spp"], "fb": [W, H], "tree_cut_nodes": res["tree_cut_nodes"],
               "loss_first": res["losses"][0], "loss_last_mean_50": float(np.mean(res["losses"][-50:])),
               "learning_ms_per_step_median": float(np.median(res["learning_ms"])),
               "inference_ms_per_frame_median": float(np.median(res["inference_ms"])),
               "reload_reproduces_first_frame_bitwise": same, "files": {k: os.path.relpath(v, out) for k, v in res["paths"].items()}}
    json.dump(summary, open(os.path.join(out, "summary.json"), "w"), indent=1)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    json.dump(summary, open(os.path.join(REPO, "gpurun_out", "validation_summary.json"), "w"), indent=1)
    print(json.dumps(summary))
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
