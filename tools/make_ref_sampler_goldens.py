#!/usr/bin/env python3
"""Golden vectors from the reference's OWN training-sample generator (K10; TEST INFRASTRUCTURE, run ON THE GPU BOX).

Drives oracle/_ref/ref_render_harness in sampler mode: granular_sample_bvh_network_input_data
(/root/reference/src/cuda/neural/neural_bvh_module_kernels.cu:119-264) on the shipped Chess tree-cut and the neural
mesh's classical BLAS with its real normals / texcoords / material segments.  The glTF's JPEG textures are replaced by
deterministic procedural materials (nbvh_b200.procedural_materials) and the renderer-owned tree-cut statistics by
nbvh_b200.procedural_tree_cut, so the tests can rebuild the same inputs without an image decoder.

Writes gpurun_out/ref_golden/ref_chess_sampler.npz (+ ref_chess_sampler_timing.json); copy to tests/golden/ and commit."""
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tools.make_ref_render_goldens import HARNESS, OUT, prepare  # noqa: E402

MAX_LOSS = 1.0
CASES = [("t0", 4096, 0, 2), ("t2", 4096, 2, 2)]  # name, batch, log2_bvh_max_traversals, steps


def write_sampler_inputs(wd, blas, mats, texs, cut):
    for i, b in enumerate(blas):
        p = os.path.join(wd, "mesh%d" % i)
        for key, suffix, dt in (("normals", "_normals.f32", "<f4"), ("texcoords", "_texcoords.f32", "<f4"),
                                ("material_ids", "_material_ids.i32", "<i4"), ("material_map", "_material_map.u32", "<u4")):
            if b.get(key) is not None:
                np.ascontiguousarray(b[key], dtype=dt).tofile(p + suffix)
    mats.tofile(os.path.join(wd, "materials.bin"))
    for k, img in enumerate(texs):
        img.tofile(os.path.join(wd, "tex%d.rgba8" % k))
        np.array([img.shape[1], img.shape[0]], dtype="<u4").tofile(os.path.join(wd, "tex%d.dims" % k))
    cut.tofile(os.path.join(wd, "tree_cut.bin"))


def run_sampler(wd, log2_hashmap, batch, log2_trav, steps):
    out = subprocess.run([HARNESS, wd, "64", "64", "1", str(log2_hashmap), "0", "sampler", str(batch), str(log2_trav), str(steps), str(MAX_LOSS)],
                         capture_output=True, text=True)
    print(out.stdout[-800:], out.stderr[-3000:], flush=True)
    if out.returncode != 0:
        raise SystemExit("ref_render_harness (sampler) failed")
    f = lambda n, dt: np.fromfile(os.path.join(wd, n), dtype=dt)
    return (f("sampler_in.f32", "<f4").reshape(steps, batch, 9), f("sampler_tgt.f32", "<f4").reshape(steps, batch, 8),
            f("sampler_nodes.u32", "<u4").reshape(steps, batch), json.load(open(os.path.join(wd, "sampler.json"))))


def main():
    import nbvh_b200
    from nbvh_b200.scene_io import CHESS_BASE, chess_fixture_dir
    root = chess_fixture_dir()
    cfg = json.load(open(os.path.join(root, CHESS_BASE + "_neural.json")))
    log2_hashmap = int(CHESS_BASE.split("hash_")[1].split("_")[0])
    mod = nbvh_b200.load_chess_module(max_depth=1, envmap_eval=False)
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root, with_neural_mesh=True, with_attributes=True)
    n_mat = int(max(b["material_ids"].max() for b in bl)) + 1
    mats, texs = nbvh_b200.procedural_materials(n_mat)
    cut = nbvh_b200.procedural_tree_cut(len(mod.neural_bvh_nodes()))
    wd = os.path.join(OUT, "work_sampler")
    prepare(mod, cfg, wd, 64, 64, (tl, ti, bl))
    write_sampler_inputs(wd, bl, mats, texs, cut)
    gold = {}
    for name, batch, log2_trav, steps in CASES:
        x, t, n, _ = run_sampler(wd, log2_hashmap, batch, log2_trav, steps)
        gold.update({name + "_in": x.view(np.uint32), name + "_tgt": t.view(np.uint32), name + "_nodes": n,
                     name + "_cfg": np.array([batch, log2_trav, steps], dtype=np.uint32)})
        print("[sampler goldens] %s: hit rate %.3f, untouched slots %.3f" % (name, t[..., 0].mean(), (np.abs(x).sum(axis=-1) == 0).mean()))
    np.savez_compressed(os.path.join(OUT, "ref_chess_sampler.npz"), max_loss=MAX_LOSS, **gold)
    # timing at the shipped batch size (reported beside nbvh_b200's sampler; outputs discarded)
    _, _, _, info = run_sampler(wd, log2_hashmap, 1 << 18, 0, 6)
    info["what"] = "reference granular_sample_bvh_network_input_data, sm_100 build, Chess, batch 2^18, procedural materials"
    json.dump(info, open(os.path.join(OUT, "ref_chess_sampler_timing.json"), "w"), indent=1)
    for fn in os.listdir(wd):
        if not fn.endswith(".json"):
            os.remove(os.path.join(wd, fn))
    mod.close()


if __name__ == "__main__":
    main()
