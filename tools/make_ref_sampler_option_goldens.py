#!/usr/bin/env python3
"""Golden vectors from the reference's OWN training-sample generator for the sampler options no shipped config uses
(TEST INFRASTRUCTURE, run ON THE GPU BOX).

Same harness and inputs as tools/make_ref_sampler_goldens.py (granular_sample_bvh_network_input_data,
/root/reference/src/cuda/neural/neural_bvh_module_kernels.cu:119-264, on the shipped Chess tree-cut and the neural mesh's BLAS
with procedural materials), with the constant blocks changed per case:

  uniform        SegmentDistribution::Uniform (neural_bvh_module_kernels.cu:53-69)
  uniform_t2     the same with 4 slots per thread (rays continuing through the tree-cut)
  worldt         DepthLearningMode::WorldT          (neural_bvh_module_common.cuh:599-606)
  position       DepthLearningMode::Position        (:607-613; 3 depth outputs, albedo dropped -> 7 outputs)
  localposition  DepthLearningMode::LocalPosition   (:614-622)
  viewdep        view_dependent_sampling, ratio 0.5 (:734-737)
  leafboundary   leaf_boundary_sampling_ratio 0.5   (:765-777)
  bounce         max_training_bounces 2, cosine-weighted bounce, 4 slots per thread (neural_bvh_module_kernels.cu:227-262)
  bounce_sphere  max_training_bounces 3, uniform-sphere bounce, back-face bounces allowed

Writes gpurun_out/ref_golden/ref_chess_sampler_options.npz; copy to tests/golden/ and commit."""
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tools.make_ref_render_goldens import HARNESS, OUT, prepare  # noqa: E402
from tools.make_ref_sampler_goldens import MAX_LOSS, write_sampler_inputs  # noqa: E402

BATCH, STEPS = 2048, 2
# name -> (log2_bvh_max_traversals, {block: {field: value}})
CASES = {
    "uniform": (0, {"module": {"segment_distribution": 2}}),
    "uniform_t2": (2, {"module": {"segment_distribution": 2}}),
    "worldt": (0, {"pt_output": {"depth_learning_mode": 1}}),
    "position": (0, {"pt_output": {"depth_learning_mode": 2, "output_depth_dim": 3, "output_albedo_dim": 0, "output_network_dims": 7}}),
    "localposition": (0, {"pt_output": {"depth_learning_mode": 3, "output_depth_dim": 3, "output_albedo_dim": 0, "output_network_dims": 7}}),
    "viewdep": (0, {"renderer": {"view_dependent_sampling": True, "view_dependent_ratio": 0.5}}),
    "leafboundary": (0, {"renderer": {"leaf_boundary_sampling_ratio": 0.5}}),
    "bounce": (2, {"renderer": {"max_training_bounces": 2, "sample_cosine_weighted": True}}),
    "bounce_sphere": (2, {"renderer": {"max_training_bounces": 3, "sample_cosine_weighted": False, "backface_bounce_training": True}}),
}


def apply_case(mod, changes):
    """Sets the case's fields on the module (the library validates them) and returns the blocks to restore afterwards."""
    from nbvh_b200 import _capi as A
    which = {"module": A.CONST_MODULE, "pt_output": A.CONST_PT_OUTPUT, "renderer": A.CONST_RENDERER}
    saved = []
    for block, fields in changes.items():
        pod = mod.get_constants(which[block])
        saved.append((which[block], bytes(pod)))
        for k, v in fields.items():
            setattr(pod, k, v)
        mod.set_constants(which[block], pod)
    return saved


def restore(mod, saved):
    from nbvh_b200 import _capi as A
    for which, raw in reversed(saved):
        pod = A.CONST_TYPES[which].from_buffer_copy(raw)
        mod.set_constants(which, pod)


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
    gold = {}
    for name, (log2_trav, changes) in CASES.items():
        saved = apply_case(mod, changes)
        wd = os.path.join(OUT, "work_sampler_" + name)
        prepare(mod, cfg, wd, 64, 64, (tl, ti, bl))
        write_sampler_inputs(wd, bl, mats, texs, cut)
        out = subprocess.run([HARNESS, wd, "64", "64", "1", str(log2_hashmap), "0", "sampler", str(BATCH), str(log2_trav), str(STEPS), str(MAX_LOSS)],
                             capture_output=True, text=True)
        print(name, out.stdout[-300:], out.stderr[-2000:], flush=True)
        if out.returncode != 0:
            raise SystemExit("ref_render_harness (sampler, %s) failed" % name)
        n_out = int(mod.get_constants(nbvh_b200._capi.CONST_PT_OUTPUT).output_network_dims)
        f = lambda n, dt: np.fromfile(os.path.join(wd, n), dtype=dt)
        x, t, n = f("sampler_in.f32", "<u4").reshape(STEPS, BATCH, 9), f("sampler_tgt.f32", "<u4").reshape(STEPS, BATCH, n_out), f("sampler_nodes.u32", "<u4").reshape(STEPS, BATCH)
        gold.update({name + "_in": x, name + "_tgt": t, name + "_nodes": n, name + "_cfg": np.array([BATCH, log2_trav, STEPS], dtype=np.uint32)})
        print("[sampler option goldens] %s: hit rate %.3f, untouched slots %.3f" % (name, t.view(np.float32)[..., 0].mean(), (x.sum(axis=-1) == 0).mean()), flush=True)
        restore(mod, saved)
        for fn in os.listdir(wd):
            os.remove(os.path.join(wd, fn))
        os.rmdir(wd)
    np.savez_compressed(os.path.join(OUT, "ref_chess_sampler_options.npz"), max_loss=MAX_LOSS, **gold)
    mod.close()


if __name__ == "__main__":
    main()
