#!/usr/bin/env python3
"""Optimisation step with the sample generation of step k + 1 on a second stream, overlapping the forward / backward / Adam of step k
(double-buffered inputs / targets; the samples of a step depend on the tree-cut and the RNG only, so the pipelined loop produces the
same batches as the sequential one).  Prints ms per step for both loops.  usage: train_overlap_probe.py [log2_batch]"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch

import nbvh_b200


def main():
    lb = int(sys.argv[1]) if len(sys.argv) > 1 else 18
    B = 1 << lb
    dev = torch.device("cuda", 0)
    root = nbvh_b200.chess_fixture_dir()
    mod = nbvh_b200.load_chess_module(max_depth=1, envmap_eval=False)
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root, with_neural_mesh=True, with_attributes=True)
    mats, texs = nbvh_b200.procedural_materials(int(max(b["material_ids"].max() for b in bl)) + 1)
    mod.set_scene(tl, ti, bl, materials=mats, textures=texs)
    cut = nbvh_b200.procedural_tree_cut(len(mod.neural_bvh_nodes()), seed=11)
    cut_dev = torch.from_numpy(cut.view(np.uint8).reshape(-1).copy()).to(dev)
    in_dims = int(mod.get_constants(nbvh_b200.CONST_MODULE).input_network_dims)
    out_dims = int(mod.network_output_dims())
    bufs = [(torch.zeros(B, dtype=torch.int32, device=dev), torch.zeros((B, in_dims), device=dev), torch.zeros((B, out_dims), device=dev))
            for _ in range(2)]
    walk_ahead = 16
    mod.sampler_walk_ahead(walk_ahead, 1)
    s_train = torch.cuda.current_stream()
    s_gen = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [torch.cuda.Event(), torch.cuda.Event()]
    step = [0]

    def sequential(n):
        sn, sx, st = bufs[0]
        for _ in range(n):
            mod.optimisation_data_generation(step[0], B, 0, sn, 1.0, cut_dev, sx, st, stream=s_train.cuda_stream)
            step[0] += 1
            mod.training_forward_backward(sx, st, stream=s_train.cuda_stream)
            mod.optimizer_step(stream=s_train.cuda_stream)

    def pipelined(n):
        # generation of step k on s_gen into buffer k & 1, once the training of step k - 2 has released it
        def gen(k):
            b = k & 1
            sn, sx, st = bufs[b]
            s_gen.wait_event(free[b])
            mod.optimisation_data_generation(step[0] + k, B, 0, sn, 1.0, cut_dev, sx, st, stream=s_gen.cuda_stream)
            ready[b].record(s_gen)
        for b in range(2):
            free[b].record(s_train)
        gen(0)
        for k in range(n):
            if k + 1 < n:
                gen(k + 1)
            b = k & 1
            sn, sx, st = bufs[b]
            s_train.wait_event(ready[b])
            mod.training_forward_backward(sx, st, stream=s_train.cuda_stream)
            mod.optimizer_step(stream=s_train.cuda_stream)
            free[b].record(s_train)
        step[0] += n

    def timeit(fn, n):
        fn(walk_ahead)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s_train)
        fn(n)
        s_train.wait_stream(s_gen)
        e1.record(s_train)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    out = {"log2_batch": lb}
    for _ in range(2):
        out["sequential_ms"] = timeit(sequential, 48)
        out["pipelined_ms"] = timeit(pipelined, 48)
    out["Msamples_s_sequential"] = B / out["sequential_ms"] / 1e3
    out["Msamples_s_pipelined"] = B / out["pipelined_ms"] / 1e3
    print(json.dumps(out))
    mod.close()


if __name__ == "__main__":
    main()
