#!/usr/bin/env python3
"""Development check (GPU box): nbvh_b200's sample generator against the reference kernel's goldens, with statistics."""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def chess_sampler_module():
    import nbvh_b200
    from nbvh_b200.scene_io import chess_fixture_dir
    root = chess_fixture_dir()
    mod = nbvh_b200.load_chess_module(max_depth=1, envmap_eval=False)
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root, with_neural_mesh=True, with_attributes=True)
    n_mat = int(max(b["material_ids"].max() for b in bl)) + 1
    mats, texs = nbvh_b200.procedural_materials(n_mat)
    mod.set_scene(tl, ti, bl, materials=mats, textures=texs)
    cut = nbvh_b200.procedural_tree_cut(len(mod.neural_bvh_nodes()))
    return mod, cut


def run_case(mod, cut_dev, batch, log2_trav, steps, max_loss):
    import torch
    nodes = torch.zeros(batch, dtype=torch.int32, device="cuda")
    x = torch.zeros((batch, 9), dtype=torch.float32, device="cuda")
    t = torch.zeros((batch, 8), dtype=torch.float32, device="cuda")
    xs, ts, ns = [], [], []
    for s in range(steps):
        mod.optimisation_data_generation(s, batch, log2_trav, nodes, max_loss, cut_dev, x, t)
        torch.cuda.synchronize()
        xs.append(x.cpu().numpy().copy()), ts.append(t.cpu().numpy().copy()), ns.append(nodes.cpu().numpy().astype(np.uint32))
    return np.stack(xs), np.stack(ts), np.stack(ns)


def compare(name, x, t, n, gx, gt, gn):
    same_node = n == gn
    xe = np.abs(x - gx).max(axis=-1)
    te = np.abs(t - gt).max(axis=-1)
    bit_x = (x.view(np.uint32) == gx.view(np.uint32)).all(axis=-1)
    bit_t = (t.view(np.uint32) == gt.view(np.uint32)).all(axis=-1)
    print("[%s] slots %d: node ids equal %.5f | inputs bit-equal %.5f, max|dx| %.3g (99.9%%: %.3g) | targets bit-equal %.5f, hit flag equal %.5f, "
          "max|dt| %.3g (99.9%%: %.3g)" % (name, n.size, same_node.mean(), bit_x.mean(), xe.max(), np.quantile(xe, 0.999), bit_t.mean(),
                                          (t[..., 0] == gt[..., 0]).mean(), te.max(), np.quantile(te, 0.999)))
    for c, lbl in ((1, "depth"), (2, "normal.x"), (5, "albedo.r")):
        e = np.abs(t[..., c] - gt[..., c])[same_node & (t[..., 0] == gt[..., 0])]
        print("    %-9s max %.3g  99.9%% %.3g  bit-equal %.5f" % (lbl, e.max(), np.quantile(e, 0.999),
                                                              (t[..., c].view(np.uint32) == gt[..., c].view(np.uint32)).mean()))
    return same_node.mean()


def main():
    import torch
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "tests", "golden", "ref_chess_sampler.npz")
    g = np.load(path)
    mod, cut = chess_sampler_module()
    cut_dev = torch.from_numpy(cut.view(np.uint8).reshape(-1).copy()).cuda()
    for name in ("t0", "t2"):
        batch, log2_trav, steps = [int(v) for v in g[name + "_cfg"]]
        x, t, n = run_case(mod, cut_dev, batch, log2_trav, steps, float(g["max_loss"]))
        compare(name, x, t, n, g[name + "_in"].view(np.float32), g[name + "_tgt"].view(np.float32), g[name + "_nodes"])
    # throughput at the shipped batch size
    B = 1 << 18
    nodes = torch.zeros(B, dtype=torch.int32, device="cuda")
    x = torch.zeros((B, 9), dtype=torch.float32, device="cuda")
    t = torch.zeros((B, 8), dtype=torch.float32, device="cuda")
    for s in range(3):
        mod.optimisation_data_generation(s, B, 0, nodes, 1.0, cut_dev, x, t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(10):
        mod.optimisation_data_generation(3 + s, B, 0, nodes, 1.0, cut_dev, x, t)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"sampler_ms_per_step_2p18": e0.elapsed_time(e1) / 10}))
    mod.close()


if __name__ == "__main__":
    main()
