#!/usr/bin/env python3
"""Data-parallel optimisation WITH tree-cut learning, one process per GPU (run under torchrun on >= 2 GPUs):
gradients and per-node tree-cut statistics are allreduced over NCCL, every rank runs the deterministic device-side split.
Checks that weights and tree-cut stay bit-identical on all ranks and that the loss falls.  Prints one JSON line on rank 0."""
import hashlib
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import nbvh_b200  # noqa: E402
from nbvh_b200.treecut import BvhSplitScheduler, TreeCutLearner  # noqa: E402


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    cfg, nbvh, tlas, tlas_idx, blas = nbvh_b200.synthetic_scene()
    mod = nbvh_b200.NeuralPathModule(cfg, device=local)
    mod.set_neural_bvh(nbvh)
    mod.set_scene(tlas, tlas_idx, blas)
    mod.init_params(1337)
    mod.tree_cut_init(0)
    mode = os.environ.get("NBVH_DP", "lib")
    if world > 1 and mode == "lib":  # the library's own NCCL communicator (comm.cu); "torch": the host's process group sums
        mod.setup_data_parallel(rank, world)
    sched = BvhSplitScheduler(splitting_interval=10, last_split_iteration=100, start_split_num=1, split_scaling=2.0, split_interval_scaling=2.0)
    sched.update_stats_and_reset(1, steps)
    learner = TreeCutLearner(mod, sched, batch_size=(1 << 15) // world, rank=rank, world=world)
    losses = []
    for _ in range(steps):
        loss, _ = learner.step()
        losses.append(loss)
    torch.cuda.synchronize()
    nodes, cut, nmap = mod.tree_cut_state()
    h = hashlib.sha256(nodes.tobytes() + nmap.tobytes() + cut["samples"].tobytes() + cut["loss"].tobytes() + mod.get_params().tobytes()).digest()
    ht = torch.tensor(list(h), dtype=torch.uint8, device="cuda")
    same = True
    if world > 1:
        gathered = [torch.empty_like(ht) for _ in range(world)]
        dist.all_gather(gathered, ht)
        same = all(bool((g == gathered[0]).all()) for g in gathered)
    if rank == 0:
        print(json.dumps({"world": world, "steps": steps, "gradient_exchange": mode if world > 1 else None, "identical_tree_cut_and_weights_on_all_ranks": same, "tree_cut_nodes": len(nodes),
                          "loss_first10": float(np.mean(losses[:10])), "loss_last10": float(np.mean(losses[-10:]))}))
    ok = same and np.mean(losses[-10:]) < np.mean(losses[:10]) and len(nodes) > 1
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    mod.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
