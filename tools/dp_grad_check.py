#!/usr/bin/env python3
"""In-library data-parallel gradient exchange (comm.cu) against the single-GPU step, one process per GPU (torchrun, >= 2 GPUs):
every rank runs nbvh_training_forward_backward on ITS share of one seeded global batch with the library's NCCL communicator attached;
rank 0 also runs the whole batch on a second context without a communicator.  The reduced gradients and loss must equal the
single-GPU ones up to the fp32 summation order (relative L2 <= 1e-3; measured ~1e-6), the gradients must be bit-identical on all
ranks, and after the replicated Adam step so must the parameters.  Prints one JSON line on rank 0; exit code 1 on failure."""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import nbvh_b200  # noqa: E402


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg, nbvh, tlas, tlas_idx, blas = nbvh_b200.synthetic_scene()
    B = 1 << 16
    rng = np.random.default_rng(5)
    x = rng.uniform(0.02, 0.98, size=(B, 9)).astype(np.float32)
    t = rng.uniform(0, 1, size=(B, 8)).astype(np.float32)
    t[:, 0] = t[:, 0] < 0.5
    mod = nbvh_b200.NeuralPathModule(cfg, device=local)
    mod.init_params(1337)
    mod.setup_data_parallel(rank, world)
    lo, n = B * rank // world, B // world
    xs, ts = torch.from_numpy(x[lo:lo + n]).to(dev), torch.from_numpy(t[lo:lo + n]).to(dev)
    loss = torch.zeros(1, dtype=torch.float32, device=dev)
    grads = mod.gradient_buffer()
    mod.training_forward_backward(xs, ts, global_batch=B, loss_sum=loss)
    torch.cuda.synchronize()
    g = grads.clone()
    mod.optimizer_step()
    torch.cuda.synchronize()
    p = torch.from_numpy(mod.get_params().astype(np.int32)).to(dev)
    same_g = same_p = True
    if world > 1:
        for tensor, name in ((g, "g"), (p, "p")):
            ref = tensor.clone()
            dist.broadcast(ref, src=0)
            eq = torch.tensor([int(bool((ref == tensor).all()))], device=dev)
            dist.all_reduce(eq, op=dist.ReduceOp.MIN)
            if name == "g":
                same_g = bool(eq.item())
            else:
                same_p = bool(eq.item())
    ok, out = True, None
    if rank == 0:
        single = nbvh_b200.NeuralPathModule(cfg, device=local)
        single.init_params(1337)
        loss1 = torch.zeros(1, dtype=torch.float32, device=dev)
        g1 = single.gradient_buffer()
        single.training_forward_backward(torch.from_numpy(x).to(dev), torch.from_numpy(t).to(dev), loss_sum=loss1)
        torch.cuda.synchronize()
        rel = float(torch.linalg.norm(g - g1) / torch.linalg.norm(g1))
        dl = abs(float(loss.item()) - float(loss1.item())) / abs(float(loss1.item()))
        ok = rel <= 1e-3 and dl <= 1e-4 and same_g and same_p
        out = {"world": world, "global_batch": B, "gradient_rel_l2_vs_single_gpu": rel, "loss_rel_diff": dl, "loss": float(loss.item()),
               "gradients_identical_on_all_ranks": same_g, "parameters_identical_after_adam": same_p, "ok": ok}
        print(json.dumps(out), flush=True)
        single.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    mod.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
