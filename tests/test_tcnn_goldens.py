"""CPU tests pinning the oracle to outputs of the reference's OWN tcnn CUDA path (tests/golden/tcnn_*.npz, produced
on a B200 by tools/make_tcnn_goldens.py through oracle/_ref/tcnn_harness = the reference's tiny-cuda-nn fork +
PathTracingComposite loss rebuilt for sm_100).  Inputs are regenerated from the recorded seeds."""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tools.make_tcnn_goldens import N_MLP, golden_inputs  # noqa: E402

GOLDEN = os.path.join(REPO, "tests", "golden")
HALF_ULP_AT_HALF = 2.0 ** -11  # fp16 spacing in [0.5, 1)


def small_cfg(log2):
    return {"neural_bvh_renderer": {"neural_module": {
        "pt_output_constants": {"output_network_dims": 8, "output_normal_dim": 3, "output_albedo_dim": 3},
        "bvh_module_constants": {"input_segment_points": 3},
        "segment_encoding": {"hashgrid_options": {"log2_hashmap_size": log2}}}}}


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="module")
def small():
    g = np.load(os.path.join(GOLDEN, "tcnn_small.npz"))
    log2, nq, B, nb, steps, seed = [int(x) for x in g["meta"]]
    p, q, tin, tgt = golden_inputs(log2, nq, B, nb, seed)
    return dict(g=g, log2=log2, nq=nq, B=B, nb=nb, steps=steps, p=p, q=q, tin=tin, tgt=tgt, cfg=small_cfg(log2))


def test_query_fp16_accumulate_emulation_matches_tcnn(oracle, small):
    """The oracle's tcnn emulation (fp16 accumulate per 16-wide K chunk) reproduces tcnn's inference to within one
    fp16 ulp of the post-sigmoid output, and exactly for > 99 % of the outputs."""
    net = oracle.make_net(small["cfg"], small["p"])
    y = oracle.query(net, small["q"], mode=1)
    d = np.abs(y - small["g"]["query_out"])
    assert d.max() <= HALF_ULP_AT_HALF * 1.001
    assert (d == 0).mean() > 0.99


def test_query_fp32_accumulate_within_tolerance_of_tcnn(oracle, small):
    """fp16-in / fp32-accumulate arithmetic (what the B200 kernels compute) vs tcnn: max-abs 2e-3 on the outputs."""
    net = oracle.make_net(small["cfg"], small["p"])
    y = oracle.query(net, small["q"], mode=0)
    assert np.abs(y - small["g"]["query_out"]).max() < 2e-3


def test_chess_weights_query_matches_tcnn(oracle, chess_root):
    from nbvh_b200.scene_io import CHESS_BASE
    g = np.load(os.path.join(GOLDEN, "tcnn_chess.npz"))
    log2, nq, B, nb, steps, seed = [int(x) for x in g["meta"]]
    _, q, _, _ = golden_inputs(log2, nq, B, nb, seed)
    params = oracle.read_network_json(os.path.join(chess_root, CHESS_BASE + "_network.json"))
    net = oracle.make_net(small_cfg(18), params)
    y1 = oracle.query(net, q, mode=1)
    d1 = np.abs(y1 - g["query_out"])
    # trained weights: longer fp16 accumulation chains with large terms; tcnn's wmma adds inside a 16-chunk in fp16-ish
    # hardware order, so the emulation is close but not exact
    assert np.percentile(d1, 99) < 4e-3 and d1.max() < 3e-2
    y0 = oracle.query(net, q, mode=0)
    d0 = np.abs(y0 - g["query_out"])
    assert np.percentile(d0, 99) < 6e-3 and d0.max() < 3e-2


def test_training_step_matches_tcnn(oracle, small):
    """Loss values, per-output losses, gradients (loss scale 128) and the 30-step loss curve vs tcnn's Trainer."""
    g = small["g"]
    tr = oracle.Trainer(small["cfg"], small["p"])
    L, loss = tr.forward_backward(small["tin"][0], small["tgt"][0])
    assert abs(loss - g["train_loss"][0]) < 1e-5 * g["train_loss"][0]
    assert np.abs(L - g["train_L0"]).max() < 1e-6
    g0 = g["grads0"].view(np.float16).astype(np.float32)
    # tcnn accumulates weight gradients in fp16 (CUTLASS split-K) and grid gradients with fp16 atomics; the oracle
    # accumulates in fp32/double, so agreement is bounded by tcnn's own rounding noise
    off = 0
    for (o, i), tol in zip([(64, 96), (64, 64), (64, 64), (64, 64), (16, 64)], [0.06, 0.02, 0.015, 0.01, 0.005]):
        assert rel(tr.grads[off:off + o * i], g0[off:off + o * i]) < tol, (o, i)
        off += o * i
    assert rel(tr.grads[N_MLP:], g0[N_MLP:]) < 0.06
    nz_t = (g0[N_MLP:] != 0).sum()
    nz_o = (tr.grads[N_MLP:].astype(np.float16) != 0).sum()
    assert abs(int(nz_t) - int(nz_o)) < 0.01 * nz_t  # same set of touched grid entries (fp16 underflow aside)
    losses = [loss]
    tr.optimizer_step()
    p0 = small["p"].view(np.float16).astype(np.float32)
    up_t = g["params_after1"].view(np.float16).astype(np.float32) - p0
    up_o = tr.params.view(np.float16).astype(np.float32) - p0
    # first Adam step moves every touched parameter by ~lr * sign(g): agreement except for noise-level gradients
    assert (np.sign(up_t) == np.sign(up_o)).mean() > 0.97
    for s in range(1, small["steps"]):
        _, l = tr.training_step(small["tin"][s % small["nb"]], small["tgt"][s % small["nb"]])
        losses.append(l)
    losses = np.array(losses)
    assert np.abs(losses - g["train_loss"]).max() < 0.01 * g["train_loss"].max()
    assert losses[-1] < losses[0]


def test_learning_rate_schedule(oracle):
    """ExponentialDecay: x0.33 at steps 10000, 12000, ... (exponential_decay.h:60-71)."""
    L = oracle.lib()
    L.onbvh_learning_rate.restype = __import__("ctypes").c_float
    oc = oracle.OOpt(1e-2, 10000, 2000, 0.33, 0.9, 0.999, 1e-8, 1e-8)
    import ctypes as C
    lr = lambda s: L.onbvh_learning_rate(C.byref(oc), C.c_uint32(s))
    assert abs(lr(0) - 1e-2) < 1e-9 and abs(lr(9999) - 1e-2) < 1e-9
    assert abs(lr(10000) - 3.3e-3) < 1e-8 and abs(lr(11999) - 3.3e-3) < 1e-8
    assert abs(lr(12000) - 1.089e-3) < 1e-8
