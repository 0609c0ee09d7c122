"""LDR-FLIP (replaces ext/flip-cuda behind compute_and_display_flip_error, neural_bvh_renderer.cu:34-96).

CPU: the numpy restatement (oracle/flip.py) against error maps produced by the reference's own code on seeded image pairs --
     figures/flip/flip_api.py (ref_flip_py_64x48.npz) and the vendored flip-cuda run on a B200 (ref_flip_cuda_160x90.npz, HDR values
     and exact zeros included) -- tools/make_ref_flip_goldens.py.  Tolerance 2e-4 on the error map (values in [0, 1], fp32 pipeline
     with pow() calls; flip-cuda itself is built with FMA contraction).
GPU: nbvh_flip_images / nbvh_flip_error against the same goldens and the oracle, same tolerance; the mean within 1e-5."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-4


def _golden(name):
    p = os.path.join(GOLD, name)
    if not os.path.exists(p):
        pytest.skip(name + " not present")
    return np.load(p)


@pytest.mark.parametrize("name", ["ref_flip_py_64x48.npz", "ref_flip_cuda_160x90.npz"])
def test_oracle_matches_reference_flip(name):
    from oracle import flip as OF
    g = _golden(name)
    err = OF.flip(g["reference"], g["test"], float(g["ppd"]))
    assert g["error_map"].std() > 0.01
    assert np.abs(err - g["error_map"]).max() < TOL


def test_oracle_flip_properties():
    from oracle import flip as OF
    g = _golden("ref_flip_py_64x48.npz")
    assert np.abs(OF.flip(g["reference"], g["reference"])).max() == 0.0            # identical images
    e = OF.flip(np.zeros((32, 40, 3), np.float32), np.ones((32, 40, 3), np.float32))
    assert e.min() > 0.9 and e.max() <= 1.0                                        # black vs white
    rs, _, rf, (gg, dg, ddg) = OF.filters(OF.DEFAULT_PPD)
    assert (rs, rf) == (10, 9) and abs(gg.sum() - 1) < 1e-6 and abs(dg[dg > 0].sum() - 1) < 1e-6 and abs(ddg[ddg > 0].sum() - 1) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ref_flip_py_64x48.npz", "ref_flip_cuda_160x90.npz"])
def test_gpu_flip_matches_reference_flip(name, synthetic):
    import torch
    import nbvh_b200
    from oracle import flip as OF
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    g = _golden(name)
    mod = nbvh_b200.NeuralPathModule(synthetic[0])

    def rgba(a, scale=1.0):
        return torch.from_numpy(np.concatenate([a * np.float32(scale), np.ones(a.shape[:2] + (1,), np.float32)], -1)).cuda().contiguous()

    mean, emap = mod.flip_images(rgba(g["reference"]), rgba(g["test"]), ppd=float(g["ppd"]))
    emap = emap.cpu().numpy()
    assert np.abs(emap - g["error_map"]).max() < TOL
    assert abs(mean - float(g["error_map"].astype(np.float64).mean())) < 1e-5
    assert np.abs(emap - OF.flip(g["reference"], g["test"], float(g["ppd"]))).max() < TOL
    # the test image divided by the sample count, default viewing conditions, identical images, bad arguments
    mean4, emap4 = mod.flip_images(rgba(g["reference"]), rgba(g["test"], 4.0), test_divisor=4.0)
    assert np.abs(emap4.cpu().numpy() - g["error_map"]).max() < TOL
    mean0, emap0 = mod.flip_images(rgba(g["reference"]), rgba(g["reference"]))
    assert mean0 == 0.0 and float(emap0.abs().max()) == 0.0
    with pytest.raises(nbvh_b200.NbvhError):
        mod.flip_images(rgba(g["reference"]), rgba(g["test"]), ppd=1e4)            # filter radius out of range
    mod.close()


@pytest.mark.gpu
def test_gpu_flip_error_of_accumulated_frame(synthetic):
    """nbvh_flip_error == FLIP(reference image, accum / spp): a 4-spp frame against a shifted, dimmed copy of the 1-spp frame."""
    import torch
    import nbvh_b200
    from oracle import flip as OF
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg, nbvh, tlas, tlas_idx, blas = synthetic
    w, h = cfg["cuda_scene"]["fb_size"]
    mod = nbvh_b200.NeuralPathModule(cfg)
    mod.set_neural_bvh(nbvh)
    mod.set_scene(tlas, tlas_idx, blas)
    mod.init_params(1337)
    sc = mod.get_constants(nbvh_b200.CONST_SCENE)
    sc.envmap_eval, sc.skylike_background = True, True      # a sky gradient behind the geometry instead of a black background
    mod.set_constants(nbvh_b200.CONST_SCENE, sc)
    accum, _ = mod.attach_framebuffer((w, h))
    mod.neural_render(0, 0)
    torch.cuda.synchronize()
    one = (torch.roll(accum, 3, dims=1) * 0.8).contiguous()
    for s in range(1, 4):
        mod.neural_render(s, 0)
    torch.cuda.synchronize()
    mean, emap = mod.flip_error(one, 4, want_map=True)
    want = OF.flip(one.cpu().numpy()[..., :3], accum.cpu().numpy()[..., :3] / np.float32(4.0))
    assert want.mean() > 1e-3                                                      # edges moved by three pixels
    assert np.abs(emap.cpu().numpy() - want).max() < TOL and abs(mean - float(want.astype(np.float64).mean())) < 1e-5
    with pytest.raises(nbvh_b200.NbvhError):
        mod.flip_error(one, 0)
    mod.close()
