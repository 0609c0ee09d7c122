#!/usr/bin/env python3
"""Goldens for LDR-FLIP, produced by the reference's own code on seeded image pairs:
  python tools/make_ref_flip_goldens.py py    -> tests/golden/ref_flip_py_64x48.npz   AUTHORING CONTAINER: imports the reference's
                                                 figures/flip/flip_api.py (compute_ldrflip) from /root/reference
  python tools/make_ref_flip_goldens.py cuda  -> gpurun_out/ref_flip_cuda_160x90.npz  GPU BOX: runs oracle/_ref/flip_harness, i.e. the
                                                 reference's vendored flip-cuda driven like compute_and_display_flip_error
Both files hold the two linear-RGB input images and the error map; copy the second into tests/golden/ after the GPU call."""
import os
import subprocess
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def image_pair(w, h, seed, hdr):
    """Smooth colour gradients + edges + noise + a darkened block; hdr adds values above 1 and exact zeros (clamp branches)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    ref = np.stack([0.5 + 0.5 * np.sin(xx / 5.0), 0.5 + 0.5 * np.cos(yy / 7.0), (xx + yy) / float(h + w)], -1) ** 2.2
    ref[(xx // 16 + yy // 12) % 2 == 0] *= 0.6
    test = np.clip(ref + rng.normal(0, 0.05, ref.shape), 0, None)
    test[h // 5:h // 5 + 10, w // 2:w // 2 + 12] *= 0.3
    if hdr:
        ref[h // 2:h // 2 + 8, 4:20] *= 3.5
        test[h // 2:h // 2 + 8, 10:26] *= 2.5
        ref[-6:, :12] = 0.0
        test[-6:, 6:18] = 0.0
    else:
        ref, test = np.clip(ref, 0, 1), np.clip(test, 0, 1)
    return ref.astype(np.float32), test.astype(np.float32)


F = np.float32
DEFAULT_PPD = float(F(0.7) * (F(3840.0) / F(0.7)) * (F(3.14159265358979) / F(180.0)))  # neural_bvh_renderer.cu:34-38,43-47


def linear_to_srgb(l):  # the reference converts both images with LinearRGB2sRGB before FLIP (neural_bvh_renderer.cu:65-66)
    l = l.astype(F)
    return np.where(l <= F(0.0031308), l * F(12.92), F(1.055) * np.power(l, F(1.0 / 2.4), dtype=F) - F(0.055)).astype(F)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "py"
    if mode == "py":
        import types
        sys.path.insert(0, "/root/reference/figures")
        for missing in ("OpenEXR", "Imath"):  # flip/data.py imports the EXR bindings for file I/O only
            sys.modules.setdefault(missing, types.ModuleType(missing))
        from flip.flip_api import compute_ldrflip
        ref, test = image_pair(64, 48, 7, hdr=False)
        err = np.squeeze(compute_ldrflip(linear_to_srgb(ref).transpose(2, 0, 1), linear_to_srgb(test).transpose(2, 0, 1), DEFAULT_PPD))
        out = os.path.join(REPO, "tests", "golden", "ref_flip_py_64x48.npz")
        np.savez_compressed(out, reference=ref, test=test, error_map=err.astype(np.float32), ppd=np.float32(DEFAULT_PPD))
        print("wrote", out, "mean", float(err.mean()))
        return 0
    tool = os.path.join(REPO, "oracle", "_ref", "flip_harness")
    ref, test = image_pair(160, 90, 11, hdr=True)
    with tempfile.TemporaryDirectory() as d:
        ref.tofile(os.path.join(d, "r.f32"))
        test.tofile(os.path.join(d, "t.f32"))
        r = subprocess.run([tool, "160", "90", os.path.join(d, "r.f32"), os.path.join(d, "t.f32"), os.path.join(d, "e.f32")],
                           capture_output=True, text=True, check=True)
        print(r.stdout.strip())
        err = np.fromfile(os.path.join(d, "e.f32"), "<f4").reshape(90, 160)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    out = os.path.join(REPO, "gpurun_out", "ref_flip_cuda_160x90.npz")
    np.savez_compressed(out, reference=ref, test=test, error_map=err, ppd=np.float32(DEFAULT_PPD))
    print("wrote", out, "mean", float(err.mean()))
    return 0


if __name__ == "__main__":
    sys.exit(main())
