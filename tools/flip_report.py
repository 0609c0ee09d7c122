#!/usr/bin/env python3
"""LDR-FLIP (the reference's own figures/flip/flip_api.py, NVIDIA's implementation) and relMSE between the B200 path-traced image and
the reference path tracer's image of the same view.  AUTHORING CONTAINER ONLY (imports from /root/reference).
usage: python tools/flip_report.py gpurun_out/b200_chess_path_256x144_256spp.npy profiles/<name>.json"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import types  # noqa: E402

sys.path.insert(0, "/root/reference/figures")
for missing in ("OpenEXR", "Imath"):  # flip/data.py imports the EXR bindings for file I/O only; compute_ldrflip never touches them
    sys.modules.setdefault(missing, types.ModuleType(missing))
from flip.flip_api import compute_ldrflip  # noqa: E402


def srgb(x):  # linear_to_srgb of the reference's framebuffer (neural_bvh_renderer_kernels.cuh)
    x = np.clip(x, 0.0, 1.0)
    return np.where(x <= 0.0031308, 12.92 * x, 1.055 * np.power(x, 1 / 2.4) - 0.055)


def main():
    ours = np.load(sys.argv[1])
    g = np.load(os.path.join(REPO, "tests", "golden", "ref_chess_path_256x144.npz"))
    ref = g["rgb_converged"]
    ppd = (0.7 * 3840 / 0.7) * np.pi / 180  # FLIP's default observer (flip_api.py:439)
    flip = compute_ldrflip(srgb(ref).transpose(2, 0, 1).astype(np.float32), srgb(ours).transpose(2, 0, 1).astype(np.float32), ppd)
    # noise floor of the metric at this sample count: the reference's 16-spp image against its own 256-spp image
    flip_ref16 = compute_ldrflip(srgb(ref).transpose(2, 0, 1).astype(np.float32), srgb(g["rgb_nspp"]).transpose(2, 0, 1).astype(np.float32), ppd)
    out = {"view": "Chess 256x144, shipped max_depth 3, %d spp" % int(g["spp_converged"]),
           "flip_mean_b200_vs_reference": float(flip.mean()), "flip_p99": float(np.quantile(flip, 0.99)),
           "relmse_b200_vs_reference": float((np.square(ref - ours) / (ref + 0.1)).mean()),
           "flip_mean_reference_16spp_vs_reference_256spp": float(flip_ref16.mean()),
           "mean_radiance": [float(ours.mean()), float(ref.mean())],
           "how": "LDR-FLIP by the reference's own figures/flip/flip_api.py (compute_ldrflip, default observer) on linear_to_srgb images"}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
