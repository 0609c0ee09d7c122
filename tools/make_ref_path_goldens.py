#!/usr/bin/env python3
"""Rendered-image goldens + GPU timings of the reference's OWN hybrid path tracer (TEST INFRASTRUCTURE; run ON THE GPU BOX).

Drives oracle/_ref/ref_render_harness in `path` mode: the reference's wavefront kernels INCLUDING the shading kernel
hybrid_scatter_and_get_next_tlas_traversal_step (material unpack, Disney BSDF sampling, Russian roulette, environment
map), its host loop NeuralPathModule::neural_path_tracing and tiny-cuda-nn inference, on the shipped Chess scene / config /
tree-cut / weights / HDRI with the chessboard's own glTF material and textures:

  * 256x144, shipped max_depth 3 -> tests/golden/ref_chess_path_256x144.npz  (RGB radiance after 1, 16 and 256 spp)
  * 1920x1080 max_depth 3 (BASELINE config 2b) and 3840x2160 max_depth 4 (config 4) -> tests/golden/ref_chess_path_timing.json

Outputs land in gpurun_out/ref_golden/; copy the two files to tests/golden/ and commit."""
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
OUT = os.path.join(REPO, "gpurun_out", "ref_golden")
SPP = 16
SPP_CONVERGED = 256


def stage_shading(root, wd, blas):
    """Shading inputs next to the arrays make_ref_render_goldens.prepare wrote (large files are linked, not copied)."""
    for i, b in enumerate(blas):
        p = os.path.join(wd, "mesh%d" % i)
        for key, ext, dt in (("normals", "_normals.f32", "<f4"), ("texcoords", "_texcoords.f32", "<f4"),
                             ("material_ids", "_material_ids.i32", "<i4"), ("material_map", "_material_map.u32", "<u4")):
            if b.get(key) is not None:
                np.ascontiguousarray(b[key], dtype=dt).tofile(p + ext)
    for fn in os.listdir(root):
        if fn == "materials.bin" or fn.startswith("tex") or fn.startswith("envmap."):
            dst = os.path.join(wd, fn)
            if os.path.lexists(dst):
                os.remove(dst)
            os.symlink(os.path.join(root, fn), dst)


def run(harness, wd, W, H, max_depth, log2_hashmap, frames, spp):
    out = subprocess.run([harness, wd, str(W), str(H), str(max_depth), str(log2_hashmap), str(frames), "path", str(spp)],
                         capture_output=True, text=True)
    print(out.stdout[-1500:], out.stderr[-3000:], flush=True)
    if out.returncode != 0:
        raise SystemExit("ref_render_harness failed in " + wd)
    return json.load(open(os.path.join(wd, "ref_frame.json")))


def main():
    import nbvh_b200
    import make_ref_render_goldens as R
    from nbvh_b200.scene_io import CHESS_BASE, chess_fixture_dir
    root = chess_fixture_dir()
    cfg = json.load(open(os.path.join(root, CHESS_BASE + "_neural.json")))
    log2_hashmap = int(CHESS_BASE.split("hash_")[1].split("_")[0])
    mod = nbvh_b200.load_chess_module(max_depth=3, envmap_eval=True)  # constants only: the harness renders
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root, with_attributes=True)
    os.makedirs(OUT, exist_ok=True)
    gpu = subprocess.run(["nvidia-smi", "--query-gpu=name", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()

    W, H = 256, 144
    wd = os.path.join(OUT, "work_path_small")
    R.prepare(mod, cfg, wd, W, H, (tl, ti, bl))
    stage_shading(root, wd, bl)
    info = run(R.HARNESS, wd, W, H, 3, log2_hashmap, 3, SPP)
    a1 = np.fromfile(os.path.join(wd, "path_accum_1spp.f32"), dtype="<f4").reshape(H, W, 4)
    an = np.fromfile(os.path.join(wd, "path_accum.f32"), dtype="<f4").reshape(H, W, 4)
    run(R.HARNESS, wd, W, H, 3, log2_hashmap, 1, SPP_CONVERGED)
    ac = np.fromfile(os.path.join(wd, "path_accum.f32"), dtype="<f4").reshape(H, W, 4)
    np.savez_compressed(os.path.join(OUT, "ref_chess_path_256x144.npz"), W=W, H=H, max_depth=3, spp=SPP, spp_converged=SPP_CONVERGED,
                        rgb_1spp=a1[..., :3].copy(), rgb_nspp=(an[..., :3] / SPP).astype(np.float32),
                        rgb_converged=(ac[..., :3] / SPP_CONVERGED).astype(np.float32), queries_1spp=info["queries"])
    print("[path goldens] 256x144: mean radiance 1spp %.5f, %dspp %.5f" % (a1[..., :3].mean(), SPP, an[..., :3].mean() / SPP))

    timing = {"gpu": gpu}
    for name, (W, H, depth) in {"config2b_1080p_depth3": (1920, 1080, 3), "config4_2160p_depth4": (3840, 2160, 4)}.items():
        wd = os.path.join(OUT, "work_path_" + name)
        R.prepare(mod, cfg, wd, W, H, (tl, ti, bl))
        stage_shading(root, wd, bl)
        info = run(R.HARNESS, wd, W, H, depth, log2_hashmap, 10, 1)
        info["what"] = ("reference kernels incl. hybrid_scatter_and_get_next_tlas_traversal_step + tiny-cuda-nn inference rebuilt for sm_100, host loop "
                        "of path_module.cu:139-245, Chess shipped config + HDRI + chessboard material, 1 spp per frame")
        timing[name] = info
        print("[path goldens] %s: %.3f ms best, %d queries, %d launches, %d host syncs" %
              (name, info["ms_best"], info["queries"], info["launches"], info["host_syncs"]))
    json.dump(timing, open(os.path.join(OUT, "ref_chess_path_timing.json"), "w"), indent=1)
    for d in os.listdir(OUT):
        if d.startswith("work_path_"):
            for fn in os.listdir(os.path.join(OUT, d)):
                if not fn.endswith(".json"):
                    os.remove(os.path.join(OUT, d, fn))
    mod.close()


if __name__ == "__main__":
    main()
