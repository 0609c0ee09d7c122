#!/usr/bin/env python3
"""Golden vectors + GPU timing from the reference's OWN wavefront kernels (TEST INFRASTRUCTURE; run ON THE GPU BOX).

Drives oracle/_ref/ref_render_harness (the reference's kernels + tcnn compiled from /root/reference by
`make -C oracle -f Makefile.tcnn render_harness`) on the shipped Chess scene / config / tree-cut / weights:

  * 256x144 frame  -> tests/golden/ref_chess_256x144.npz   (per-pixel rays, unconditional interval lists, classical
                      closest hits, first wavefront step's network inputs/outputs, final primary hit records)
  * 1920x1080 frame -> tests/golden/ref_chess_1080p.npz     (interval counts, hit kinds, per-row checksums) and
                      tests/golden/ref_chess_timing.json     (the reference frame time on this GPU)

Outputs land in gpurun_out/ref_golden/; copy the three files to tests/golden/ and commit.
The constant blocks handed to the harness are the PODs nbvh_b200 parsed from the shipped config (the harness checks
each blob's size against the reference's own sizeof)."""
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
HARNESS = os.path.join(REPO, "oracle", "_ref", "ref_render_harness")
OUT = os.path.join(REPO, "gpurun_out", "ref_golden")
MAX_IV = 32
HIT_DTYPE = np.dtype([("valid_hit", "<u4"), ("t", "<f4"), ("a", "<f4", 3), ("b", "<f4", 3)])
BLAS_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4"), ("mesh", "<u4")])


def row_checksums(count, node, mint, maxt, W):
    """Per-image-row u64 checksum of the ragged interval lists (order-sensitive within a pixel)."""
    k = np.arange(MAX_IV, dtype=np.uint64)[None, :]
    m = k < np.minimum(count, MAX_IV)[:, None].astype(np.uint64)
    mix = (node.astype(np.uint64) * np.uint64(2654435761) + mint.view(np.uint32).astype(np.uint64) * np.uint64(40503)
           + maxt.view(np.uint32).astype(np.uint64) * np.uint64(2246822519)) * (k + np.uint64(1))
    per_pixel = (mix * m).sum(axis=1, dtype=np.uint64) + count.astype(np.uint64)
    return per_pixel.reshape(-1, W).sum(axis=1, dtype=np.uint64)


def prepare(mod, cfg, wd, W, H, arrays):
    import nbvh_b200
    from nbvh_b200 import _capi as A
    os.makedirs(wd, exist_ok=True)
    mod.reinitialise_inference_data((W, H))
    cam = cfg["cuda_scene"]["camera"]
    mod.set_camera(cam["matrix"], cam["yfov"])
    names = {A.CONST_RENDERER: "renderer", A.CONST_SCENE_BBOX: "scene_bbox", A.CONST_MODULE: "module",
             A.CONST_PT_OUTPUT: "pt_output", A.CONST_PATH: "path", A.CONST_CAMERA: "camera", A.CONST_SCENE: "scene"}
    for which, name in names.items():
        open(os.path.join(wd, "const_%s.bin" % name), "wb").write(bytes(mod.get_constants(which)))
    tlas_nodes, tlas_idx, blas = arrays
    mod.neural_bvh_nodes().tofile(os.path.join(wd, "nbvh_nodes.bin"))
    tlas_nodes.tofile(os.path.join(wd, "tlas_nodes.bin"))
    tlas_idx.astype("<u4").tofile(os.path.join(wd, "tlas_indices.u32"))
    for i, b in enumerate(blas):
        p = os.path.join(wd, "mesh%d" % i)
        np.ascontiguousarray(b["nodes"]).tofile(p + "_nodes.bin")
        np.ascontiguousarray(b["prim_ids"]).astype("<u8").tofile(p + "_prim_ids.u64")
        np.ascontiguousarray(b["vertices"], dtype="<f4").tofile(p + "_vertices.f32")
        np.ascontiguousarray(b["indices"], dtype="<u4").tofile(p + "_indices.u32")
    mod.get_params().tofile(os.path.join(wd, "params.u16"))


def run(wd, W, H, max_depth, log2_hashmap, frames):
    out = subprocess.run([HARNESS, wd, str(W), str(H), str(max_depth), str(log2_hashmap), str(frames)], capture_output=True, text=True)
    print(out.stdout[-1500:], out.stderr[-3000:], flush=True)
    if out.returncode != 0:
        raise SystemExit("ref_render_harness failed in " + wd)
    return json.load(open(os.path.join(wd, "ref_frame.json")))


def load_outputs(wd, n):
    f = lambda name, dt: np.fromfile(os.path.join(wd, name), dtype=dt)
    return dict(rays=f("rays.f32", "<f4").reshape(n, 6), iv_count=f("iv_count.u32", "<u4"),
                iv_node=f("iv_node.u32", "<u4").reshape(n, MAX_IV), iv_mint=f("iv_mint.f32", "<f4").reshape(n, MAX_IV),
                iv_maxt=f("iv_maxt.f32", "<f4").reshape(n, MAX_IV), blas=f("blas_hit.bin", BLAS_DTYPE),
                hits=f("hits.bin", HIT_DTYPE), s0_in=f("step0_in.f32", "<f4").reshape(-1, 9),
                s0_out=f("step0_out.f32", "<f4").reshape(-1, 8), s0_px=f("step0_pixel.u32", "<u4"))


def main():
    import nbvh_b200
    from nbvh_b200.scene_io import CHESS_BASE, chess_fixture_dir
    root = chess_fixture_dir()
    cfg = json.load(open(os.path.join(root, CHESS_BASE + "_neural.json")))
    log2_hashmap = int(CHESS_BASE.split("hash_")[1].split("_")[0])
    mod = nbvh_b200.load_chess_module(max_depth=1, envmap_eval=False)
    tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root)
    os.makedirs(OUT, exist_ok=True)

    # ---- small frame: full per-pixel goldens
    W, H = 256, 144
    wd = os.path.join(OUT, "work_small")
    prepare(mod, cfg, wd, W, H, (tl, ti, bl))
    info_s = run(wd, W, H, 1, log2_hashmap, 3)
    o = load_outputs(wd, W * H)
    cnt = np.minimum(o["iv_count"], MAX_IV)
    m = np.arange(MAX_IV)[None, :] < cnt[:, None]
    order = np.argsort(o["s0_px"], kind="stable")  # queue slots are atomic-ordered: sort by pixel
    np.savez_compressed(
        os.path.join(OUT, "ref_chess_256x144.npz"), W=W, H=H, rays=o["rays"].view(np.uint32), iv_count=o["iv_count"].astype(np.uint16),
        iv_node=o["iv_node"][m], iv_mint=o["iv_mint"][m].view(np.uint32), iv_maxt=o["iv_maxt"][m].view(np.uint32),
        blas_t=o["blas"]["t"].view(np.uint32), blas_u=o["blas"]["u"].view(np.uint32), blas_v=o["blas"]["v"].view(np.uint32),
        blas_prim=o["blas"]["prim"], blas_mesh=o["blas"]["mesh"].astype(np.uint8),
        hit_kind=o["hits"]["valid_hit"].astype(np.uint8), hit_t=o["hits"]["t"].view(np.uint32),
        hit_a=o["hits"]["a"].view(np.uint32), hit_b=o["hits"]["b"].view(np.uint32),
        s0_px=o["s0_px"][order], s0_in=o["s0_in"][order].view(np.uint32), s0_out=o["s0_out"][order].astype(np.float16),
        queries=info_s["queries"])
    print("[goldens] 256x144: %d intervals, %d first-step queries, hit kinds %s" %
          (int(cnt.sum()), len(order), np.bincount(o["hits"]["valid_hit"], minlength=3).tolist()))

    # ---- full frame: summaries + the reference's frame time on this GPU
    W, H = 1920, 1080
    wd = os.path.join(OUT, "work_full")
    prepare(mod, cfg, wd, W, H, (tl, ti, bl))
    info = run(wd, W, H, 1, log2_hashmap, 20)
    o = load_outputs(wd, W * H)
    np.savez_compressed(
        os.path.join(OUT, "ref_chess_1080p.npz"), W=W, H=H, iv_count=o["iv_count"].astype(np.uint8),
        iv_row_checksum=row_checksums(o["iv_count"], o["iv_node"], o["iv_mint"], o["iv_maxt"], W),
        hit_kind=o["hits"]["valid_hit"].astype(np.uint8), blas_prim_row_sum=o["blas"]["prim"].astype(np.uint64).reshape(H, W).sum(axis=1),
        blas_t_row_sum=o["blas"]["t"].view(np.uint32).astype(np.uint64).reshape(H, W).sum(axis=1), queries=info["queries"])
    info["gpu"] = subprocess.run(["nvidia-smi", "--query-gpu=name", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    info["mrays_s_best"] = W * H / info["ms_best"] / 1e3
    info["what"] = ("reference kernels (path_inference_kernels.cu, neural_bvh_module_kernels.cu) + tiny-cuda-nn inference rebuilt for sm_100, "
                    "host loop of path_module.cu:54-245, Chess shipped config, max_depth 1, envmap off, 1 spp")
    json.dump(info, open(os.path.join(OUT, "ref_chess_timing.json"), "w"), indent=1)
    print("[goldens] 1080p reference frame: %.3f ms best (%.1f Mrays/s), %d queries, %d launches, %d host syncs" %
          (info["ms_best"], info["mrays_s_best"], info["queries"], info["launches"], info["host_syncs"]))
    # keep gpurun_out small: the raw work dirs are scratch
    for d in ("work_small", "work_full"):
        for fn in os.listdir(os.path.join(OUT, d)):
            if not fn.endswith(".json"):
                os.remove(os.path.join(OUT, d, fn))
    mod.close()


if __name__ == "__main__":
    main()
