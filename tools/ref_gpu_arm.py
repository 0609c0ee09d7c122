#!/usr/bin/env python3
"""BENCH INFRASTRUCTURE: the reference's OWN GPU path timed LIVE on the box bench.py runs on.

bench.py (rank 0, outside its timed regions) calls `measure()`: it stages the shipped Chess scene / config / tree-cut /
weights for the two reference-side harnesses that travel to the GPU box as prebuilt binaries

  * oracle/_ref/ref_render_harness  -- the reference's wavefront kernels (path_inference_kernels.cu,
    neural_bvh_module_kernels.cu) + its tiny-cuda-nn fork, compiled from /root/reference for sm_100 and driven in the
    order of NeuralPathModule::neural_path_tracing / neural_bvh_traversal
    (/root/reference/src/cuda/neural/path/path_module.cu:54-245, `network->inference` at :92), every host
    synchronisation of that loop included as the reference's own FrameTimer does;
  * oracle/_ref/tcnn_harness        -- `network->inference` on 2^20 queries and `trainer->training_step` at batch 2^18,

runs them with the same warm-up (3) and frame counts as bench.py's own arm, each under a ClockSampler, and returns
{"live": True, ...} blocks with ms, Mrays/s, launches, host syncs and the clocks seen during that run.  Nothing here is
product code: the harnesses are the checker / baseline, the product library is not involved."""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RENDER_HARNESS = os.path.join(REPO, "oracle", "_ref", "ref_render_harness")
TCNN_HARNESS = os.path.join(REPO, "oracle", "_ref", "tcnn_harness")


def available():
    return os.path.exists(RENDER_HARNESS) and os.access(RENDER_HARNESS, os.X_OK)


def _run(cmd, sampler_cls, device_index, timeout=300):
    smp = sampler_cls(device_index) if sampler_cls else None
    if smp:
        smp.start()
    t0 = time.time()
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    wall = time.time() - t0
    clocks = None
    if smp:
        smp.stop_flag = True
        smp.join()
        clocks = smp.summary()
    if out.returncode != 0:
        raise RuntimeError("%s failed (rc %d): %s" % (os.path.basename(cmd[0]), out.returncode, out.stderr[-400:]))
    return clocks, wall


def _stage_shading(root, wd, blas):
    for i, b in enumerate(blas):
        p = os.path.join(wd, "mesh%d" % i)
        for key, ext, dt in (("normals", "_normals.f32", "<f4"), ("texcoords", "_texcoords.f32", "<f4"),
                             ("material_ids", "_material_ids.i32", "<i4"), ("material_map", "_material_map.u32", "<u4")):
            if b.get(key) is not None:
                np.ascontiguousarray(b[key], dtype=dt).tofile(p + ext)
    for fn in os.listdir(root):
        if fn == "materials.bin" or fn.startswith("tex") or fn.startswith("envmap."):
            dst = os.path.join(wd, fn)
            if not os.path.lexists(dst):
                os.symlink(os.path.join(root, fn), dst)


def measure(root, sampler_cls=None, device_index=0, frames=20, do_path=True, do_train=True, log=None):
    """Returns a dict of live reference-GPU timings (see module docstring); raises when the harness is absent."""
    if not available():
        raise RuntimeError("oracle/_ref/ref_render_harness is absent (built in the authoring container from /root/reference)")
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tools"))
    import nbvh_b200
    import make_ref_render_goldens as R
    from nbvh_b200.scene_io import CHESS_BASE
    say = log or (lambda *_a: None)
    cfg = json.load(open(os.path.join(root, CHESS_BASE + "_neural.json")))
    log2_hashmap = int(CHESS_BASE.split("hash_")[1].split("_")[0])
    res = {"live": True, "harness": "oracle/_ref/ref_render_harness (reference kernels + tiny-cuda-nn rebuilt for sm_100, "
                                    "host loop of path_module.cu:54-245 incl. its host syncs)", "warmup_frames": 3}
    tmp = tempfile.mkdtemp(prefix="nbvh_refarm_")
    try:
        # ---- config 2a: 1080p, max_depth 1, envmap off (the headline workload)
        mod = nbvh_b200.load_chess_module(device=device_index, max_depth=1, envmap_eval=False)
        tl, ti, bl, _ = nbvh_b200.load_scene_arrays(root, with_attributes=True)
        wd = os.path.join(tmp, "c2a")
        R.prepare(mod, cfg, wd, 1920, 1080, (tl, ti, bl))
        mod.close()
        clocks, wall = _run([RENDER_HARNESS, wd, "1920", "1080", "1", str(log2_hashmap), str(frames), "timing"], sampler_cls, device_index)
        j = json.load(open(os.path.join(wd, "ref_frame.json")))
        res["config2a"] = {"ms_best": j["ms_best"], "ms_mean": j["ms_mean"], "Mrays_s_best": j["rays"] / j["ms_best"] / 1e3,
                           "Mrays_s_mean": j["rays"] / j["ms_mean"] / 1e3, "queries": j["queries"], "launches": j["launches"],
                           "host_syncs": j["host_syncs"], "frames_timed": j["frames_timed"], "clocks": clocks, "process_wall_s": wall}
        say("[ref arm] 2a: %.3f ms best / %.3f mean" % (j["ms_best"], j["ms_mean"]))
        shutil.rmtree(wd, ignore_errors=True)
        if do_path and nbvh_b200.has_shading_fixture(root):
            pmod = nbvh_b200.load_chess_module(device=device_index, max_depth=3, envmap_eval=True)
            for name, (W, H, depth, n) in {"config2b": (1920, 1080, 3, max(5, frames // 2)), "config4_1spp": (3840, 2160, 4, 5)}.items():
                wd = os.path.join(tmp, name)
                R.prepare(pmod, cfg, wd, W, H, (tl, ti, bl))
                _stage_shading(root, wd, bl)
                clocks, wall = _run([RENDER_HARNESS, wd, str(W), str(H), str(depth), str(log2_hashmap), str(n), "timing"], sampler_cls, device_index)
                j = json.load(open(os.path.join(wd, "ref_frame.json")))
                res[name] = {"ms_best": j["ms_best"], "ms_mean": j["ms_mean"], "Mpixels_s_best": j["rays"] / j["ms_best"] / 1e3,
                             "queries": j["queries"], "launches": j["launches"], "host_syncs": j["host_syncs"],
                             "frames_timed": j["frames_timed"], "clocks": clocks, "process_wall_s": wall}
                say("[ref arm] %s: %.3f ms best" % (name, j["ms_best"]))
                shutil.rmtree(wd, ignore_errors=True)
            pmod.close()
        if do_train:
            # ---- K10 sample generation at batch 2^18 (procedural materials / tree-cut statistics, as bench.py's own train block)
            import make_ref_sampler_goldens as S
            mod = nbvh_b200.load_chess_module(device=device_index, max_depth=1, envmap_eval=False)
            tl2, ti2, bl2, _ = nbvh_b200.load_scene_arrays(root, with_neural_mesh=True, with_attributes=True)
            mats, texs = nbvh_b200.procedural_materials(int(max(b["material_ids"].max() for b in bl2)) + 1)
            cut = nbvh_b200.procedural_tree_cut(len(mod.neural_bvh_nodes()), seed=11)
            wd = os.path.join(tmp, "sampler")
            R.prepare(mod, cfg, wd, 64, 64, (tl2, ti2, bl2))
            params = mod.get_params().copy()
            mod.close()
            S.write_sampler_inputs(wd, bl2, mats, texs, cut)
            clocks, wall = _run([RENDER_HARNESS, wd, "64", "64", "1", str(log2_hashmap), "0", "sampler", str(1 << 18), "0", "8", "1.0"],
                                sampler_cls, device_index)
            j = json.load(open(os.path.join(wd, "sampler.json")))
            res["sampler_2p18"] = {"ms_per_step": j["ms_per_step"], "steps_timed": j["steps"] - 1, "clocks": clocks, "process_wall_s": wall}
            say("[ref arm] sampler: %.3f ms" % j["ms_per_step"])
            shutil.rmtree(wd, ignore_errors=True)
            del tl2, ti2, bl2
            # ---- tcnn: network->inference on 2^20 queries, trainer->training_step at batch 2^18
            if os.path.exists(TCNN_HARNESS):
                wd = os.path.join(tmp, "tcnn")
                os.makedirs(wd)
                rng = np.random.default_rng(9)
                params.tofile(os.path.join(wd, "params_init.u16"))
                rng.uniform(0.02, 0.98, size=(1 << 20, 9)).astype("<f4").tofile(os.path.join(wd, "query_in.f32"))
                rng.uniform(0.02, 0.98, size=(1 << 18, 9)).astype("<f4").tofile(os.path.join(wd, "train_in.f32"))
                rng.uniform(0.0, 1.0, size=(1 << 18, 8)).astype("<f4").tofile(os.path.join(wd, "train_tgt.f32"))
                clocks, wall = _run([TCNN_HARNESS, wd, str(log2_hashmap), str(1 << 20), str(1 << 18), "20"], sampler_cls, device_index)
                j = json.load(open(os.path.join(wd, "timing.json")))
                res["tcnn"] = {"inference_2p20_ms": j["inference_ms"], "training_step_2p18_ms": j["training_step_ms"], "clocks": clocks,
                               "process_wall_s": wall}
                say("[ref arm] tcnn: inference %.3f ms, training_step %.3f ms" % (j["inference_ms"], j["training_step_ms"]))
                if "sampler_2p18" in res:
                    step = res["sampler_2p18"]["ms_per_step"] + j["training_step_ms"]
                    res["train_2p18"] = {"ms_per_step": step, "samples_s": (1 << 18) / (step * 1e-3),
                                         "what": "granular_sample_bvh_network_input_data + tcnn training_step, batch 2^18"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return res


if __name__ == "__main__":
    sys.path.insert(0, REPO)
    import nbvh_b200
    print(json.dumps(measure(nbvh_b200.chess_fixture_dir(), log=lambda s: print(s, file=sys.stderr)), indent=1))
