#!/usr/bin/env python3
"""Generate the Chess fixtures the GPU box needs (it has no /root/reference).

TEST/BENCH INFRASTRUCTURE.  Runs only in the authoring container, where /root/reference exists:

  * copies the shipped DATA files the hot path loads unchanged (config JSON, .nbvh tree-cut,
    network_models weights JSON) into baseline/_ref/chess/  (git-ignored, travels with gpurun);
  * parses chess.gltf (JSON only) and drives oracle/ref_tools/scene_tool.cpp -- compiled against the
    reference's own vendored glm + ext/bvh headers where they lie -- to produce the flattened neural
    mesh, the non-neural meshes, their BLAS (reference builder settings) and the TLAS as raw arrays.

No reference *source* is copied into the repo; only data and built arrays land in baseline/_ref/.
Scene-flattening order follows /root/reference/src/base/scene.cpp:143-262 (children before the
node's own mesh; neural sub-scene first; meshes whose glTF mesh name already went into the neural
mesh are skipped in the other scenes).
"""
import json
import os
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
SCENE_DIR = os.path.join(REF, "scene_data/scenes_path_tracing/chess")
BASE = "nodes_32940_hash_18_lvl_8_tsteps_10000"
OUT = os.path.join(REPO, "baseline/_ref/chess")

COMP_SIZE = {5123: 2, 5125: 4, 5126: 4}
TYPE_N = {"SCALAR": 1, "VEC2": 2, "VEC3": 3}


def accessor_offset(g, idx):
    a = g["accessors"][idx]
    bv = g["bufferViews"][a["bufferView"]]
    assert "byteStride" not in bv, "strided buffer views not needed for the shipped scene"
    return bv.get("byteOffset", 0) + a.get("byteOffset", 0), a["count"], a["componentType"]


def node_line(n):
    s = "node"
    if "matrix" in n:
        raise SystemExit("matrix node transforms not needed for the shipped scene")
    if "translation" in n:
        s += " T " + " ".join(repr(float(x)) for x in n["translation"])
    if "rotation" in n:
        s += " R " + " ".join(repr(float(x)) for x in n["rotation"])
    if "scale" in n:
        s += " S " + " ".join(repr(float(x)) for x in n["scale"])
    return s


def emit_node(g, node_idx, chain, lines, neural, neural_names, separate):
    n = g["nodes"][node_idx]
    chain = chain + [n]
    for c in n.get("children", []):
        emit_node(g, c, chain, lines, neural, neural_names, separate)
    if "mesh" not in n:
        return
    mesh = g["meshes"][n["mesh"]]
    if neural:
        neural_names.add(mesh["name"])
    elif mesh["name"] in neural_names:
        return
    if separate:
        lines.append("mesh %s" % mesh["name"].replace(" ", "_"))
    lines.append("inst")
    for cn in chain:
        lines.append(node_line(cn))
    for p in mesh["primitives"]:
        assert p.get("mode", 4) == 4
        pos_off, pos_cnt, _ = accessor_offset(g, p["attributes"]["POSITION"])
        nrm_off = accessor_offset(g, p["attributes"]["NORMAL"])[0] if "NORMAL" in p["attributes"] else -1
        uv_off = accessor_offset(g, p["attributes"]["TEXCOORD_0"])[0] if "TEXCOORD_0" in p["attributes"] else -1
        idx_off, idx_cnt, idx_ct = accessor_offset(g, p["indices"])
        lines.append("prim %d %d %d %d %d %d %d %d" % (pos_off, pos_cnt, nrm_off, uv_off, idx_off, idx_cnt,
                                                        COMP_SIZE[idx_ct], p.get("material", -1)))


def write_shading_fixture(g):
    """Materials, the textures the NON-neural meshes reference, and the shipped environment map, for the path-tracing
    rows.  Image decoding belongs to the host application (tinygltf/stb and tinyexr in the reference);
    here OpenCV decodes the same files once, in the authoring container:
      materials.bin     glTF material i -> 96-byte PackedMaterial, rules of /root/reference/src/base/scene.cpp:308-417
                        (tinygltf defaults: baseColorFactor 1, metallicFactor 1, roughnessFactor 1, emissiveFactor 0)
      tex<k>.rgba8/.dims  image k as RGBA8 (tinygltf expands to 4 channels); images no non-neural mesh uses -> 1x1 white
      envmap.f32/.dims  scene_data/hdris/<envmap_filename> as RGBA32F (tinyexr LoadEXR layout)"""
    import numpy as np
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    if os.path.exists(os.path.join(OUT, "materials.bin")) and os.path.exists(os.path.join(OUT, "envmap.f32")) and "--force" not in sys.argv:
        print("[make_fixtures] materials / textures / envmap already present")
        return
    sys.path.insert(0, REPO)
    from nbvh_b200.module import MATERIAL_DTYPE

    def handle(tex_id, channel=0):
        return np.array([0x80000000 | ((channel & 3) << 29) | (tex_id & 0x1FFFFFFF)], dtype=np.uint32).view(np.float32)[0]

    mats = np.zeros(len(g["materials"]), dtype=MATERIAL_DTYPE)
    for i, m in enumerate(g["materials"]):
        pbr = m.get("pbrMetallicRoughness", {})
        r = mats[i]
        r["base_color"] = np.asarray(pbr.get("baseColorFactor", [1, 1, 1, 1])[:3], np.float32)
        r["metallic"] = pbr.get("metallicFactor", 1.0)
        r["roughness"] = pbr.get("roughnessFactor", 1.0)
        r["emission_scale"] = 1.0
        r["ior"] = 1.5
        if "baseColorTexture" in pbr:
            r["base_color"][0] = handle(g["textures"][pbr["baseColorTexture"]["index"]]["source"])
        if "normalTexture" in m:
            r["normal"][0] = handle(g["textures"][m["normalTexture"]["index"]]["source"])
        if "metallicRoughnessTexture" in pbr:
            src = g["textures"][pbr["metallicRoughnessTexture"]["index"]]["source"]
            r["metallic"] = handle(src, 2)
            r["roughness"] = handle(src, 1)
        ext = m.get("extensions", {})
        if "KHR_materials_transmission" in ext:
            r["specular_transmission"] = ext["KHR_materials_transmission"].get("transmissionFactor", 0.0)
        if "KHR_materials_ior" in ext:
            r["ior"] = ext["KHR_materials_ior"].get("ior", 1.5)
        r["base_emission"] = np.asarray(m.get("emissiveFactor", [0, 0, 0]), np.float32)
        if "emissiveTexture" in m:
            r["base_emission"][0] = handle(g["textures"][m["emissiveTexture"]["index"]]["source"])
        if "KHR_materials_emissive_strength" in ext:
            r["emission_scale"] = ext["KHR_materials_emissive_strength"].get("emissiveStrength", 1.0)
    mats.tofile(os.path.join(OUT, "materials.bin"))

    # images referenced by the materials of the non-neural meshes
    man = json.load(open(os.path.join(OUT, "manifest.json")))
    used = set()
    for mi in range(len(man["meshes"])):
        if mi == man["neural_mesh"]:
            continue
        for mid in np.fromfile(os.path.join(OUT, "mesh%d_material_ids.i32" % mi), dtype="<i4"):
            if mid < 0:
                continue
            for fld in ("base_color", "metallic", "roughness", "normal", "base_emission"):
                v = np.atleast_1d(mats[mid][fld])[:1].view(np.uint32)[0]
                if v & 0x80000000:
                    used.add(int(v & 0x1FFFFFFF))
    for k, img in enumerate(g.get("images", [])):
        if k in used:
            bgr = cv2.imread(os.path.join(SCENE_DIR, img["uri"]), cv2.IMREAD_COLOR)
            rgba = np.dstack([bgr[..., ::-1], np.full(bgr.shape[:2], 255, np.uint8)])
        else:
            rgba = np.full((1, 1, 4), 255, np.uint8)
        np.ascontiguousarray(rgba).tofile(os.path.join(OUT, "tex%d.rgba8" % k))
        np.asarray([rgba.shape[1], rgba.shape[0]], np.uint32).tofile(os.path.join(OUT, "tex%d.dims" % k))
    cfg = json.load(open(os.path.join(OUT, BASE + "_neural.json")))["cuda_scene"]
    exr = cv2.imread(os.path.join(REF, "scene_data/hdris", cfg["envmap_filename"]), cv2.IMREAD_UNCHANGED)
    assert exr is not None and exr.dtype == np.float32, "EXR decode failed"
    env = np.dstack([exr[..., 2], exr[..., 1], exr[..., 0], exr[..., 3] if exr.shape[2] > 3 else np.ones(exr.shape[:2], np.float32)])
    np.ascontiguousarray(env, dtype=np.float32).tofile(os.path.join(OUT, "envmap.f32"))
    np.asarray([env.shape[1], env.shape[0]], np.uint32).tofile(os.path.join(OUT, "envmap.dims"))
    print("[make_fixtures] %d materials, textures used by non-neural meshes: %s, envmap %dx%d" % (len(mats), sorted(used), env.shape[1], env.shape[0]))


def main():
    if not os.path.isdir(SCENE_DIR):
        print("[make_fixtures] no reference tree here; nothing to do")
        return 0
    os.makedirs(OUT, exist_ok=True)
    # 1) shipped data files, unchanged
    shutil.copy(os.path.join(SCENE_DIR, "configs", BASE + "_neural.json"), os.path.join(OUT, BASE + "_neural.json"))
    shutil.copy(os.path.join(SCENE_DIR, "nbvhs", BASE + ".nbvh"), os.path.join(OUT, BASE + ".nbvh"))
    shutil.copy(os.path.join(SCENE_DIR, "network_models", BASE + "_network.json"),
                os.path.join(OUT, BASE + "_network.json"))
    # a second, partial config to exercise defaults (83-line variant)
    shutil.copy(os.path.join(SCENE_DIR, "configs", "nodes_155_hash_14_lvl_8_tsteps_10000_neural.json"),
                os.path.join(OUT, "nodes_155_hash_14_lvl_8_tsteps_10000_neural.json"))

    # every shipped configuration of the scene (7 node budgets x 5 hash-table sizes), unchanged: tests/test_shipped_configs.py
    os.makedirs(os.path.join(OUT, "configs"), exist_ok=True)
    for name in sorted(os.listdir(os.path.join(SCENE_DIR, "configs"))):
        if name.endswith("_neural.json"):
            shutil.copy(os.path.join(SCENE_DIR, "configs", name), os.path.join(OUT, "configs", name))

    g = json.load(open(os.path.join(SCENE_DIR, "chess.gltf")))
    if os.path.exists(os.path.join(OUT, "manifest.json")) and "--force" not in sys.argv:
        print("[make_fixtures] scene arrays already present (use --force to rebuild)")
        write_shading_fixture(g)
        return 0

    # 2) scene arrays through the reference's own glm + ext/bvh
    neural_scenes = [i for i, s in enumerate(g["scenes"]) if "neural" in s.get("name", "")]
    if not neural_scenes:
        neural_scenes = [0]
    lines = ["bin " + os.path.join(SCENE_DIR, g["buffers"][0]["uri"]), "out " + OUT, "neural 0"]
    cfg = json.load(open(os.path.join(OUT, BASE + "_neural.json")))["cuda_scene"]
    lines.append("camera %d %d %r %s" % (cfg["fb_size"][0], cfg["fb_size"][1], cfg["camera"]["yfov"],
                                         " ".join(repr(float(x)) for x in cfg["camera"]["matrix"])))
    neural_names = set()
    lines.append("mesh neural_mesh")
    for si in neural_scenes[:1]:
        for ni in g["scenes"][si]["nodes"]:
            emit_node(g, ni, [], lines, True, neural_names, False)
    for si, s in enumerate(g["scenes"]):
        if si in neural_scenes:
            continue
        for ni in s["nodes"]:
            emit_node(g, ni, [], lines, False, neural_names, True)
    manifest = os.path.join(OUT, "scene_manifest.txt")
    open(manifest, "w").write("\n".join(lines) + "\n")

    tool = os.path.join(REPO, "oracle/_ref/scene_tool")
    os.makedirs(os.path.dirname(tool), exist_ok=True)
    cmd = ["g++", "-std=c++20", "-O2", "-pthread", "-I", os.path.join(REF, "ext/bvh/src"), "-I",
           os.path.join(REF, "ext/glm"), os.path.join(REPO, "oracle/ref_tools/scene_tool.cpp"), "-o", tool]
    print("[make_fixtures]", " ".join(cmd))
    subprocess.check_call(cmd)
    subprocess.check_call([tool, manifest])
    write_shading_fixture(g)
    return 0


if __name__ == "__main__":
    sys.exit(main())
