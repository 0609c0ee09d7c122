"""Harness-side scene loading: stands in for the reference's CudaBackend (scene upload) when driving the module.

  * load_scene_arrays / load_chess_module: the Chess fixture written by tools/make_fixtures.py
    (baseline/_ref/chess: BLAS/TLAS built by the reference's own ext/bvh settings, shipped config / .nbvh / weights)
  * synthetic_scene: procedural scene + tree-cut + random weights of the shipped network shape, for boxes
    without the fixture and for the synthetic configuration of BASELINE.json.
"""
import json
import os

import numpy as np

from .module import MATERIAL_DTYPE, NODE_DTYPE, TREE_CUT_DTYPE, NeuralPathModule

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHESS_BASE = "nodes_32940_hash_18_lvl_8_tsteps_10000"


def chess_fixture_dir():
    d = os.path.join(REPO, "baseline", "_ref", "chess")
    return d if os.path.exists(os.path.join(d, "manifest.json")) else None


def texture_handle(tex_id, channel=0):
    """A material parameter that refers to a texture: sign bit set, channel in bits 29-30, id below
    (include/base/material.h:8-28), as the float32 whose bit pattern it is."""
    return np.array([0x80000000 | ((channel & 3) << 29) | (tex_id & 0x1FFFFFFF)], dtype=np.uint32).view(np.float32)[0]


def procedural_materials(n_materials, tex_size=64, seed=7):
    """Deterministic stand-in for the glTF materials (JPEG decoding belongs to the host application): every second
    material has a base-colour texture, every third a tangent-space normal map; smooth RGBA8 patterns.
    Returns (MATERIAL_DTYPE array, list of (H, W, 4) uint8 textures)."""
    rng = np.random.default_rng(seed)
    mats = np.zeros(n_materials, dtype=MATERIAL_DTYPE)
    mats["base_color"] = 0.9
    mats["roughness"] = 1.0
    mats["emission_scale"] = 1.0
    mats["ior"] = 1.5
    yy, xx = np.meshgrid(np.arange(tex_size) / tex_size, np.arange(tex_size) / tex_size, indexing="ij")
    textures = []
    for m in range(n_materials):
        if m % 2 == 0:
            img = np.zeros((tex_size, tex_size, 4), np.float64)
            for ch in range(3):
                kx, ky = rng.integers(1, 4, 2)
                img[..., ch] = 0.5 + 0.4 * np.sin(2 * np.pi * (kx * xx + ky * yy + rng.uniform()))
            img[..., 3] = 1.0
            mats["base_color"][m, 0] = texture_handle(len(textures))
            textures.append(np.clip(np.rint(img * 255), 0, 255).astype(np.uint8))
        else:
            mats["base_color"][m] = rng.uniform(0.1, 0.9, 3).astype(np.float32)
        if m % 3 == 0:
            kx, ky = rng.integers(1, 3, 2)
            nx = 0.25 * np.sin(2 * np.pi * (kx * xx + rng.uniform()))
            ny = 0.25 * np.cos(2 * np.pi * (ky * yy + rng.uniform()))
            img = np.stack([0.5 + 0.5 * nx, 0.5 + 0.5 * ny, np.full_like(nx, 1.0), np.ones_like(nx)], axis=-1)
            mats["normal"][m, 0] = texture_handle(len(textures))
            textures.append(np.clip(np.rint(img * 255), 0, 255).astype(np.uint8))
    return mats, textures


def procedural_tree_cut(n_nodes, seed=11):
    """Deterministic NeuralTreeCutData stand-in (the renderer owns the real statistics): ~10 % of the nodes unsampled."""
    rng = np.random.default_rng(seed)
    cut = np.zeros(n_nodes, dtype=TREE_CUT_DTYPE)
    cut["loss"] = rng.uniform(0.0, 400.0, n_nodes).astype(np.float32)
    cut["iter_alive"] = rng.integers(1, 100, n_nodes)
    cut["depth"] = rng.integers(0, 20, n_nodes)
    cut["samples"] = rng.integers(1, 5000, n_nodes) * (rng.uniform(size=n_nodes) > 0.1)
    return cut


def load_scene_arrays(root, with_neural_mesh=False, with_attributes=False):
    man = json.load(open(os.path.join(root, "manifest.json")))
    blas = []
    for i, m in enumerate(man["meshes"]):
        p = os.path.join(root, "mesh%d" % i)
        if i == man["neural_mesh"] and not with_neural_mesh:
            # placeholder BLAS: never traversed while ignore_neural_blas is false
            blas.append(dict(nodes=np.zeros(1, NODE_DTYPE), prim_ids=np.zeros(1, np.uint32),
                             vertices=np.zeros(9, np.float32), indices=np.zeros(3, np.uint32)))
            continue
        blas.append(dict(nodes=np.fromfile(p + "_bvh_nodes.bin", dtype=NODE_DTYPE),
                         prim_ids=np.fromfile(p + "_prim_ids.u32", dtype="<u4"),
                         vertices=np.fromfile(p + "_vertices.f32", dtype="<f4"),
                         indices=np.fromfile(p + "_indices.u32", dtype="<u4")))
        if with_attributes:
            blas[-1].update(normals=np.fromfile(p + "_normals.f32", dtype="<f4"), texcoords=np.fromfile(p + "_texcoords.f32", dtype="<f4"),
                            material_ids=np.fromfile(p + "_material_ids.i32", dtype="<i4"),
                            material_map=np.fromfile(p + "_material_map.u32", dtype="<u4"))
    tlas_nodes = np.fromfile(os.path.join(root, "tlas_nodes.bin"), dtype=NODE_DTYPE)
    tlas_idx = np.fromfile(os.path.join(root, "tlas_indices.u32"), dtype="<u4")
    return tlas_nodes, tlas_idx, blas, man


def load_shading_arrays(root, with_envmap=True):
    """Materials (PackedMaterial records), RGBA8 textures and the RGBA32F environment map of the Chess fixture
    (tools/make_fixtures.py: write_shading_fixture).  Returns (materials, textures, envmap or None)."""
    mats = np.fromfile(os.path.join(root, "materials.bin"), dtype=MATERIAL_DTYPE)
    textures = []
    k = 0
    while os.path.exists(os.path.join(root, "tex%d.rgba8" % k)):
        w, h = np.fromfile(os.path.join(root, "tex%d.dims" % k), dtype="<u4")
        textures.append(np.fromfile(os.path.join(root, "tex%d.rgba8" % k), dtype=np.uint8).reshape(int(h), int(w), 4))
        k += 1
    env = None
    if with_envmap and os.path.exists(os.path.join(root, "envmap.f32")):
        w, h = np.fromfile(os.path.join(root, "envmap.dims"), dtype="<u4")
        env = np.fromfile(os.path.join(root, "envmap.f32"), dtype="<f4").reshape(int(h), int(w), 4)
    return mats, textures, env


def has_shading_fixture(root=None):
    root = root or chess_fixture_dir()
    return root is not None and os.path.exists(os.path.join(root, "materials.bin")) and os.path.exists(os.path.join(root, "envmap.f32"))


def load_chess_module(device=0, max_depth=1, envmap_eval=False, with_neural_mesh=False, root=None, shading=False):
    """Shipped Chess configuration: config JSON, .nbvh and network_models JSON load unchanged through the C ABI.
    shading=True also uploads the shading attributes, materials, textures and the shipped environment map (max_depth > 1)."""
    root = root or chess_fixture_dir()
    if root is None:
        raise FileNotFoundError("Chess fixture missing (baseline/_ref/chess); run tools/make_fixtures.py")
    cfg_text = open(os.path.join(root, CHESS_BASE + "_neural.json")).read()
    mod = NeuralPathModule(cfg_text, device=device)
    mod.load_neural_bvh(os.path.join(root, CHESS_BASE + ".nbvh"))
    mod.load_network_model(os.path.join(root, CHESS_BASE + "_network.json"))
    tlas_nodes, tlas_idx, blas, man = load_scene_arrays(root, with_neural_mesh, with_attributes=shading)
    if shading:
        mats, textures, env = load_shading_arrays(root, with_envmap=envmap_eval)
        mod.set_scene(tlas_nodes, tlas_idx, blas, materials=mats, textures=textures)
        mod.set_envmap(env)
    else:
        mod.set_scene(tlas_nodes, tlas_idx, blas)
    mod.max_depth = max_depth
    from ._capi import CONST_SCENE
    sc = mod.get_constants(CONST_SCENE)
    sc.envmap_eval = envmap_eval
    mod.set_constants(CONST_SCENE, sc)
    return mod


# ------------------------------------------------------------------------------------------------ synthetic
def _build_median_bvh(centers, bmin, bmax, leaf_size):
    """Simple median-split BVH2 in the reference's 32-byte node layout (children adjacent: left, left+1)."""
    n = centers.shape[0]
    order = np.arange(n, dtype=np.uint32)
    nodes = [None]
    prim_ids = []

    def bounds(ids):
        return bmin[ids].min(axis=0), bmax[ids].max(axis=0)

    stack = [(0, order)]
    while stack:
        ni, ids = stack.pop()
        lo, hi = bounds(ids)
        if len(ids) <= leaf_size:
            nodes[ni] = (lo, len(ids), hi, len(prim_ids))
            prim_ids.extend(ids.tolist())
            continue
        axis = int(np.argmax(hi - lo))
        srt = ids[np.argsort(centers[ids, axis], kind="stable")]
        half = len(srt) // 2
        left = len(nodes)
        nodes.append(None)
        nodes.append(None)
        nodes[ni] = (lo, 0, hi, left)
        stack.append((left, srt[:half]))
        stack.append((left + 1, srt[half:]))
    out = np.zeros(len(nodes), dtype=NODE_DTYPE)
    for i, (lo, a, hi, b) in enumerate(nodes):
        out[i] = (lo, a, hi, b)
    return out, np.asarray(prim_ids, dtype=np.uint32)


def synthetic_scene(seed=1234, n_spheres=24, subdiv=12, cut_depth=9, fb=(256, 192), log2_hashmap=14):
    """Procedural stand-in: jittered lattice of tessellated spheres (neural mesh, used only to derive the tree-cut),
    a ground quad grid as the classical BLAS, a fixed-depth N-BVH cut and the shipped network shape.
    Returns (config dict, nbvh nodes, tlas_nodes, tlas_indices, blas_list)."""
    rng = np.random.default_rng(seed)
    # neural mesh: spheres
    verts, tris = [], []
    for s in range(n_spheres):
        c = rng.uniform(-0.4, 0.4, 3) * np.array([1.0, 0.3, 1.0]) + np.array([0, 0.25, 0])
        r = rng.uniform(0.03, 0.08)
        base = len(verts)
        for i in range(subdiv + 1):
            th = np.pi * i / subdiv
            for j in range(2 * subdiv):
                ph = np.pi * j / subdiv
                verts.append(c + r * np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)]))
        w = 2 * subdiv
        for i in range(subdiv):
            for j in range(w):
                a = base + i * w + j
                b = base + i * w + (j + 1) % w
                tris.append((a, b, a + w))
                tris.append((b, b + w, a + w))
    verts = np.asarray(verts, dtype=np.float32)
    tris = np.asarray(tris, dtype=np.uint32)
    tv = verts[tris]
    n_nodes, n_prims = _build_median_bvh(tv.mean(axis=1), tv.min(axis=1), tv.max(axis=1), 2)
    # tree-cut at fixed depth: traverse_and_build_neural_bvh (neural_bvh_renderer_accel.cpp:24-74)
    cut = [None]
    work = [(0, 0, 0)]
    while work:
        bi, ci, depth = work.pop()
        nd = n_nodes[bi]
        if nd["a"] > 0 or depth == cut_depth:
            cut[ci] = (nd["mn"], 1 if nd["a"] > 0 else 8, nd["mx"], 0xFFFFFFFF)
            continue
        left = len(cut)
        cut.append(None)
        cut.append(None)
        cut[ci] = (nd["mn"], 0, nd["mx"], left)
        work.append((int(nd["b"]), left, depth + 1))
        work.append((int(nd["b"]) + 1, left + 1, depth + 1))
    nbvh = np.zeros(len(cut), dtype=NODE_DTYPE)
    for i, c in enumerate(cut):
        nbvh[i] = c
    # classical BLAS: ground grid
    g = 24
    gx, gz = np.meshgrid(np.linspace(-0.6, 0.6, g + 1), np.linspace(-0.6, 0.6, g + 1))
    gverts = np.stack([gx.ravel(), 0.02 * np.sin(7 * gx.ravel()) * np.cos(5 * gz.ravel()), gz.ravel()], axis=1).astype(np.float32)
    gtris = []
    for i in range(g):
        for j in range(g):
            a = i * (g + 1) + j
            gtris.append((a, a + 1, a + g + 1))
            gtris.append((a + 1, a + g + 2, a + g + 1))
    gtris = np.asarray(gtris, dtype=np.uint32)
    gv = gverts[gtris]
    gmin, gmax = gv.min(axis=1), gv.max(axis=1)
    flat = (gmax - gmin) <= 1e-4
    gmin = gmin - flat * 5e-5
    gmax = gmax + flat * 5e-5
    g_nodes, g_prims = _build_median_bvh(gv.mean(axis=1), gmin, gmax, 2)
    blas = [dict(nodes=n_nodes, prim_ids=n_prims, vertices=verts.ravel(), indices=tris.ravel()),
            dict(nodes=g_nodes, prim_ids=g_prims, vertices=gverts.ravel(), indices=gtris.ravel())]
    # TLAS: root + 2 leaves (neural first)
    tlas = np.zeros(3, dtype=NODE_DTYPE)
    tlas[0] = (np.minimum(n_nodes[0]["mn"], g_nodes[0]["mn"]), 1, np.maximum(n_nodes[0]["mx"], g_nodes[0]["mx"]), 0)
    tlas[1] = (n_nodes[0]["mn"], 0, n_nodes[0]["mx"], 2)
    tlas[2] = (g_nodes[0]["mn"], 1, g_nodes[0]["mx"], 1)
    tlas_idx = np.array([0, 1], dtype=np.uint32)
    # camera looking at the origin from above
    eye = np.array([0.9, 0.7, 0.8], dtype=np.float64)
    fwd = -eye / np.linalg.norm(eye)
    right = np.cross(fwd, [0, 1, 0])
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    view = np.eye(4)
    view[0, :3], view[1, :3], view[2, :3] = right, up, -fwd
    view[:3, 3] = -view[:3, :3] @ eye
    cfg = {
        "cuda_scene": {"fb_size": list(fb), "camera": {"yfov": 0.7, "matrix": [float(x) for x in view.T.ravel()]},
                       "scene_constants": {"envmap_eval": False, "ema_accumulate": False}},
        "neural_bvh_renderer": {
            "neural_module": {
                "max_depth": 1,
                "pt_output_constants": {"output_albedo_dim": 3, "output_depth_dim": 1, "output_network_dims": 8,
                                        "output_normal_dim": 3, "output_visibility_dim": 1},
                "bvh_module_constants": {"input_network_dims": 9, "input_segment_points": 3,
                                         "input_segment_points_dim": 9, "segment_distribution": "EquidistantExterior"},
                "segment_encoding": {"type": "HashGrid", "hashgrid_options": {"log2_hashmap_size": log2_hashmap}},
            },
            "bvh_constants": {"rendered_lod": 0, "current_learned_lod": 6},
        },
    }
    return cfg, nbvh, tlas, tlas_idx, blas


# ------------------------------------------------------------------------------------------------ BASELINE config 5
def _icosphere(subdiv):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(subdiv):
        a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
        ab, bc, ca = a + b, b + c, c + a
        n0 = len(v)
        nf = len(f)
        v = np.concatenate([v, ab / np.linalg.norm(ab, axis=1, keepdims=True), bc / np.linalg.norm(bc, axis=1, keepdims=True),
                            ca / np.linalg.norm(ca, axis=1, keepdims=True)])
        iab, ibc, ica = n0 + np.arange(nf), n0 + nf + np.arange(nf), n0 + 2 * nf + np.arange(nf)
        f = np.concatenate([np.stack([f[:, 0], iab, ica], 1), np.stack([f[:, 1], ibc, iab], 1), np.stack([f[:, 2], ica, ibc], 1),
                            np.stack([iab, ibc, ica], 1)])
    return v, f  # (vertices are duplicated along shared edges: irrelevant for bounds)


def synthetic_instanced_scene(lattice=(12, 13, 13), subdiv=4, cut_depth=17, seed=1234, log2_hashmap=18, fb=(1920, 1080), device=None,
                              return_mesh=False):
    """BASELINE.json configs[4]: K = prod(lattice) instances of a displaced, subdivided icosphere "piece" (20 * 4^subdiv triangles
    each; defaults: 2,028 x 5,120 = 10.38 M triangles) on a jittered 3-D lattice inside the unit cube, pre-transformed and
    flattened into ONE neural mesh (the reference supports exactly one, scene.cpp:225-251); std::mt19937-free: numpy PCG64
    with the given seed.  The neural BLAS is a median-split BVH2 over the triangles, cut at a fixed depth (neural_bvh_min_depth,
    neural_bvh_renderer_accel.cpp:24-74): 2^(cut_depth+1) - 1 tree-cut nodes, 2^cut_depth neural leaves.  Only the tree-cut is
    needed for inference, so the build stops at the cut (level-synchronous, vectorised with torch on `device`).
    Returns (config dict, nbvh nodes, tlas_nodes, tlas_indices, blas_list, info)."""
    import torch
    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    rng = np.random.default_rng(seed)
    K = int(np.prod(lattice))
    bv, bf = _icosphere(subdiv)
    # displaced "piece": radial bumps, same for every instance; per-instance scale + jittered lattice position
    disp = 1.0 + 0.25 * np.sin(5.0 * bv[:, 0]) * np.cos(4.0 * bv[:, 1]) + 0.15 * np.sin(7.0 * bv[:, 2])
    piece = (bv * disp[:, None]).astype(np.float32)
    gx, gy, gz = np.meshgrid(*[(np.arange(n) + 0.5) / n for n in lattice], indexing="ij")
    cell = 1.0 / np.array(lattice)
    centers = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], 1) + rng.uniform(-0.2, 0.2, (K, 3)) * cell
    radii = rng.uniform(0.22, 0.32, K) * cell.min()
    tri = torch.from_numpy(piece[bf]).to(dev)                                   # (F, 3, 3)
    c_t = torch.from_numpy(centers.astype(np.float32)).to(dev)
    r_t = torch.from_numpy(radii.astype(np.float32)).to(dev)
    F = tri.shape[0]
    lo = (tri.amin(dim=1)[None] * r_t[:, None, None] + c_t[:, None, :]).reshape(-1, 3)   # (K*F, 3) triangle bounds
    hi = (tri.amax(dim=1)[None] * r_t[:, None, None] + c_t[:, None, :]).reshape(-1, 3)
    cen = (tri.mean(dim=1)[None] * r_t[:, None, None] + c_t[:, None, :]).reshape(-1, 3)
    N = lo.shape[0]
    n_nodes = (1 << (cut_depth + 1)) - 1
    node_lo = torch.empty((n_nodes, 3), dtype=torch.float32, device=dev)
    node_hi = torch.empty((n_nodes, 3), dtype=torch.float32, device=dev)
    seg = torch.zeros(N, dtype=torch.int64, device=dev)  # index of each triangle's node within the current level
    for depth in range(cut_depth + 1):
        n_seg = 1 << depth
        first = n_seg - 1
        idx3 = seg[:, None].expand(-1, 3)
        blo = torch.full((n_seg, 3), float("inf"), device=dev).scatter_reduce_(0, idx3, lo, "amin")
        bhi = torch.full((n_seg, 3), float("-inf"), device=dev).scatter_reduce_(0, idx3, hi, "amax")
        node_lo[first:first + n_seg], node_hi[first:first + n_seg] = blo, bhi
        if depth == cut_depth:
            break
        clo = torch.full((n_seg, 3), float("inf"), device=dev).scatter_reduce_(0, idx3, cen, "amin")
        chi = torch.full((n_seg, 3), float("-inf"), device=dev).scatter_reduce_(0, idx3, cen, "amax")
        axis = (chi - clo).argmax(dim=1)                                         # split axis of each node: longest centroid extent
        key = cen.gather(1, axis[seg][:, None]).squeeze(1)
        order = torch.argsort(key, stable=True)
        order = order[torch.argsort(seg[order], stable=True)]                    # sorted by (node, coordinate)
        count = torch.bincount(seg, minlength=n_seg)
        start = torch.cumsum(count, 0) - count
        rank = torch.empty(N, dtype=torch.int64, device=dev)
        rank[order] = torch.arange(N, device=dev) - start[seg[order]]
        seg = 2 * seg + (rank >= (count[seg] + 1) // 2).to(torch.int64)          # median split: left gets ceil(n/2)
    nbvh = np.zeros(n_nodes, dtype=NODE_DTYPE)
    nbvh["mn"], nbvh["mx"] = node_lo.cpu().numpy(), node_hi.cpu().numpy()
    n_int = (1 << cut_depth) - 1
    nbvh["a"][:n_int] = 0
    nbvh["b"][:n_int] = 2 * np.arange(n_int, dtype=np.uint32) + 1              # children adjacent: (2i+1, 2i+2)
    nbvh["a"][n_int:] = 8                                                       # LOWEST_BVH_LOD: a cut leaf that could still be split
    nbvh["b"][n_int:] = 0xFFFFFFFF
    leaf_count = torch.bincount(seg, minlength=1 << cut_depth)
    stub = dict(nodes=np.zeros(1, NODE_DTYPE), prim_ids=np.zeros(1, np.uint32), vertices=np.zeros(9, np.float32), indices=np.zeros(3, np.uint32))
    tlas = np.zeros(1, dtype=NODE_DTYPE)
    tlas[0] = (nbvh[0]["mn"], 0, nbvh[0]["mx"], 2)                              # a single neural TLAS leaf
    tlas_idx = np.array([0], dtype=np.uint32)
    cfg = {
        "cuda_scene": {"fb_size": list(fb), "camera": {"yfov": 0.7, "matrix": [1.0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, -0.5, -0.5, -2.5, 1]},
                       "scene_constants": {"envmap_eval": False, "ema_accumulate": False}},
        "neural_bvh_renderer": {
            "neural_module": {
                "max_depth": 1,
                "pt_output_constants": {"output_albedo_dim": 3, "output_depth_dim": 1, "output_network_dims": 8,
                                        "output_normal_dim": 3, "output_visibility_dim": 1},
                "bvh_module_constants": {"input_network_dims": 9, "input_segment_points": 3,
                                         "input_segment_points_dim": 9, "segment_distribution": "EquidistantExterior"},
                "segment_encoding": {"type": "HashGrid", "hashgrid_options": {"log2_hashmap_size": log2_hashmap}},
            },
            "bvh_constants": {"rendered_lod": 0, "current_learned_lod": 6},
        },
    }
    info = dict(triangles=int(N), instances=K, triangles_per_instance=int(F), nbvh_nodes=int(n_nodes), neural_leaves=1 << cut_depth,
                triangles_per_leaf_min=int(leaf_count.min()), triangles_per_leaf_max=int(leaf_count.max()))
    if return_mesh:  # the flattened neural mesh itself (device tensors): (K * Vp, 3) float32 vertices, (K * F, 3) int32 indices
        pv = torch.from_numpy(piece).to(dev)
        Vp = pv.shape[0]
        info["mesh_vertices"] = (pv[None] * r_t[:, None, None] + c_t[:, None, :]).reshape(-1, 3).contiguous()
        info["mesh_indices"] = (torch.from_numpy(bf.astype(np.int32)).to(dev)[None] +
                                (torch.arange(K, device=dev, dtype=torch.int32) * Vp)[:, None, None]).reshape(-1, 3).contiguous()
    return cfg, nbvh, tlas, tlas_idx, [stub], info


def synthetic_ray_batch(nbvh_root, n, seed=42, device=None):
    """BASELINE config 5 rays: uniform random origins on the surface of the 1.5x-scaled scene box, directions toward uniformly
    sampled interior points.  Returns float32 (n, 3) origin and (normalised) direction tensors on `device`."""
    import torch
    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    mn = torch.tensor(np.asarray(nbvh_root["mn"], np.float32), device=dev)
    mx = torch.tensor(np.asarray(nbvh_root["mx"], np.float32), device=dev)
    c, e = 0.5 * (mn + mx), 0.5 * (mx - mn)
    u = torch.rand((n, 3), generator=g, device=dev) * 2.0 - 1.0
    face = torch.randint(0, 6, (n,), generator=g, device=dev)                   # equal-area faces of the (near-)cube
    axis, sign = face // 2, (face % 2).to(torch.float32) * 2.0 - 1.0
    u.scatter_(1, axis[:, None], sign[:, None])
    o = c + 1.5 * e * u
    tgt = mn + (mx - mn) * torch.rand((n, 3), generator=g, device=dev)
    d = tgt - o
    d = d / d.norm(dim=1, keepdim=True)
    return o.contiguous(), d.contiguous()
