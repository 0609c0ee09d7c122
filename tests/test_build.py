"""Device-side BLAS build (nbvh_build_blas): an LBVH in the reference's BVH2Node layout.

  * structure: 2n-1 nodes, children adjacent, every interior box is exactly the union of its children's boxes, every leaf holds
    one triangle and encloses it, every triangle appears once, depth within the 32-entry traversal stack
  * function: closest hits through the built tree are bit-identical (t, u, v) to the closest hits through the median-split BVH
    of the same mesh (same triangle test on the same triangles: only the visiting order differs), on the GPU and in the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def validate_tree(nodes, prims, verts, tris):
    n = len(tris)
    assert len(nodes) == 2 * n - 1 and len(prims) == n
    assert sorted(prims.tolist()) == list(range(n))
    interior = nodes["a"] == 0
    assert interior.sum() == n - 1 and (nodes["a"][~interior] == 1).all()
    left = nodes["b"][interior].astype(np.int64)
    assert (left >= 1).all() and (left + 1 < len(nodes)).all()
    kids = np.concatenate([left, left + 1])
    assert len(np.unique(kids)) == 2 * (n - 1) and 0 not in kids  # every non-root node has exactly one parent
    assert (nodes["mn"][interior] == np.minimum(nodes["mn"][left], nodes["mn"][left + 1])).all()
    assert (nodes["mx"][interior] == np.maximum(nodes["mx"][left], nodes["mx"][left + 1])).all()
    leaf_tri = prims[nodes["b"][~interior]]
    tv = verts[tris[leaf_tri]]
    assert (nodes["mn"][~interior] <= tv.min(axis=1)).all() and (nodes["mx"][~interior] >= tv.max(axis=1)).all()


@pytest.mark.parametrize("which", [0, 1])
def test_gpu_lbvh_structure_and_closest_hits(oracle, synthetic, which):
    import torch
    import nbvh_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg, nbvh, tlas, tlas_idx, blas = synthetic
    mesh = blas[which]
    verts = np.asarray(mesh["vertices"], np.float32).reshape(-1, 3)
    tris = np.asarray(mesh["indices"], np.uint32).reshape(-1, 3)
    mod = nbvh_b200.NeuralPathModule(cfg)
    nodes_t, prims_t, depth = mod.build_blas(torch.from_numpy(verts).cuda(), torch.from_numpy(tris.astype(np.int32)).cuda())
    nodes = nodes_t.cpu().numpy().view(nbvh_b200.NODE_DTYPE)
    prims = prims_t.cpu().numpy().astype(np.uint32)
    assert 0 < depth <= 32
    validate_tree(nodes, prims, verts, tris)
    # closest hits through the built tree vs through the median-split tree of the same mesh
    built = dict(nodes=nodes, prim_ids=prims, vertices=mesh["vertices"], indices=mesh["indices"])
    blas2 = list(blas)
    blas2[which] = built
    rng = np.random.default_rng(4)
    n = 20000
    lo, hi = verts.min(axis=0), verts.max(axis=0)
    o = (lo + (hi - lo) * rng.uniform(-0.5, 1.5, (n, 3))).astype(np.float32)
    tgt = (lo + (hi - lo) * rng.uniform(0, 1, (n, 3))).astype(np.float32)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    res = []
    for b in (blas, blas2):
        mod.set_neural_bvh(nbvh)
        mod.set_scene(tlas, tlas_idx, b)
        t, u, v, prim = mod.blas_intersect(which, torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda())
        res.append((t.cpu().numpy(), u.cpu().numpy(), v.cpu().numpy(), prim.cpu().numpy()))
    (t0, u0, v0, p0), (t1, u1, v1, p1) = res
    assert (t0 < 3e38).sum() > n // 10
    assert (t0.view(np.uint32) == t1.view(np.uint32)).all()
    same_prim = p0 == p1
    assert same_prim.mean() > 0.999  # exact ties between two triangles may resolve to either one
    assert (u0.view(np.uint32)[same_prim] == u1.view(np.uint32)[same_prim]).all()
    # the CPU oracle walks the built tree to the same hits
    sc = oracle.Scene(tlas, tlas_idx, blas2)
    tc, uc, vc, pc, _ = oracle.blas_intersect(sc, which, o[:4000], d[:4000])
    assert (tc.view(np.uint32) == t1[:4000].view(np.uint32)).all() and (pc.astype(np.uint32) == p1[:4000].astype(np.uint32)).all()
    mod.close()


def test_gpu_lbvh_degenerate_inputs(synthetic):
    import torch
    import nbvh_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    mod = nbvh_b200.NeuralPathModule(synthetic[0])
    # one triangle; two identical triangles; a flat grid of coincident centroids
    v = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=torch.float32, device="cuda")
    for tris in ([[0, 1, 2]], [[0, 1, 2], [0, 1, 2]], [[0, 1, 2]] * 37):
        idx = torch.tensor(tris, dtype=torch.int32, device="cuda")
        nodes_t, prims_t, depth = mod.build_blas(v, idx)
        nodes = nodes_t.cpu().numpy().view(nbvh_b200.NODE_DTYPE)
        prims = prims_t.cpu().numpy().astype(np.uint32)
        if len(tris) == 1:
            assert nodes["a"][0] == 1 and nodes["b"][0] == 0 and depth == 0
        else:
            validate_tree(nodes, prims, v.cpu().numpy(), np.asarray(tris, np.uint32))
            assert depth <= 8
    mod.close()


# ------------------------------------------------------------------------------------------------ device TLAS build (nbvh_build_tlas)
def _one_triangle_mesh(rng):
    c = rng.uniform(-3, 3, 3)
    v = (c + rng.uniform(-0.4, 0.4, (3, 3))).astype(np.float32)
    import nbvh_b200
    nodes = np.zeros(1, nbvh_b200.NODE_DTYPE)
    nodes[0] = (v.min(axis=0), 1, v.max(axis=0), 0)
    return dict(nodes=nodes, prim_ids=np.zeros(1, np.uint32), vertices=v.reshape(-1), indices=np.arange(3, dtype=np.uint32))


@pytest.mark.parametrize("n_meshes", [1, 2, 5, 16, 40, 300])
def test_gpu_tlas_structure(synthetic, n_meshes):
    """TLASNode layout (tlas.cuh:64-68): 2n-1 nodes, adjacent children, every mesh in exactly one leaf whose box is the mesh's root box,
    interior boxes are the union of their children, neural meshes carry leaf type 2."""
    import torch
    import nbvh_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    rng = np.random.default_rng(n_meshes)
    meshes = [_one_triangle_mesh(rng) for _ in range(n_meshes)]
    neural = [0] if n_meshes < 5 else [1, n_meshes - 2]
    mod = nbvh_b200.NeuralPathModule(synthetic[0])
    mod.set_scene(None, None, meshes, neural_meshes=neural)
    nodes, idx = mod.tlas_arrays()
    assert len(nodes) == 2 * n_meshes - 1 and sorted(idx.tolist()) == list(range(n_meshes))
    leaf = nodes["b"] >= 1
    assert leaf.sum() == n_meshes
    mesh_of_leaf = idx[nodes["a"][leaf]]
    assert sorted(mesh_of_leaf.tolist()) == list(range(n_meshes))
    for node, m in zip(nodes[leaf], mesh_of_leaf):
        assert (node["mn"] == meshes[m]["nodes"]["mn"][0]).all() and (node["mx"] == meshes[m]["nodes"]["mx"][0]).all()
        assert node["b"] == (2 if m in neural else 1)
    left = nodes["a"][~leaf].astype(np.int64)
    kids = np.concatenate([left, left + 1])
    assert len(np.unique(kids)) == 2 * (n_meshes - 1) and 0 not in kids and (kids < len(nodes)).all()
    assert (nodes["mn"][~leaf] == np.minimum(nodes["mn"][left], nodes["mn"][left + 1])).all()
    assert (nodes["mx"][~leaf] == np.maximum(nodes["mx"][left], nodes["mx"][left + 1])).all()
    with pytest.raises(nbvh_b200.NbvhError):
        mod.set_scene(None, None, meshes, neural_meshes=[n_meshes])   # neural mesh index out of range
    mod.close()


def test_gpu_tlas_renders_like_host_tlas(synthetic):
    """The frame through the device-built TLAS is bit-identical (hit records and accumulation) to the frame through the host-built one."""
    import torch
    import nbvh_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg, nbvh, tlas, tlas_idx, blas = synthetic
    w, h = cfg["cuda_scene"]["fb_size"]
    out = []
    for device_tlas in (False, True):
        mod = nbvh_b200.NeuralPathModule(cfg)
        mod.set_neural_bvh(nbvh)
        if device_tlas:
            mod.set_scene(None, None, blas, neural_meshes=[int(tlas_idx[tlas["a"][tlas["b"] == 2][0]])])
            assert len(mod.tlas_arrays()[0]) == len(tlas)
        else:
            mod.set_scene(tlas, tlas_idx, blas)
        mod.init_params(1337)
        accum, _ = mod.attach_framebuffer((w, h))
        hits = mod.attach_hit_buffer(w * h)
        mod.neural_render(0, 0)
        torch.cuda.synchronize()
        out.append((accum.cpu().numpy().copy(), hits.cpu().numpy().copy()))
        mod.close()
    assert out[0][0].std() > 0
    assert (out[0][0].view(np.uint32) == out[1][0].view(np.uint32)).all()
    assert (out[0][1] == out[1][1]).all()


def test_gpu_scene_with_more_than_16_meshes(synthetic):
    """The reference's TLAS has no bound on the mesh count (cuda_backend.cpp:312-385).  The classical mesh of the synthetic scene is cut
    into 40 meshes with their own device-built BLAS (41 meshes with the neural one: the descriptors no longer fit the kernel
    parameters and travel through the device table, TLAS built on the device): the closest hits -- kind, t, and for neural hits normal
    and albedo -- are bit-identical to the frame of the 2-mesh scene; classical hits name another (mesh, primitive) pair, of course."""
    import torch
    import nbvh_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg, nbvh, tlas, tlas_idx, blas = synthetic
    w, h = cfg["cuda_scene"]["fb_size"]
    neural_idx = int(tlas_idx[tlas["a"][tlas["b"] == 2][0]])
    classical_idx = 1 - neural_idx
    assert len(blas) == 2

    def frame(meshes, t_nodes, t_idx, neural):
        mod = nbvh_b200.NeuralPathModule(cfg)
        mod.set_neural_bvh(nbvh)
        if t_nodes is None:
            mod.set_scene(None, None, meshes, neural_meshes=neural)
        else:
            mod.set_scene(t_nodes, t_idx, meshes)
        mod.init_params(1337)
        hits = mod.attach_hit_buffer(w * h)
        mod.neural_render(0, 0)
        torch.cuda.synchronize()
        out = hits.cpu().numpy().view(nbvh_b200.HIT_DTYPE).copy()
        mod.close()
        return out

    ref = frame(blas, tlas, tlas_idx, None)
    # cut the classical mesh into 40 index ranges, one device-built BLAS each
    c = blas[classical_idx]
    verts = np.ascontiguousarray(c["vertices"], dtype=np.float32).reshape(-1, 3)
    tris = np.ascontiguousarray(c["indices"], dtype=np.uint32).reshape(-1, 3)
    builder = nbvh_b200.NeuralPathModule(cfg)
    parts = []
    for chunk in np.array_split(np.arange(len(tris)), 40):
        v_t = torch.from_numpy(verts).cuda()
        i_t = torch.from_numpy(tris[chunk].astype(np.int32)).cuda()
        nodes, prim, _ = builder.build_blas(v_t, i_t)
        torch.cuda.synchronize()
        parts.append(dict(nodes=nodes.cpu().numpy().view(nbvh_b200.NODE_DTYPE).copy(), prim_ids=prim.cpu().numpy().astype(np.uint32),
                          vertices=verts.reshape(-1), indices=tris[chunk].reshape(-1)))
    builder.close()
    meshes = parts[:7] + [blas[neural_idx]] + parts[7:]
    got = frame(meshes, None, None, [7])
    assert (ref["valid_hit"] == 2).sum() > 500 and (ref["valid_hit"] == 1).sum() > 500
    assert (got["valid_hit"] == ref["valid_hit"]).all()
    assert (got["t"].view(np.uint32) == ref["t"].view(np.uint32)).all()
    neu = ref["valid_hit"] == 1
    assert got[neu].tobytes() == ref[neu].tobytes()
    cls = ref["valid_hit"] == 2
    # barycentrics of the same triangle (a ray through a shared edge may report either neighbour: same t, other u / v)
    assert (got["normal_or_uv"][cls].view(np.uint32) == ref["normal_or_uv"][cls].view(np.uint32)).all(axis=1).mean() > 0.999
