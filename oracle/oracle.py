"""ctypes front-end of the CPU oracle (oracle/libnbvh_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (nbvh_b200) never imports this module.

Also holds independent numpy readers for the reference's on-disk formats (config JSON, .nbvh,
network_models JSON) so the C++ loaders of the product can be checked against a second implementation.
"""
import ctypes as C
import json
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnbvh_oracle.so")

NODE_DTYPE = np.dtype([("mn", "<f4", 3), ("a", "<u4"), ("mx", "<f4", 3), ("b", "<u4")])
HIT_DTYPE = np.dtype([("valid_hit", "u1"), ("t", "<f4"), ("normal_or_uv", "<f4", 3), ("albedo_or_ids", "<f4", 3)],
                     align=True)


class ONode(C.Structure):
    _fields_ = [("mn", C.c_float * 3), ("a", C.c_uint32), ("mx", C.c_float * 3), ("b", C.c_uint32)]


class OBlas(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("n_nodes", C.c_uint32), ("prim_ids", C.c_void_p), ("vertices", C.c_void_p),
                ("indices", C.c_void_p)]


class OScene(C.Structure):
    _fields_ = [("tlas_nodes", C.c_void_p), ("n_tlas_nodes", C.c_uint32), ("blas_indices", C.c_void_p),
                ("blas", C.POINTER(OBlas)), ("n_blas", C.c_uint32)]


class OConst(C.Structure):
    _fields_ = [("rendered_lod", C.c_uint32), ("fixed_node_dilation", C.c_float), ("node_extent_dilation", C.c_float),
                ("inference_mint_epsilon", C.c_float), ("inference_maxt_epsilon", C.c_float),
                ("learning_mint_epsilon", C.c_float), ("learning_maxt_epsilon", C.c_float),
                ("primary_visibility_threshold", C.c_float), ("secondary_visibility_threshold", C.c_float),
                ("depth_epsilon", C.c_float), ("fixed_depth_epsilon", C.c_float), ("ignore_neural_blas", C.c_uint32),
                ("n_segment_points", C.c_uint32), ("segment_distribution", C.c_uint32), ("bbox_min", C.c_float * 3),
                ("bbox_max", C.c_float * 3), ("cam_origin", C.c_float * 3), ("cam_du", C.c_float * 3),
                ("cam_dv", C.c_float * 3), ("cam_top_left", C.c_float * 3), ("fb_w", C.c_uint32), ("fb_h", C.c_uint32)]


class ONet(C.Structure):
    _fields_ = [("n_levels", C.c_uint32), ("n_features", C.c_uint32), ("base_resolution", C.c_uint32),
                ("log2_hashmap_size", C.c_uint32), ("per_level_scale", C.c_float), ("n_hidden_layers", C.c_uint32),
                ("n_neurons", C.c_uint32), ("n_input", C.c_uint32), ("n_output_padded", C.c_uint32),
                ("n_output", C.c_uint32), ("params", C.c_void_p), ("n_params", C.c_uint64)]


_lib = None


def build():
    subprocess.check_call(["make", "-C", HERE, "-s"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.onbvh_h2f.restype = C.c_float
        _lib.onbvh_h2f.argtypes = [C.c_uint16]
        _lib.onbvh_f2h.restype = C.c_uint16
        _lib.onbvh_f2h.argtypes = [C.c_float]
        _lib.onbvh_lcg_rng_init.restype = C.c_uint32
        _lib.onbvh_lcg_randomf.restype = C.c_float
        _lib.onbvh_blas_intersect.restype = C.c_float
        _lib.onbvh_grid_n_params.restype = C.c_uint64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------- file formats (independent readers)
def read_nbvh(path):
    """.nbvh = bvh::v2::Bvh<Node<float,3,64,32>>::serialize (ext/bvh/src/bvh/v2/bvh.h:133-153, node.h:119-133).
    Returns 32-byte device nodes laid out like neural_bvh_renderer_accel.cpp:229-238 does."""
    d = open(path, "rb").read()
    n_nodes, n_prims = struct.unpack("<QQ", d[:16])
    rec = np.dtype([("b", "<f4", 6), ("first", "<u8"), ("cnt", "<u8")])
    assert len(d) == 16 + n_nodes * rec.itemsize + 8 * n_prims, "unexpected .nbvh size"
    raw = np.frombuffer(d, dtype=rec, count=n_nodes, offset=16)
    out = np.zeros(n_nodes, dtype=NODE_DTYPE)
    out["mn"] = raw["b"][:, 0::2]
    out["mx"] = raw["b"][:, 1::2]
    out["a"] = raw["cnt"].astype(np.uint32)      # prim_count (LoD tag)
    out["b"] = raw["first"].astype(np.uint32)    # left child / NO_NEURAL_CHILD
    return out


def read_network_json(path):
    """network_models/*_network.json (trainer.h:275-320, gpu_memory_json.h:36-60): returns fp16 bit patterns."""
    j = json.load(open(path))
    n = int(j["n_params"])
    raw = np.asarray(j["params_binary"]["bytes"], dtype=np.uint8)
    if j["params_type"] == "__half":
        bits = raw.view("<u2").copy()
    elif j["params_type"] == "float":
        bits = raw.view("<f4").astype(np.float16).view("<u2").copy()
    else:
        raise ValueError("unknown params_type " + str(j["params_type"]))
    assert bits.size == n
    return bits


SEGMENT_DISTRIBUTIONS = ["EquidistantInterior", "EquidistantExterior", "Uniform"]


def config_constants(cfg, nbvh_nodes=None):
    """Config JSON -> OConst + network description, with the reference's defaults
    (neural_bvh_renderer_constants.cuh:13-47, neural_bvh_module.h:79-94,243-274, tcnn_config.h)."""
    scene = cfg.get("cuda_scene", {})
    r = cfg.get("neural_bvh_renderer", {})
    m = r.get("neural_module", {})
    bc = r.get("bvh_constants", {})
    pt = m.get("pt_output_constants", {})
    mc = m.get("bvh_module_constants", {})
    c = OConst()
    c.rendered_lod = bc.get("rendered_lod", 0)
    c.fixed_node_dilation = bc.get("fixed_node_dilation", 1e-5)
    c.node_extent_dilation = bc.get("node_extent_dilation", 1e-4)
    c.inference_mint_epsilon = bc.get("inference_mint_epsilon", 1e-4)
    c.inference_maxt_epsilon = bc.get("inference_maxt_epsilon", 1e-3)
    c.learning_mint_epsilon = bc.get("learning_mint_epsilon", 1e-3)
    c.learning_maxt_epsilon = bc.get("learning_maxt_epsilon", 1e-3)
    c.primary_visibility_threshold = pt.get("primary_visibility_threshold", 0.5)
    c.secondary_visibility_threshold = pt.get("secondary_visibility_threshold", 0.5)
    c.depth_epsilon = pt.get("depth_epsilon", 0.2)
    c.fixed_depth_epsilon = pt.get("fixed_depth_epsilon", 1e-5)
    c.ignore_neural_blas = int(bc.get("ignore_neural_blas", False))
    c.n_segment_points = mc.get("input_segment_points", 3)
    c.segment_distribution = SEGMENT_DISTRIBUTIONS.index(mc.get("segment_distribution", "EquidistantInterior"))
    fb = scene.get("fb_size", [1920, 1080])
    c.fb_w, c.fb_h = fb
    cam = scene.get("camera", {})
    if "matrix" in cam:
        mtx = (C.c_float * 16)(*[float(x) for x in cam["matrix"]])
        o, du, dv, tl = (C.c_float * 3)(), (C.c_float * 3)(), (C.c_float * 3)(), (C.c_float * 3)()
        lib().onbvh_camera_params(mtx, C.c_float(cam.get("yfov", np.radians(45.0))), C.c_uint32(c.fb_w),
                                  C.c_uint32(c.fb_h), o, du, dv, tl)
        c.cam_origin, c.cam_du, c.cam_dv, c.cam_top_left = o, du, dv, tl
    if nbvh_nodes is not None:
        c.bbox_min = (C.c_float * 3)(*nbvh_nodes["mn"][0])
        c.bbox_max = (C.c_float * 3)(*nbvh_nodes["mx"][0])
    return c


def make_net(cfg, params_bits):
    m = cfg.get("neural_bvh_renderer", {}).get("neural_module", {})
    hg = m.get("segment_encoding", {}).get("hashgrid_options", {})
    nw = m.get("network", {})
    pt = m.get("pt_output_constants", {})
    mc = m.get("bvh_module_constants", {})
    net = ONet()
    net.n_levels = hg.get("n_levels", 8)
    net.n_features = hg.get("n_features_per_level", 4)
    net.base_resolution = hg.get("base_resolution", 8)
    net.log2_hashmap_size = hg.get("log2_hashmap_size", 18)
    net.per_level_scale = hg.get("per_level_scale", 2.0)
    net.n_hidden_layers = nw.get("n_hidden_layers", 4)
    net.n_neurons = nw.get("n_neurons", 64)
    n_points = mc.get("input_segment_points", 3)
    net.n_input = n_points * net.n_levels * net.n_features
    # derived like NeuralPathModule::update_neural_model_module (path_module_init.cu:259-265): the sum of the four dims with the
    # struct defaults 1, 1, 0, 0 (neural_bvh_module.h:243-248); the file's own "output_network_dims" is ignored by the reference
    net.n_output = (pt.get("output_visibility_dim", 1) + pt.get("output_depth_dim", 1) + pt.get("output_normal_dim", 0) +
                    pt.get("output_albedo_dim", 0))
    net.n_output_padded = (net.n_output + 15) // 16 * 16
    params_bits = np.ascontiguousarray(params_bits, dtype=np.uint16)
    net.params = params_bits.ctypes.data
    net.n_params = params_bits.size
    net._keep = params_bits
    return net


def expected_n_params(net):
    offs = (C.c_uint32 * 65)()
    grid = lib().onbvh_grid_n_params(C.byref(net), offs)
    mlp = net.n_neurons * net.n_input + (net.n_hidden_layers - 1) * net.n_neurons ** 2 + net.n_output_padded * net.n_neurons
    return mlp + grid, list(offs[: net.n_levels + 1])


# ----------------------------------------------------------------------------- scene assembly
class Scene:
    """Keeps numpy arrays alive and exposes an OScene."""

    def __init__(self, tlas_nodes, tlas_indices, blas_list):
        self.tlas_nodes = np.ascontiguousarray(tlas_nodes, dtype=NODE_DTYPE)
        self.tlas_indices = np.ascontiguousarray(tlas_indices, dtype=np.uint32)
        self.blas_arrays = []
        self.blas = (OBlas * len(blas_list))()
        for i, b in enumerate(blas_list):
            nodes = np.ascontiguousarray(b["nodes"], dtype=NODE_DTYPE)
            prim = np.ascontiguousarray(b["prim_ids"], dtype=np.uint32)
            verts = np.ascontiguousarray(b["vertices"], dtype=np.float32)
            idx = np.ascontiguousarray(b["indices"], dtype=np.uint32)
            self.blas_arrays.append((nodes, prim, verts, idx))
            self.blas[i] = OBlas(_p(nodes), nodes.size, _p(prim), _p(verts), _p(idx))
        self.c = OScene(_p(self.tlas_nodes), self.tlas_nodes.size, _p(self.tlas_indices), self.blas, len(blas_list))


def load_chess_fixture(root, with_neural_mesh=False):
    """Loads baseline/_ref/chess (written by tools/make_fixtures.py).  The neural mesh's triangle BLAS
    (1.4 M triangles) is only needed for training targets / the non-neural CPU baseline."""
    man = json.load(open(os.path.join(root, "manifest.json")))
    blas = []
    for i, m in enumerate(man["meshes"]):
        p = os.path.join(root, "mesh%d" % i)
        if i == man["neural_mesh"] and not with_neural_mesh:
            # a stub: never traversed while ignore_neural_blas == false
            blas.append(dict(nodes=np.zeros(1, NODE_DTYPE), prim_ids=np.zeros(1, np.uint32),
                             vertices=np.zeros(9, np.float32), indices=np.zeros(3, np.uint32)))
            continue
        blas.append(dict(nodes=np.fromfile(p + "_bvh_nodes.bin", dtype=NODE_DTYPE),
                         prim_ids=np.fromfile(p + "_prim_ids.u32", dtype="<u4"),
                         vertices=np.fromfile(p + "_vertices.f32", dtype="<f4"),
                         indices=np.fromfile(p + "_indices.u32", dtype="<u4")))
    tlas_nodes = np.fromfile(os.path.join(root, "tlas_nodes.bin"), dtype=NODE_DTYPE)
    tlas_idx = np.fromfile(os.path.join(root, "tlas_indices.u32"), dtype="<u4")
    return Scene(tlas_nodes, tlas_idx, blas), man


# ----------------------------------------------------------------------------- wrappers
def trace_intervals(nodes, const, o, d, max_intervals):
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
    o = np.ascontiguousarray(o, dtype=np.float32)
    d = np.ascontiguousarray(d, dtype=np.float32)
    n = o.shape[0]
    counts = np.zeros(n, np.uint32)
    node_idx = np.zeros((n, max_intervals), np.uint32)
    mint = np.zeros((n, max_intervals), np.float32)
    maxt = np.zeros((n, max_intervals), np.float32)
    steps = C.c_uint64(0)
    lib().onbvh_trace_intervals(_p(nodes), C.byref(const), C.c_uint32(n), _p(o), _p(d), C.c_uint32(max_intervals),
                                _p(counts), _p(node_idx), _p(mint), _p(maxt), C.byref(steps))
    return counts, node_idx, mint, maxt, steps.value


def raygen(const, pixels, sample_idx=0, rnd_offset=0):
    pixels = np.asarray(pixels, dtype=np.uint32)
    o = np.zeros((pixels.size, 3), np.float32)
    d = np.zeros((pixels.size, 3), np.float32)
    oo, dd = (C.c_float * 3)(), (C.c_float * 3)()
    f = lib().onbvh_raygen
    for i, p in enumerate(pixels):
        f(C.byref(const), C.c_uint32(int(p)), C.c_int(sample_idx), C.c_int(rnd_offset), oo, dd, None)
        o[i] = oo[:]
        d[i] = dd[:]
    return o, d


def blas_intersect(scene, blas_idx, o, d, maxt=None):
    o = np.ascontiguousarray(o, dtype=np.float32)
    d = np.ascontiguousarray(d, dtype=np.float32)
    n = o.shape[0]
    t = np.zeros(n, np.float32)
    u = np.zeros(n, np.float32)
    v = np.zeros(n, np.float32)
    prim = np.zeros(n, np.uint32)
    cnt = (C.c_uint64 * 2)()
    mt = None if maxt is None else _p(np.ascontiguousarray(maxt, dtype=np.float32))
    lib().onbvh_blas_intersect_batch(C.byref(scene.blas[blas_idx]), C.c_uint32(n), _p(o), _p(d), mt, _p(t), _p(u),
                                     _p(v), _p(prim), cnt)
    return t, u, v, prim, (cnt[0], cnt[1])


def encode_segments(const, o, d, mint, maxt):
    n = o.shape[0]
    out = np.zeros((n, 3 * const.n_segment_points), np.float32)
    f = lib().onbvh_encode_segment
    for i in range(n):
        oo = (C.c_float * 3)(*o[i])
        dd = (C.c_float * 3)(*d[i])
        f(C.byref(const), oo, dd, C.c_float(mint[i]), C.c_float(maxt[i]), _p(out[i]))
    return out


def raygen_rng(const, pixels, sample_idx=0, rnd_offset=0):
    """State of each pixel's generator after ray generation (what the wavefront's traversal data carries)."""
    pixels = np.asarray(pixels, dtype=np.uint32)
    out = np.zeros(pixels.size, np.uint32)
    oo, dd, st = (C.c_float * 3)(), (C.c_float * 3)(), C.c_uint32(0)
    f = lib().onbvh_raygen
    for i, p in enumerate(pixels):
        f(C.byref(const), C.c_uint32(int(p)), C.c_int(sample_idx), C.c_int(rnd_offset), oo, dd, C.byref(st))
        out[i] = st.value
    return out


def encode_segments_rng(const, o, d, mint, maxt, rng_states):
    """encode_segments with each ray's generator (SegmentDistribution::Uniform); returns (inputs, advanced generator states)."""
    n = o.shape[0]
    out = np.zeros((n, 3 * const.n_segment_points), np.float32)
    states = np.array(rng_states, dtype=np.uint32).copy()
    f = lib().onbvh_encode_segment_rng
    for i in range(n):
        oo = (C.c_float * 3)(*o[i])
        dd = (C.c_float * 3)(*d[i])
        st = C.c_uint32(int(states[i]))
        f(C.byref(const), oo, dd, C.c_float(mint[i]), C.c_float(maxt[i]), C.byref(st), _p(out[i]))
        states[i] = st.value
    return out, states


def query(net, inputs, mode=0):
    """inputs: (N, 9) float32 rows (== the reference's column-major 9xN matrix in memory)."""
    x = np.ascontiguousarray(inputs, dtype=np.float32)
    out = np.zeros((x.shape[0], net.n_output), np.float32)
    lib().onbvh_query(C.byref(net), _p(x), C.c_uint32(x.shape[0]), C.c_int(mode), _p(out))
    return out


def grid_encode(net, points):
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
    feats = np.zeros(pts.shape[0] * net.n_levels * net.n_features, np.uint16)
    lib().onbvh_grid_encode(C.byref(net), _p(pts), C.c_uint32(pts.shape[0]), _p(feats))
    return feats


def render_primary(scene, nbvh_nodes, const, net, pixel_begin, pixel_end, stride=1, sample_idx=0, rnd_offset=0,
                   mlp_mode=0, n_threads=0):
    nbvh_nodes = np.ascontiguousarray(nbvh_nodes, dtype=NODE_DTYPE)
    n = (pixel_end - pixel_begin + stride - 1) // stride
    hits = np.zeros(n, dtype=HIT_DTYPE)
    cnt = (C.c_uint64 * 4)()
    lib().onbvh_render_primary(C.byref(scene.c), _p(nbvh_nodes), C.byref(const), C.byref(net), C.c_uint32(pixel_begin),
                               C.c_uint32(pixel_end), C.c_uint32(stride), C.c_int(sample_idx), C.c_int(rnd_offset),
                               C.c_int(mlp_mode), _p(hits), cnt, C.c_int(n_threads))
    return hits, [cnt[i] for i in range(4)]


# ----------------------------------------------------------------------------- sample generator (a17, K10)
class OTexture(C.Structure):
    _fields_ = [("rgba", C.c_void_p), ("w", C.c_uint32), ("h", C.c_uint32)]


class OSampler(C.Structure):
    _fields_ = [("blas", OBlas), ("normals", C.c_void_p), ("texcoords", C.c_void_p), ("material_ids", C.c_void_p),
                ("material_map", C.c_void_p), ("num_materials", C.c_uint32), ("materials", C.c_void_p),
                ("textures", C.POINTER(OTexture)), ("sampling_bbox_scaling", C.c_float), ("current_learned_lod", C.c_uint32),
                ("training_node_survival", C.c_uint32), ("survival_probability_scale", C.c_float),
                ("min_rr_survival_prob", C.c_float), ("max_training_bounces", C.c_int32),
                ("backface_bounce_training", C.c_uint32), ("sample_cosine_weighted", C.c_uint32),
                ("depth_linear_activation", C.c_uint32)]


class Sampler:
    """Keeps the arrays of the neural mesh (classical BLAS + shading attributes), materials and textures alive."""

    def __init__(self, cfg, blas, materials=None, textures=None):
        bc = cfg.get("neural_bvh_renderer", {}).get("bvh_constants", {})
        pt = cfg.get("neural_bvh_renderer", {}).get("neural_module", {}).get("pt_output_constants", {})
        self.keep = [np.ascontiguousarray(blas["nodes"], dtype=NODE_DTYPE), np.ascontiguousarray(blas["prim_ids"], dtype=np.uint32),
                     np.ascontiguousarray(blas["vertices"], dtype=np.float32), np.ascontiguousarray(blas["indices"], dtype=np.uint32)]
        s = OSampler()
        s.blas = OBlas(_p(self.keep[0]), self.keep[0].size, _p(self.keep[1]), _p(self.keep[2]), _p(self.keep[3]))
        for key, dt in (("normals", np.float32), ("texcoords", np.float32), ("material_ids", np.int32), ("material_map", np.uint32)):
            if blas.get(key) is not None:
                a = np.ascontiguousarray(blas[key], dtype=dt)
                self.keep.append(a)
                setattr(s, key, a.ctypes.data)
        s.num_materials = len(blas["material_map"]) if blas.get("material_map") is not None else 0
        if materials is not None:
            m = np.ascontiguousarray(materials)
            assert m.dtype.itemsize == 96
            self.keep.append(m)
            s.materials = m.ctypes.data
        if textures:
            self.tex = (OTexture * len(textures))()
            for k, img in enumerate(textures):
                img = np.ascontiguousarray(img, dtype=np.uint8)
                self.keep.append(img)
                self.tex[k] = OTexture(img.ctypes.data, img.shape[1], img.shape[0])
            s.textures = self.tex
        s.sampling_bbox_scaling = bc.get("sampling_bbox_scaling", 1.5)
        s.current_learned_lod = bc.get("current_learned_lod", 0)
        s.training_node_survival = int(bc.get("training_node_survival", True))
        s.survival_probability_scale = bc.get("survival_probability_scale", 1.0)
        s.min_rr_survival_prob = bc.get("min_rr_survival_prob", 0.005)
        s.max_training_bounces = bc.get("max_training_bounces", 1)
        s.backface_bounce_training = int(bc.get("backface_bounce_training", False))
        s.sample_cosine_weighted = int(bc.get("sample_cosine_weighted", True))
        s.depth_linear_activation = int(pt.get("output_depth_linear_activation", False))
        self.c = s


def generate_batch(nbvh_nodes, const, sampler, training_step, batch, log2_max_trav, sampled_nodes, max_loss, tree_cut, inputs, targets):
    """In place on sampled_nodes (uint32 [B]), inputs (float32 [B, 9]), targets (float32 [B, 8])."""
    nbvh_nodes = np.ascontiguousarray(nbvh_nodes, dtype=NODE_DTYPE)
    tc = np.ascontiguousarray(tree_cut)
    assert tc.dtype.itemsize == 16 and inputs.flags.c_contiguous and targets.flags.c_contiguous
    lib().onbvh_generate_batch(_p(nbvh_nodes), C.byref(const), C.byref(sampler.c), C.c_uint32(training_step), C.c_uint32(batch),
                               C.c_uint32(log2_max_trav), _p(sampled_nodes), C.c_float(max_loss), _p(tc), _p(inputs), _p(targets))


def tex2d(img, u, v):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    t = OTexture(img.ctypes.data, img.shape[1], img.shape[0])
    out = (C.c_float * 4)()
    lib().onbvh_tex2d(C.byref(t), C.c_float(u), C.c_float(v), out)
    return np.array(out[:], dtype=np.float32)


def trace_rays(scene, nbvh_nodes, const, net, o, d, ray_depth=1, mlp_mode=0, n_threads=0):
    nbvh_nodes = np.ascontiguousarray(nbvh_nodes, dtype=NODE_DTYPE)
    o = np.ascontiguousarray(o, dtype=np.float32)
    d = np.ascontiguousarray(d, dtype=np.float32)
    n = o.shape[0]
    hits = np.zeros(n, dtype=HIT_DTYPE)
    cnt = (C.c_uint64 * 4)()
    lib().onbvh_trace_rays(C.byref(scene.c), _p(nbvh_nodes), C.byref(const), C.byref(net), C.c_uint32(n), _p(o), _p(d),
                           C.c_uint32(ray_depth), C.c_uint32(const.rendered_lod), C.c_int(mlp_mode), _p(hits), cnt, C.c_int(n_threads))
    return hits, [cnt[i] for i in range(4)]


# ----------------------------------------------------------------------------- hybrid neural path tracing (a15 / f1)
class OAttr(C.Structure):
    _fields_ = [("normals", C.c_void_p), ("texcoords", C.c_void_p), ("material_ids", C.c_void_p), ("material_map", C.c_void_p),
                ("num_materials", C.c_uint32)]


class OPath(C.Structure):
    _fields_ = [("attr", C.POINTER(OAttr)), ("materials", C.c_void_p), ("textures", C.POINTER(OTexture)),
                ("neural_bvh_idx", C.c_uint32), ("max_depth", C.c_uint32), ("path_lod_switch", C.c_uint32),
                ("lod_switch", C.c_uint32), ("num_lods_learned", C.c_int32), ("normal_epsilon_offset", C.c_float),
                ("apply_russian_roulette", C.c_uint32), ("rr_start_depth", C.c_uint32), ("neural_roughness", C.c_float),
                ("neural_specular", C.c_float), ("neural_metallic", C.c_float), ("force_constant_material", C.c_uint32),
                ("emission_eval", C.c_uint32), ("constant_roughness", C.c_float), ("constant_specular", C.c_float),
                ("constant_metallic", C.c_float), ("envmap_eval", C.c_uint32), ("skylike_background", C.c_uint32),
                ("envmap_rotation_deg", C.c_float), ("envmap_strength", C.c_float), ("envmap", C.c_void_p),
                ("env_w", C.c_uint32), ("env_h", C.c_uint32)]


class PathSetup:
    """Shading inputs of the path tracer: per-BLAS attributes (dicts with optional normals / texcoords / material_ids /
    material_map), PackedMaterial records (96-byte numpy records), RGBA8 textures, an RGBA32F environment map, and the
    path / scene constants of the config JSON."""

    def __init__(self, cfg, blas_list, materials=None, textures=None, envmap=None, max_depth=None, neural_bvh_idx=0):
        r = cfg.get("neural_bvh_renderer", {})
        nm = r.get("neural_module", {})
        pmc = nm.get("path_module_constants", {})
        bc = r.get("bvh_constants", {})
        sc = cfg.get("cuda_scene", {}).get("scene_constants", {})
        self.keep = []
        self.attr = (OAttr * len(blas_list))()
        for i, b in enumerate(blas_list):
            a = OAttr()
            for key, dt in (("normals", np.float32), ("texcoords", np.float32), ("material_ids", np.int32), ("material_map", np.uint32)):
                if b.get(key) is not None and len(b[key]):
                    arr = np.ascontiguousarray(b[key], dtype=dt)
                    self.keep.append(arr)
                    setattr(a, key, arr.ctypes.data)
            a.num_materials = len(b["material_map"]) if b.get("material_map") is not None else 0
            self.attr[i] = a
        p = OPath()
        p.attr = self.attr
        if materials is not None and len(materials):
            m = np.ascontiguousarray(materials)
            assert m.dtype.itemsize == 96
            self.keep.append(m)
            p.materials = m.ctypes.data
        if textures:
            self.tex = (OTexture * len(textures))()
            for k, img in enumerate(textures):
                img = np.ascontiguousarray(img, dtype=np.uint8)
                self.keep.append(img)
                self.tex[k] = OTexture(img.ctypes.data, img.shape[1], img.shape[0])
            p.textures = self.tex
        if envmap is not None:
            e = np.ascontiguousarray(envmap, dtype=np.float32)
            assert e.ndim == 3 and e.shape[2] == 4
            self.keep.append(e)
            p.envmap, p.env_w, p.env_h = e.ctypes.data, e.shape[1], e.shape[0]
        p.neural_bvh_idx = neural_bvh_idx
        p.max_depth = int(max_depth if max_depth is not None else nm.get("max_depth", 2))
        p.path_lod_switch = int(bc.get("path_lod_switch", False))
        p.lod_switch = int(bc.get("lod_switch", 0))
        p.num_lods_learned = int(bc.get("num_lods_learned", 0))
        p.normal_epsilon_offset = pmc.get("normal_epsilon_offset", 1e-4)
        p.apply_russian_roulette = int(pmc.get("apply_russian_roulette", True))
        p.rr_start_depth = int(pmc.get("rr_start_depth", 2))
        p.neural_roughness = pmc.get("neural_roughness", 0.2)
        p.neural_specular = pmc.get("neural_specular", 0.2)
        p.neural_metallic = pmc.get("neural_metallic", 0.1)
        p.force_constant_material = int(sc.get("force_constant_material", False))
        p.emission_eval = int(sc.get("emission_eval", True))
        p.constant_roughness = sc.get("constant_roughness", 1.0)
        p.constant_specular = sc.get("constant_specular", 0.0)
        p.constant_metallic = sc.get("constant_metallic", 0.0)
        p.envmap_eval = int(sc.get("envmap_eval", True))
        p.skylike_background = int(sc.get("skylike_background", False))
        p.envmap_rotation_deg = sc.get("envmap_rotation", 0.0)
        p.envmap_strength = sc.get("envmap_strength", 1.0)
        self.c = p


def render_path(scene, nbvh_nodes, const, net, path, pixel_begin, pixel_end, stride=1, sample_idx=0, rnd_offset=0,
                mlp_mode=0, n_threads=0):
    """Radiance (n, 3) of one sample per pixel, the primary-hit records and [queries, nbvh steps, blas steps, tri tests, rays]."""
    nbvh_nodes = np.ascontiguousarray(nbvh_nodes, dtype=NODE_DTYPE)
    n = (pixel_end - pixel_begin + stride - 1) // stride
    rad = np.zeros((n, 3), np.float32)
    hits = np.zeros(n, dtype=HIT_DTYPE)
    cnt = (C.c_uint64 * 5)()
    lib().onbvh_render_path(C.byref(scene.c), _p(nbvh_nodes), C.byref(const), C.byref(net), C.byref(path.c),
                            C.c_uint32(pixel_begin), C.c_uint32(pixel_end), C.c_uint32(stride), C.c_int(sample_idx),
                            C.c_int(rnd_offset), C.c_int(mlp_mode), _p(rad), _p(hits), cnt, C.c_int(n_threads))
    return rad, hits, [cnt[i] for i in range(5)]


# ----------------------------------------------------------------------------- tree-cut learning (f2)
TREE_CUT_DTYPE = np.dtype([("loss", "<f4"), ("iter_alive", "<u4"), ("depth", "<u4"), ("samples", "<u4")])
NO_NEURAL_CHILD = 0xFFFFFFFF


def tree_cut_initial(blas_nodes, min_depth, first_learned_lod):
    """traverse_and_build_neural_bvh (neural_bvh_renderer_accel.cpp:24-74), recursive exactly as written there."""
    import sys
    nodes, cut, nmap = [None], [None], [None]

    def rec(bi, ni, depth):
        b = blas_nodes[bi]
        nmap[ni] = bi
        cut[ni] = (0.0, 1, depth, 0)
        if b["a"] > 0 or depth == min_depth:
            nodes[ni] = (b["mn"], 1 if b["a"] > 0 else 2 + first_learned_lod, b["mx"], NO_NEURAL_CHILD)
            return
        left = len(nodes)
        nodes.extend([None, None]), cut.extend([None, None]), nmap.extend([None, None])
        nodes[ni] = (b["mn"], 0, b["mx"], left)
        rec(int(b["b"]), left, depth + 1)
        rec(int(b["b"]) + 1, left + 1, depth + 1)

    old = sys.getrecursionlimit()
    sys.setrecursionlimit(10000)
    rec(0, 0, 0)
    sys.setrecursionlimit(old)
    out = np.zeros(len(nodes), dtype=NODE_DTYPE)
    for i, t in enumerate(nodes):
        out[i] = t
    return out, np.array(cut, dtype=TREE_CUT_DTYPE), np.array(nmap, dtype=np.uint32)


def tree_cut_step(nodes, cut, nmap, blas_nodes, L, sampled, num_splits, current_learned_lod, dims=8):
    """One step of the reference's tree-cut learning, restated sequentially (neural_bvh_renderer_kernels.cu:202-287,
    neural_bvh_renderer.cu:199-268, neural_bvh_renderer_accel.cpp:256-392).  Loss sums in float64 then rounded (the reference's
    float atomics are unordered).  Returns (nodes, cut, nmap, max_node_loss, n_split)."""
    nodes, cut, nmap = nodes.copy(), cut.copy(), nmap.copy()
    per_sample = L[:, :dims].astype(np.float64).sum(axis=1)
    loss = cut["loss"].astype(np.float64)
    np.add.at(loss, sampled, per_sample)
    cut["loss"] = loss.astype(np.float32)
    np.add.at(cut["samples"], sampled, np.uint32(dims))
    leaf = nodes["b"] == NO_NEURAL_CHILD
    cut["iter_alive"][leaf] += 1
    cand = leaf & (nodes["a"] != 1) & (cut["samples"] > 0)
    ia = cut["iter_alive"].astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        h = cut["loss"] / (ia * np.sqrt(cut["samples"].astype(np.float32) / ia))
    h = np.where(cand, h, -np.inf).astype(np.float32)
    pos = np.flatnonzero(h > 0)
    order = pos[np.argsort(-h[pos], kind="stable")]
    max_loss = float(h[order[0]]) if len(order) else 0.0
    # distinct values, best first; at most min(k, #leaves) of them; then ascending node index (std::set)
    vals, first = np.unique(-h[order], return_index=True)
    chosen = order[np.sort(first)][: min(num_splits, (len(nodes) + 1) // 2)]
    n_split = 0
    new_nodes, new_cut, new_map = [], [], []
    n = len(nodes)
    for idx in np.sort(chosen):
        b = blas_nodes[nmap[idx]]
        if b["a"] > 0:
            continue
        left = n + 2 * n_split
        if nodes["a"][idx] == 2 + current_learned_lod:
            nodes["a"][idx] = 0
        nodes["b"][idx] = left
        nodes["mn"][idx], nodes["mx"][idx] = b["mn"], b["mx"]
        for c in (0, 1):
            ch = blas_nodes[int(b["b"]) + c]
            new_nodes.append((ch["mn"], 1 if ch["a"] > 0 else 2 + current_learned_lod, ch["mx"], NO_NEURAL_CHILD))
            new_cut.append((0.0, 1, cut["depth"][idx] + 1, 0))
            new_map.append(int(b["b"]) + c)
        n_split += 1
    if n_split:
        add = np.zeros(len(new_nodes), dtype=NODE_DTYPE)
        for i, t in enumerate(new_nodes):
            add[i] = t
        nodes = np.concatenate([nodes, add])
        cut = np.concatenate([cut, np.array(new_cut, dtype=TREE_CUT_DTYPE)])
        nmap = np.concatenate([nmap, np.array(new_map, dtype=np.uint32)])
    return nodes, cut, nmap, max_loss, n_split


# ----------------------------------------------------------------------------- online optimisation
LOSS_TYPES = ["RelativeL2", "L2", "RelativeL1", "L1", "RelativeL2Luminance", "BinaryCrossEntropy", "LogL2", "LHalf"]


class OLoss(C.Structure):
    _fields_ = [("vis_dim", C.c_uint32), ("depth_dim", C.c_uint32), ("normal_dim", C.c_uint32), ("albedo_dim", C.c_uint32),
                ("vis_loss", C.c_uint32), ("depth_loss", C.c_uint32), ("normal_loss", C.c_uint32), ("albedo_loss", C.c_uint32),
                ("vis_scale", C.c_float), ("depth_scale", C.c_float), ("normal_scale", C.c_float), ("albedo_scale", C.c_float),
                ("normal_linear", C.c_uint32), ("depth_linear", C.c_uint32), ("albedo_linear", C.c_uint32),
                ("vis_masking", C.c_uint32)]


class OOpt(C.Structure):
    _fields_ = [("learning_rate", C.c_float), ("decay_start", C.c_int32), ("decay_interval", C.c_int32),
                ("decay_base", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float),
                ("l2_reg", C.c_float)]


def make_loss(cfg):
    """PathTracingOutputConstants defaults: neural_bvh_module.h:243-274."""
    pt = cfg.get("neural_bvh_renderer", {}).get("neural_module", {}).get("pt_output_constants", {})
    lc = OLoss()
    lc.vis_dim = pt.get("output_visibility_dim", 1)
    lc.depth_dim = pt.get("output_depth_dim", 1)
    lc.normal_dim = pt.get("output_normal_dim", 0)
    lc.albedo_dim = pt.get("output_albedo_dim", 0)
    lc.vis_loss = LOSS_TYPES.index(pt.get("visibility_loss", "BinaryCrossEntropy"))
    lc.depth_loss = LOSS_TYPES.index(pt.get("depth_loss", "L1"))
    lc.normal_loss = LOSS_TYPES.index(pt.get("normal_loss", "L1"))
    lc.albedo_loss = LOSS_TYPES.index(pt.get("albedo_loss", "RelativeL2"))
    lc.vis_scale = pt.get("visibility_loss_scale", 1.0)
    lc.depth_scale = pt.get("depth_loss_scale", 1.0)
    lc.normal_scale = pt.get("normal_loss_scale", 1.0)
    lc.albedo_scale = pt.get("albedo_loss_scale", 1.0)
    lc.normal_linear = int(pt.get("output_normal_linear_activation", True))
    lc.depth_linear = int(pt.get("output_depth_linear_activation", False))
    lc.albedo_linear = int(pt.get("output_albedo_linear_activation", False))
    lc.vis_masking = int(pt.get("visibility_masking_loss", True))
    return lc


def make_opt(cfg):
    op = cfg.get("neural_bvh_renderer", {}).get("neural_module", {}).get("optimizer", {})
    return OOpt(op.get("learning_rate", 1e-2), op.get("decay_start", 10000), op.get("decay_interval", 2000),
                op.get("decay_base", 0.33), op.get("beta1", 0.9), op.get("beta2", 0.999), op.get("epsilon", 1e-8),
                op.get("l2_reg", 1e-8))


class Trainer:
    """CPU restatement of tcnn::Trainer + Adam/ExponentialDecay around the oracle network (fp32 master weights,
    fp16 parameters, per-parameter step counters; trainer.h:68-87, optimizers/adam.h:48-123)."""

    def __init__(self, cfg, params_bits):
        self.cfg = cfg
        self.params = np.ascontiguousarray(params_bits, dtype=np.uint16).copy()
        self.net = make_net(cfg, self.params)
        self.params = self.net._keep  # the array the C side reads
        self.loss_c = make_loss(cfg)
        self.opt_c = make_opt(cfg)
        n = self.params.size
        self.master = self.params.view(np.float16).astype(np.float32)
        self.m = np.zeros(n, np.float32)
        self.v = np.zeros(n, np.float32)
        self.param_steps = np.zeros(n, np.uint32)
        self.grads = np.zeros(n, np.float32)
        self.step = 0
        self.n_matrix = (self.net.n_neurons * self.net.n_input + (self.net.n_hidden_layers - 1) * self.net.n_neurons ** 2 +
                         self.net.n_output_padded * self.net.n_neurons)
        self.loss_scale = 128.0

    def forward_backward(self, x, target, global_batch=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        target = np.ascontiguousarray(target, dtype=np.float32)
        b = x.shape[0]
        L = np.zeros((b, self.net.n_output_padded), np.float32)
        s = C.c_double(0)
        lib().onbvh_train_forward_backward(C.byref(self.net), C.byref(self.loss_c), _p(x), _p(target), C.c_uint32(b),
                                           C.c_uint32(global_batch or b), C.c_float(self.loss_scale), _p(L), _p(self.grads),
                                           C.byref(s))
        return L, s.value

    def optimizer_step(self):
        lib().onbvh_adam_step(C.byref(self.opt_c), C.c_uint32(self.step), C.c_uint64(self.params.size),
                              C.c_uint64(self.n_matrix), C.c_float(self.loss_scale), _p(self.grads), _p(self.master),
                              _p(self.params), _p(self.m), _p(self.v), _p(self.param_steps))
        self.step += 1

    def training_step(self, x, target):
        L, loss = self.forward_backward(x, target)
        self.optimizer_step()
        return L, loss
