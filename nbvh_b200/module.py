"""Host-side mirror of the reference's module interface for the neural ray-query path.

`NeuralPathModule` follows `ntwr::neural::path::NeuralPathModule : NeuralBVHModule`
(/root/reference/include/cuda/neural/neural_bvh_module.h:144-241,
 /root/reference/include/cuda/neural/path/path_module.h:256-345): same method names, argument meaning
and error behaviour (the reference throws std::runtime_error; here `NbvhError`).  Every method is a thin
call into the C ABI of libnbvh_b200.so -- PyTorch is used only to own device memory and streams.
"""
import ctypes as C
import json
import os

import numpy as np

from . import _capi
from ._capi import (CONST_CAMERA, CONST_MODULE, CONST_NETWORK, CONST_OPTIMIZER, CONST_PATH, CONST_PT_OUTPUT,
                    CONST_RENDERER, CONST_SCENE, CONST_SCENE_BBOX, CONST_TYPES, COUNTER_NAMES)

NODE_DTYPE = np.dtype([("mn", "<f4", 3), ("a", "<u4"), ("mx", "<f4", 3), ("b", "<u4")])
# PackedMaterial (include/base/material.h:52-87): vtable pointer + Disney parameters; a negative float encodes a texture handle
MATERIAL_DTYPE = np.dtype([("vptr", "<u8"), ("base_color", "<f4", 3), ("metallic", "<f4"), ("specular", "<f4"), ("roughness", "<f4"),
                           ("specular_tint", "<f4"), ("anisotropy", "<f4"), ("sheen", "<f4"), ("sheen_tint", "<f4"),
                           ("clearcoat", "<f4"), ("clearcoat_gloss", "<f4"), ("base_emission", "<f4", 3), ("emission_scale", "<f4"),
                           ("normal", "<f4", 3), ("ior", "<f4"), ("specular_transmission", "<f4"), ("pad", "<f4")])
TREE_CUT_DTYPE = np.dtype([("loss", "<f4"), ("iter_alive", "<u4"), ("depth", "<u4"), ("samples", "<u4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("valid_hit", "<u4"), ("normal_or_uv", "<f4", 3), ("albedo_or_ids", "<f4", 3)])


class NbvhError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("nbvh_b200 error %s (%d): %s" % (_capi.ERR_NAMES.get(code, "?"), code, msg))
        self.code = code


def _torch():
    import torch
    return torch


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, int):  # a raw device pointer (e.g. the context-owned tree-cut statistics)
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr())


class NeuralPathModule:
    """One instance per process per device, like the reference (its __constant__ symbols are process-global)."""

    def __init__(self, config=None, device=0):
        self._lib = _capi.load_library()
        self._h = C.c_void_p()
        self.device = device
        if isinstance(config, dict):
            config = json.dumps(config)
        cfg = config.encode() if isinstance(config, str) else config
        rc = self._lib.nbvh_create(int(device), cfg, C.byref(self._h))
        if rc != 0:
            msg = self._lib.nbvh_last_error(self._h).decode() if self._h else "no usable sm_100 CUDA device"
            if self._h:
                self._lib.nbvh_destroy(self._h)
                self._h = C.c_void_p()
            raise NbvhError(rc, msg)
        self._keep = {}

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc != 0:
            raise NbvhError(rc, self._lib.nbvh_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.nbvh_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def name(self):
        return "Path Tracing Module"

    # ------------------------------------------------------------------ configuration
    def load_config_from_json(self, config):
        """== NeuralBVHRenderer::load_config_from_json (neural_bvh_renderer_init.cu:566-749)."""
        if isinstance(config, dict):
            config = json.dumps(config)
        self._check(self._lib.nbvh_load_config_json(self._h, config.encode()))

    def get_constants(self, which):
        pod = CONST_TYPES[which]()
        self._check(self._lib.nbvh_get_constants(self._h, which, C.byref(pod), C.sizeof(pod)))
        return pod

    def set_constants(self, which, pod):
        self._check(self._lib.nbvh_set_constants(self._h, which, C.byref(pod), C.sizeof(pod)))

    def setup_neural_module_constants(self, **blocks):
        """== setup_neural_module_constants(): pushes constant blocks (keyword = block id name)."""
        for which, pod in blocks.items():
            self.set_constants(which, pod)

    def set_camera(self, matrix16, yfov_rad):
        m = (C.c_float * 16)(*[float(x) for x in matrix16])
        self._check(self._lib.nbvh_set_camera(self._h, m, C.c_float(yfov_rad)))

    @property
    def max_depth(self):
        return self._lib.nbvh_get_max_depth(self._h)

    @max_depth.setter
    def max_depth(self, v):
        self._check(self._lib.nbvh_set_max_depth(self._h, int(v)))

    def network_output_dims(self):
        return self.get_constants(CONST_PT_OUTPUT).output_network_dims

    # ------------------------------------------------------------------ scene / tree-cut / weights
    def set_scene(self, tlas_nodes, tlas_indices, blas_list, materials=None, textures=None, neural_meshes=None):
        """Uploads a scene description (numpy arrays) and hands the device pointers over, standing in for
        CudaBackend's buffers (cuda_backend.cpp:186-441).  blas_list: dicts with nodes, prim_ids, vertices,
        indices (prim ids are widened to 64 bit: the reference stores size_t) and, for the training sampler, the optional
        shading attributes normals, texcoords, material_ids, material_map.  materials: MATERIAL_DTYPE array
        (PackedMaterial); textures: list of (H, W, 4) uint8 images (RGBA8, wrap + linear like the reference's).
        tlas_nodes=None builds the TLAS on the device (nbvh_build_tlas) over the uploaded BLAS; neural_meshes: indices of the meshes the
        N-BVH stands for (default: bvh_constants.neural_bvh_idx).  self.tlas_arrays() reads the result back."""
        torch = _torch()
        dev = torch.device("cuda", self.device)

        def up(a, dtype):
            a = np.ascontiguousarray(a)
            return torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(dev)

        keep = []
        build_tlas = tlas_nodes is None          # device TLAS build (nbvh_build_tlas) once the descriptors are uploaded
        if not build_tlas:
            t_nodes = up(np.ascontiguousarray(tlas_nodes, dtype=NODE_DTYPE), None)
            t_idx = up(np.ascontiguousarray(tlas_indices, dtype=np.uint32), None)
            keep += [t_nodes, t_idx]
        descs = (_capi.CudaBlas * len(blas_list))()
        for i, b in enumerate(blas_list):
            nodes = up(np.ascontiguousarray(b["nodes"], dtype=NODE_DTYPE), None)
            prim = up(np.ascontiguousarray(b["prim_ids"]).astype(np.uint64), None)
            verts = up(np.ascontiguousarray(b["vertices"], dtype=np.float32), None)
            idx = up(np.ascontiguousarray(b["indices"], dtype=np.uint32), None)
            keep += [nodes, prim, verts, idx]
            descs[i].num_nodes = len(b["nodes"])
            descs[i].prim_indices = prim.data_ptr()
            descs[i].num_prim_indices = len(b["prim_ids"])
            descs[i].vertices = verts.data_ptr()
            descs[i].indices = idx.data_ptr()
            descs[i].nodes = nodes.data_ptr()
            for key, dt in (("normals", np.float32), ("texcoords", np.float32), ("material_ids", np.int32), ("material_map", np.uint32)):
                if b.get(key) is not None:
                    t = up(np.ascontiguousarray(b[key], dtype=dt), None)
                    keep.append(t)
                    setattr(descs[i], key, t.data_ptr())
            if b.get("material_map") is not None:
                descs[i].num_materials = len(b["material_map"])
        d_descs = torch.from_numpy(np.frombuffer(bytes(descs), dtype=np.uint8).copy()).to(dev)
        keep.append(d_descs)
        n_tlas_nodes = len(tlas_nodes) if not build_tlas else 2 * len(blas_list) - 1
        if build_tlas:
            t_nodes = torch.zeros(n_tlas_nodes * 32, dtype=torch.uint8, device=dev)
            t_idx = torch.zeros(len(blas_list), dtype=torch.int32, device=dev)
            keep += [t_nodes, t_idx]
            neural = list(neural_meshes) if neural_meshes is not None else [int(self.get_constants(CONST_RENDERER).neural_bvh_idx)]
            arr = (C.c_uint32 * max(1, len(neural)))(*neural)
            self._check(self._lib.nbvh_build_tlas(self._h, d_descs.data_ptr(), len(blas_list), arr, len(neural), _ptr(t_nodes), _ptr(t_idx), None))
            torch.cuda.synchronize(dev)
        sd = _capi.SceneData()
        sd.tlas_nodes = t_nodes.data_ptr()
        sd.num_tlas_nodes = n_tlas_nodes
        sd.blas_indices = t_idx.data_ptr()
        sd.mesh_bvhs = d_descs.data_ptr()
        sd.num_meshes = len(blas_list)
        if materials is not None:
            m = up(np.ascontiguousarray(materials, dtype=MATERIAL_DTYPE), None)
            keep.append(m)
            sd.materials = m.data_ptr()
        if textures:
            self._check(self._lib.nbvh_destroy_textures(self._h))
            handles = np.zeros(len(textures), np.uint64)
            for k, img in enumerate(textures):
                img = np.ascontiguousarray(img, dtype=np.uint8)
                assert img.ndim == 3 and img.shape[2] == 4
                h = C.c_uint64(0)
                self._check(self._lib.nbvh_create_texture_rgba8(self._h, img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], C.byref(h)))
                handles[k] = h.value
            t = up(handles, None)
            keep.append(t)
            sd.textures = t.data_ptr()
        self._check(self._lib.nbvh_set_scene(self._h, C.byref(sd)))
        self._keep["scene"] = keep
        self._keep["tlas"] = (t_nodes, t_idx)

    def tlas_arrays(self):
        """Host copies of the TLAS in use: (nodes NODE_DTYPE array, blas_indices uint32 array)."""
        t_nodes, t_idx = self._keep["tlas"]
        return t_nodes.cpu().numpy().view(NODE_DTYPE).copy(), t_idx.cpu().numpy().view(np.uint32).copy()

    def set_envmap(self, pixels):
        """== CudaBackend::load_envmap (cuda_backend.cpp:455-500): (H, W, 4) float32 lat-long radiance map, or None for the
        reference's white fallback image.  Create it AFTER set_scene(textures=...) (which re-creates the texture set)."""
        if pixels is None:
            self._check(self._lib.nbvh_set_envmap(self._h, 0))
            return
        img = np.ascontiguousarray(pixels, dtype=np.float32)
        assert img.ndim == 3 and img.shape[2] == 4
        h = C.c_uint64(0)
        self._check(self._lib.nbvh_create_envmap_rgba32f(self._h, img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], C.byref(h)))
        self._check(self._lib.nbvh_set_envmap(self._h, h.value))

    def set_neural_bvh(self, nodes):
        torch = _torch()
        nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
        d = torch.from_numpy(nodes.view(np.uint8).reshape(-1).copy()).to(torch.device("cuda", self.device))
        self._check(self._lib.nbvh_set_neural_bvh(self._h, _ptr(d), len(nodes)))

    def load_neural_bvh(self, path):
        """== NeuralBVHRenderer::load_neural_bvh (neural_bvh_renderer_accel.cpp:208-254)."""
        self._check(self._lib.nbvh_load_neural_bvh_file(self._h, os.fsencode(path)))

    def save_neural_bvh(self, path):
        self._check(self._lib.nbvh_save_neural_bvh_file(self._h, os.fsencode(path)))

    def neural_bvh_nodes(self):
        n = self._lib.nbvh_neural_bvh_num_nodes(self._h)
        out = np.zeros(n, dtype=NODE_DTYPE)
        self._check(self._lib.nbvh_get_neural_bvh_host(self._h, out.ctypes.data_as(C.c_void_p), n))
        return out

    def load_network_model(self, path):
        """== NeuralBVHRenderer::load_network_model -> trainer->deserialize (neural_bvh_renderer_init.cu:394-404)."""
        self._check(self._lib.nbvh_load_network_file(self._h, os.fsencode(path)))

    def save_network_model(self, path):
        self._check(self._lib.nbvh_save_network_file(self._h, os.fsencode(path)))

    def n_params(self):
        return int(self._lib.nbvh_n_params(self._h))

    def load_params(self, fp16_bits):
        a = np.ascontiguousarray(fp16_bits, dtype=np.uint16)
        self._check(self._lib.nbvh_load_params_host(self._h, a.ctypes.data_as(C.c_void_p), a.size))

    def get_params(self):
        out = np.zeros(self.n_params(), dtype=np.uint16)
        self._check(self._lib.nbvh_get_params_host(self._h, out.ctypes.data_as(C.c_void_p), out.size))
        return out

    def init_params(self, seed=1337):
        self._check(self._lib.nbvh_init_params(self._h, int(seed)))

    # ------------------------------------------------------------------ inference
    def reinitialise_inference_data(self, fb_size):
        """== reinitialise_inference_data(fb_size): (re)allocates the frame buffers."""
        self._check(self._lib.nbvh_set_framebuffer(self._h, None, None, int(fb_size[0]), int(fb_size[1])))

    def attach_framebuffer(self, fb_size):
        """Caller-owned frame buffers (the reference's GL surfaces): returns (accum RGBA32F [H, W, 4], ldr sRGB8 [H, W, 4])."""
        torch = _torch()
        dev = torch.device("cuda", self.device)
        w, h = int(fb_size[0]), int(fb_size[1])
        accum = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
        ldr = torch.zeros((h, w, 4), dtype=torch.uint8, device=dev)
        self._check(self._lib.nbvh_set_framebuffer(self._h, _ptr(accum), _ptr(ldr), w, h))
        self._keep["fb"] = (accum, ldr)
        return accum, ldr

    def attach_hit_buffer(self, n_pixels):
        torch = _torch()
        buf = torch.zeros(n_pixels * HIT_DTYPE.itemsize, dtype=torch.uint8, device=torch.device("cuda", self.device))
        self._check(self._lib.nbvh_set_hit_buffer(self._h, _ptr(buf)))
        self._keep["hits"] = buf
        return buf

    def inference(self, inputs, outputs=None, stream=None):
        """== network->inference(stream, in 9xN, out 8xN): torch CUDA tensors, row i = query i."""
        torch = _torch()
        n = inputs.shape[0]
        if outputs is None:
            outputs = torch.empty((n, self.network_output_dims()), dtype=torch.float32, device=inputs.device)
        assert inputs.is_cuda and inputs.is_contiguous() and inputs.dtype == torch.float32
        self._check(self._lib.nbvh_query(self._h, _ptr(inputs), _ptr(outputs), n, stream))
        return outputs

    def trace_intervals(self, o, d, max_intervals, stream=None):
        torch = _torch()
        n = o.shape[0]
        dev = o.device
        counts = torch.zeros(n, dtype=torch.int32, device=dev)
        node = torch.zeros((n, max_intervals), dtype=torch.int32, device=dev)
        mint = torch.zeros((n, max_intervals), dtype=torch.float32, device=dev)
        maxt = torch.zeros((n, max_intervals), dtype=torch.float32, device=dev)
        self._check(self._lib.nbvh_trace_intervals(self._h, _ptr(o), _ptr(d), n, max_intervals, _ptr(counts), _ptr(node),
                                                   _ptr(mint), _ptr(maxt), stream))
        return counts, node, mint, maxt

    def blas_intersect(self, blas_idx, o, d, maxt=None, stream=None):
        torch = _torch()
        n = o.shape[0]
        dev = o.device
        t = torch.zeros(n, dtype=torch.float32, device=dev)
        u = torch.zeros(n, dtype=torch.float32, device=dev)
        v = torch.zeros(n, dtype=torch.float32, device=dev)
        prim = torch.zeros(n, dtype=torch.int32, device=dev)
        self._check(self._lib.nbvh_blas_intersect(self._h, blas_idx, _ptr(o), _ptr(d), _ptr(maxt), n, _ptr(t), _ptr(u),
                                                  _ptr(v), _ptr(prim), stream))
        return t, u, v, prim

    def raygen(self, first_pixel, n, sample_idx=0, sample_offset=0, stream=None):
        torch = _torch()
        dev = torch.device("cuda", self.device)
        o = torch.zeros((n, 3), dtype=torch.float32, device=dev)
        d = torch.zeros((n, 3), dtype=torch.float32, device=dev)
        self._check(self._lib.nbvh_raygen(self._h, first_pixel, n, sample_idx, sample_offset, _ptr(o), _ptr(d), stream))
        return o, d

    def encode_segments(self, o, d, mint, maxt, stream=None):
        """== encode_bvh_network_input_data (neural_bvh_module_kernels.cu:266-303): -> (n, input_network_dims) float32."""
        torch = _torch()
        n = o.shape[0]
        out = torch.zeros((n, self.get_constants(CONST_MODULE).input_network_dims), dtype=torch.float32, device=o.device)
        self._check(self._lib.nbvh_encode_segments(self._h, _ptr(o), _ptr(d), _ptr(mint), _ptr(maxt), n, _ptr(out), stream))
        return out

    def neural_render(self, accumulated_spp, sample_offset, first_pixel=0, pixel_count=0, stream=None):
        """== neural_render(accumulated_spp, sample_offset, neural_bvh) (path_module.cu:139-245)."""
        self._check(self._lib.nbvh_render(self._h, int(accumulated_spp), int(sample_offset), int(first_pixel),
                                          int(pixel_count), stream))

    def trace_rays(self, o, d, ray_depth=1, hits=None, stream=None):
        """Closest-hit neural ray query of a ray batch (o, d: (N, 3) float32 CUDA tensors).  Returns a uint8 tensor holding N
        HIT_DTYPE records."""
        torch = _torch()
        n = int(o.shape[0])
        if hits is None:
            hits = torch.empty(n * HIT_DTYPE.itemsize, dtype=torch.uint8, device=o.device)
        self._check(self._lib.nbvh_trace_rays(self._h, _ptr(o), _ptr(d), n, int(ray_depth), _ptr(hits), stream))
        return hits

    def neural_render_shard(self, accumulated_spp, sample_offset, shard_index, shard_count, stream=None):
        """neural_render restricted to an interleaved tile shard (every shard_count-th 8x4-pixel block): the multi-GPU partition."""
        self._check(self._lib.nbvh_render_interleaved(self._h, int(accumulated_spp), int(sample_offset), int(shard_index), int(shard_count), stream))

    def build_blas(self, vertices, indices, stream=None):
        """Device LBVH build (== CudaBackend's per-mesh BLAS build, on the GPU).  vertices: (V, 3) float32, indices: (T, 3) int32 CUDA
        tensors.  Returns (nodes uint8 tensor of 2T-1 32-byte records, prim_indices int64 tensor [T], tree depth)."""
        torch = _torch()
        n = int(indices.shape[0])
        nodes = torch.empty((2 * n - 1) * 32, dtype=torch.uint8, device=vertices.device)
        prims = torch.empty(n, dtype=torch.int64, device=vertices.device)
        depth = C.c_uint32(0)
        self._check(self._lib.nbvh_build_blas(self._h, _ptr(vertices), _ptr(indices), n, _ptr(nodes), _ptr(prims), C.byref(depth), stream))
        return nodes, prims, depth.value

    ERROR_MODES = {"AbsError": 3, "RelAbsError": 4, "MSEError": 5, "RelMSEError": 6}  # PostprocessMode, neural_bvh_renderer.h:83

    def image_error(self, reference, accumulated_spp, mode="RelMSEError", write_error_image=False, stream=None):
        """== compute_and_display_render_error: mean error of accum / spp against a (H, W, 4) float32 CUDA reference image."""
        out = C.c_float(0.0)
        self._check(self._lib.nbvh_image_error(self._h, _ptr(reference), int(accumulated_spp), self.ERROR_MODES[mode], int(write_error_image),
                                               C.byref(out), stream))
        return out.value

    def flip_error(self, reference, accumulated_spp, ppd=0.0, want_map=False, stream=None):
        """== compute_and_display_flip_error: mean LDR-FLIP of accum / spp against a (H, W, 4) float32 CUDA reference image; with
        want_map also the (H, W) per-pixel error map."""
        torch = _torch()
        assert reference.is_cuda and reference.dtype == torch.float32 and reference.is_contiguous()
        emap = torch.empty(reference.shape[:2], dtype=torch.float32, device=reference.device) if want_map else None
        mean = C.c_float(0)
        self._check(self._lib.nbvh_flip_error(self._h, _ptr(reference), int(accumulated_spp), float(ppd), _ptr(emap), C.byref(mean), stream))
        return (mean.value, emap) if want_map else mean.value

    def flip_images(self, reference, test, test_divisor=1.0, ppd=0.0, stream=None):
        """LDR-FLIP of two (H, W, 4) float32 CUDA images (linear RGB): returns (mean, (H, W) error map)."""
        torch = _torch()
        for t in (reference, test):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape == reference.shape and t.shape[2] == 4
        emap = torch.empty(reference.shape[:2], dtype=torch.float32, device=reference.device)
        mean = C.c_float(0)
        self._check(self._lib.nbvh_flip_images(self._h, _ptr(reference), _ptr(test), float(test_divisor), int(reference.shape[1]),
                                               int(reference.shape[0]), float(ppd), _ptr(emap), C.byref(mean), stream))
        return mean.value, emap

    # ------------------------------------------------------------------ results (f4)
    def save_image(self, path, accumulated_spp, stream=None):
        """== NeuralBVHRenderer::save_image (neural_bvh_renderer_init.cu:361-383): accumulation / accumulated_spp as an A,B,G,R float EXR."""
        self._check(self._lib.nbvh_save_image_exr(self._h, os.fsencode(path), int(accumulated_spp), stream))

    def save_config(self):
        """== the module-owned part of NeuralBVHRenderer::save_config (neural_bvh_renderer_init.cu:406-509); returns the JSON text."""
        need = C.c_size_t(0)
        self._lib.nbvh_save_config_json(self._h, None, 0, C.byref(need))
        buf = C.create_string_buffer(need.value)
        self._check(self._lib.nbvh_save_config_json(self._h, buf, need.value, None))
        return buf.value.decode()

    def neural_render_samples(self, first_spp, n_spp, sample_offset=0, shard_index=0, shard_count=1, stream=None):
        """n_spp samples per pixel of an interleaved shard in one launch sequence (max_depth > 1): keeps the persistent kernels full."""
        self._check(self._lib.nbvh_render_samples(self._h, int(first_spp), int(n_spp), int(sample_offset), int(shard_index), int(shard_count),
                                                  stream))

    def neural_render_host(self, accumulated_spp, sample_offset, ldr_host=None, hits_host=None, camera=None):
        """End-to-end call with HOST buffers (numpy arrays, ideally pinned)."""
        cam = C.byref(camera) if camera is not None else None
        lp = ldr_host.ctypes.data_as(C.c_void_p) if ldr_host is not None else None
        hp = hits_host.ctypes.data_as(C.c_void_p) if hits_host is not None else None
        self._check(self._lib.nbvh_render_host(self._h, int(accumulated_spp), int(sample_offset), cam, lp, hp))

    def neural_render_host_submit(self, accumulated_spp, sample_offset, ldr_host, slot, camera=None):
        """Frames in flight: like neural_render_host, but returns once the work is queued; ldr_host (pinned numpy array) is valid
        after neural_render_host_wait(slot).  Two slots: the copy of one frame overlaps the kernels of the next."""
        cam = C.byref(camera) if camera is not None else None
        self._check(self._lib.nbvh_render_host_submit(self._h, int(accumulated_spp), int(sample_offset), cam,
                                                      ldr_host.ctypes.data_as(C.c_void_p), int(slot)))

    def neural_render_host_wait(self, slot):
        self._check(self._lib.nbvh_render_host_wait(self._h, int(slot)))

    # ------------------------------------------------------------------ online optimisation
    def gradient_buffer(self):
        """fp32 gradient vector (n_params) owned by torch and registered with the library: NCCL allreduces it in place."""
        torch = _torch()
        if "grads" not in self._keep or self._keep["grads"].numel() != self.n_params():
            g = torch.zeros(self.n_params(), dtype=torch.float32, device=torch.device("cuda", self.device))
            self._check(self._lib.nbvh_set_gradient_buffer(self._h, _ptr(g)))
            self._keep["grads"] = g
        return self._keep["grads"]

    def optimisation_data_generation(self, training_step, batch_size, log2_bvh_max_traversals, sampled_node_indices, current_max_loss,
                                     tree_cut_data, inputs, targets, stream=None):
        """== NeuralBVHModule::optimisation_data_generation (neural_bvh_module.h:172-180, path_module.cu:20-52): torch CUDA
        tensors; sampled_node_indices int32/uint32 [B], tree_cut_data uint8 view of TREE_CUT_DTYPE[num_nodes] (or None),
        inputs float32 [B, in_dims], targets float32 [B, out_dims] -- all updated in place."""
        self._check(self._lib.nbvh_optimisation_data_generation(self._h, int(training_step), int(batch_size), int(log2_bvh_max_traversals),
                                                                _ptr(sampled_node_indices), C.c_float(current_max_loss),
                                                                _ptr(tree_cut_data), _ptr(inputs), _ptr(targets), stream))

    def sampler_walk_ahead(self, steps, step_stride=1):
        """One walk launch serves the next `steps` optimisation steps (training_step, training_step + stride, ...): same samples, the
        generator's rejection walks run at several slots per lane instead of 1.5 (nbvh_sampler_walk_ahead)."""
        self._check(self._lib.nbvh_sampler_walk_ahead(self._h, int(steps), int(step_stride)))

    def training_forward_backward(self, inputs, targets, global_batch=None, per_sample_loss=None, loss_sum=None, stream=None):
        """Forward + composite loss + backward; leaves loss-scaled (x128) fp32 gradients in the gradient buffer."""
        n = inputs.shape[0]
        assert inputs.is_cuda and inputs.is_contiguous() and targets.is_contiguous() and targets.shape[0] == n
        self._check(self._lib.nbvh_training_forward_backward(self._h, _ptr(inputs), _ptr(targets), n, int(global_batch or n),
                                                             _ptr(per_sample_loss), _ptr(loss_sum), stream))

    def optimizer_step(self, stream=None):
        self._check(self._lib.nbvh_optimizer_step(self._h, stream))

    def training_step(self, inputs, targets, per_sample_loss=None, want_loss=True, stream=None):
        """== trainer->training_step(stream, in 9xB, target 8xB) (+ trainer->loss when want_loss): returns the batch loss."""
        n = inputs.shape[0]
        loss = C.c_float(0)
        self._check(self._lib.nbvh_training_step(self._h, _ptr(inputs), _ptr(targets), n, _ptr(per_sample_loss),
                                                 C.byref(loss) if want_loss else None, stream))
        return loss.value if want_loss else None

    # ------------------------------------------------------------------ tree-cut learning (NeuralBVHRenderer's part of a step)
    def tree_cut_init(self, min_depth=0):
        """== build_and_upload_neural_bvh(neural_bvh_min_depth): initial cut of the neural mesh's BLAS (0 = the root alone)."""
        self._check(self._lib.nbvh_tree_cut_init(self._h, int(min_depth)))

    def tree_cut_attach(self):
        """== reinitialise_neural_bvh_learning_states for the tree-cut already loaded (node map, depths, zero statistics)."""
        self._check(self._lib.nbvh_tree_cut_attach(self._h))

    def tree_cut_step(self, per_sample_loss, sampled_node_indices, num_splits, dims=None, stream=None):
        """== update_neural_tree_cut_loss + update_neural_tree_cut_decisions + split_neural_bvh_at.  per_sample_loss: (B, 16) float32
        CUDA tensor (tcnn's ctx.L).  Returns (max_node_loss, leaves split)."""
        batch, stride = int(per_sample_loss.shape[0]), int(per_sample_loss.shape[1])
        mx, ns = C.c_float(0.0), C.c_uint32(0)
        self._check(self._lib.nbvh_tree_cut_step(self._h, _ptr(per_sample_loss), stride, int(dims if dims is not None else self.network_output_dims()),
                                                 _ptr(sampled_node_indices), batch, int(num_splits), C.byref(mx), C.byref(ns), stream))
        return mx.value, ns.value

    # ------------------------------------------------------------------ data-parallel optimisation inside the library (comm.cu)
    def setup_data_parallel(self, rank=None, world=None, group=None):
        """Gives the library its own NCCL communicator (nbvh_allreduce_setup_unique_id): rank 0 draws the unique id, the host's
        launcher -- here torch.distributed, any backend -- only carries those 128 bytes to the other ranks.  From then on
        training_forward_backward / training_step sum the gradients across the ranks themselves (two buckets on the library's
        communication stream, the first overlapping the grid backward)."""
        import torch.distributed as dist
        torch = _torch()
        from ._capi import NCCL_UNIQUE_ID_BYTES
        if rank is None:
            rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        buf = (C.c_ubyte * NCCL_UNIQUE_ID_BYTES)()
        if rank == 0:
            self._check(self._lib.nbvh_nccl_unique_id(buf))
        if world > 1:
            on_gpu = dist.get_backend(group) == "nccl"
            t = torch.tensor(list(buf), dtype=torch.uint8, device=torch.device("cuda", self.device) if on_gpu else "cpu")
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            buf = (C.c_ubyte * NCCL_UNIQUE_ID_BYTES)(*t.cpu().tolist())
        self._check(self._lib.nbvh_allreduce_setup_unique_id(self._h, buf, int(rank), int(world)))
        return rank, world

    def share_data_parallel(self, other):
        """Attaches the communicator `other` (a module of this process, which must outlive this one) already owns."""
        comm = other._lib.nbvh_nccl_comm(other._h)
        if not comm:
            raise NbvhError(-1, "share_data_parallel: the other module has no communicator")
        self._check(self._lib.nbvh_allreduce_setup(self._h, C.c_void_p(comm)))

    def teardown_data_parallel(self):
        self._check(self._lib.nbvh_allreduce_teardown(self._h))

    def data_parallel_world(self):
        r = C.c_int(0)
        w = self._lib.nbvh_data_parallel_world(self._h, C.byref(r))
        return r.value, w

    def allreduce(self, tensor, stream=None):
        """Sums a CUDA tensor over the library's ranks in place (no-op without a communicator)."""
        torch = _torch()
        from ._capi import DTYPE_F32, DTYPE_F64, DTYPE_I32, DTYPE_U32, DTYPE_U64
        kinds = {torch.float32: DTYPE_F32, torch.int32: DTYPE_I32, torch.float64: DTYPE_F64, torch.int64: DTYPE_U64, torch.uint8: None}
        if tensor.dtype == torch.uint8:  # bytes of disjoint tiles: summed as 32-bit words (no carries between disjoint non-zero bytes)
            assert tensor.numel() % 4 == 0
            self._check(self._lib.nbvh_allreduce_buffer(self._h, _ptr(tensor), tensor.numel() // 4, DTYPE_U32, stream))
            return tensor
        self._check(self._lib.nbvh_allreduce_buffer(self._h, _ptr(tensor), tensor.numel(), kinds[tensor.dtype], stream))
        return tensor

    def tree_cut_step_data_parallel(self, per_sample_loss, sampled_node_indices, num_splits, dims=None, group=None, stream=None):
        """tree_cut_step for one process per GPU: per-rank statistics -> allreduce (NCCL) -> identical decisions and splits everywhere."""
        import torch.distributed as dist
        torch = _torch()
        if self.data_parallel_world()[1] > 1:  # the library's own communicator
            n = self._lib.nbvh_neural_bvh_num_nodes(self._h)
            dev = per_sample_loss.device
            loss_d = torch.empty(n, dtype=torch.float32, device=dev)
            samp_d = torch.empty(n, dtype=torch.int32, device=dev)
            batch, stride = int(per_sample_loss.shape[0]), int(per_sample_loss.shape[1])
            self._check(self._lib.nbvh_tree_cut_accumulate(self._h, _ptr(per_sample_loss), stride,
                                                           int(dims if dims is not None else self.network_output_dims()),
                                                           _ptr(sampled_node_indices), batch, _ptr(loss_d), _ptr(samp_d), stream))
            self.allreduce(loss_d, stream=stream)
            self.allreduce(samp_d, stream=stream)
            mx, ns = C.c_float(0.0), C.c_uint32(0)
            self._check(self._lib.nbvh_tree_cut_decide(self._h, _ptr(loss_d), _ptr(samp_d), int(num_splits), C.byref(mx), C.byref(ns), stream))
            return mx.value, ns.value
        n = self._lib.nbvh_neural_bvh_num_nodes(self._h)
        dev = per_sample_loss.device
        loss_d = torch.empty(n, dtype=torch.float32, device=dev)
        samp_d = torch.empty(n, dtype=torch.int32, device=dev)
        batch, stride = int(per_sample_loss.shape[0]), int(per_sample_loss.shape[1])
        self._check(self._lib.nbvh_tree_cut_accumulate(self._h, _ptr(per_sample_loss), stride, int(dims if dims is not None else self.network_output_dims()),
                                                       _ptr(sampled_node_indices), batch, _ptr(loss_d), _ptr(samp_d), stream))
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(loss_d, group=group)
            dist.all_reduce(samp_d, group=group)
        mx, ns = C.c_float(0.0), C.c_uint32(0)
        self._check(self._lib.nbvh_tree_cut_decide(self._h, _ptr(loss_d), _ptr(samp_d), int(num_splits), C.byref(mx), C.byref(ns), stream))
        return mx.value, ns.value

    def tree_cut_assign_lod(self):
        """== assign_lod_to_current_neural_leaf_nodes.  Returns True when all LoDs have been assigned."""
        done = C.c_int(0)
        self._check(self._lib.nbvh_tree_cut_assign_lod(self._h, C.byref(done)))
        return bool(done.value)

    def tree_cut_data_ptr(self):
        return self._lib.nbvh_tree_cut_data_device(self._h)

    def set_tree_cut_data(self, cut):
        cut = np.ascontiguousarray(cut, dtype=TREE_CUT_DTYPE)
        self._check(self._lib.nbvh_set_tree_cut_host(self._h, cut.ctypes.data_as(C.c_void_p), len(cut)))

    def tree_cut_state(self):
        """Host copies of (tree-cut nodes, NeuralTreeCutData, node map)."""
        n = self._lib.nbvh_neural_bvh_num_nodes(self._h)
        cut = np.zeros(n, dtype=TREE_CUT_DTYPE)
        nmap = np.zeros(n, dtype=np.uint32)
        self._check(self._lib.nbvh_get_tree_cut_host(self._h, cut.ctypes.data_as(C.c_void_p), nmap.ctypes.data_as(C.c_void_p), n))
        return self.neural_bvh_nodes(), cut, nmap

    def training_step_count(self):
        return int(self._lib.nbvh_training_step_count(self._h))

    def selftest_division(self, n_trials, seed=1):
        bad = C.c_uint64(0)
        self._check(self._lib.nbvh_selftest_division(self._h, int(seed), int(n_trials), C.byref(bad)))
        return bad.value

    # ------------------------------------------------------------------ counters / timing
    def counters(self):
        out = (C.c_uint64 * len(COUNTER_NAMES))()
        self._check(self._lib.nbvh_get_counters(self._h, out, len(COUNTER_NAMES)))
        return dict(zip(COUNTER_NAMES, [int(x) for x in out]))

    def reset_counters(self):
        self._check(self._lib.nbvh_reset_counters(self._h))

    def enable_kernel_timing(self, on=True):
        self._check(self._lib.nbvh_enable_kernel_timing(self._h, int(on)))

    def kernel_time(self):
        ms, n = C.c_float(0), C.c_uint32(0)
        self._check(self._lib.nbvh_get_kernel_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def set_queue_order(self, limits=(4, 3, 2, 2), flags=1):
        """Drain order of the neural ray queue by expected query-chain length (scheduling only; limits (0, 0, 0, 0) = arrival order)."""
        t = [int(x) for x in limits]
        self._check(self._lib.nbvh_set_queue_order(self._h, t[0], t[1], t[2], t[3], int(flags)))

    def enable_stage_timing(self, max_frames):
        """Event triples around k_primary / k_neural for the next `max_frames` frames (no host synchronisation)."""
        self._check(self._lib.nbvh_enable_stage_timing(self._h, int(max_frames)))

    def enable_training_stage_timing(self, max_steps):
        self._check(self._lib.nbvh_enable_training_stage_timing(self._h, int(max_steps)))

    def training_stage_times(self):
        """-> ({encode, mlp, grid_backward, exchange_wait, adam} total ms, steps) since the last call."""
        ms, n = (C.c_float * 5)(), C.c_uint32(0)
        self._check(self._lib.nbvh_get_training_stage_times(self._h, ms, C.byref(n)))
        return dict(zip(("encode", "mlp", "grid_backward", "exchange_wait", "adam"), [float(v) for v in ms])), n.value

    def stage_times(self):
        """-> (k_primary total ms, k_neural total ms, frames) since the last call; synchronises on the events."""
        a, b, n = C.c_float(0), C.c_float(0), C.c_uint32(0)
        self._check(self._lib.nbvh_get_stage_times(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value
