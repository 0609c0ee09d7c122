"""Result writers of the reference's validation mode, host side: thin calls into the C ABI.

  write_exr      == write_exr (/root/reference/include/base/bitmap.h:97-182): A,B,G,R float scanline EXR, no compression
  FrameTimer     == ntwr::FrameTimer (/root/reference/include/utils/frame_timer.h)
  result_path    == the file names of save_image / save_performance_metrics
                    (/root/reference/src/cuda/neural/neural_bvh_renderer_init.cu:328-372)
  read_exr       a minimal reader of exactly that layout (tests and tools read the images back without an EXR library)
"""
import ctypes as C
import os
import struct
import time

import numpy as np

from . import _capi


def _check(rc, what):
    if rc != 0:
        raise OSError("%s failed (code %d)" % (what, rc))


def write_exr(path, rgba, rgb_renormalisation=1.0):
    """rgba: float32 [H, W, 4] host array; every value (alpha too) is divided by rgb_renormalisation, like the reference."""
    a = np.ascontiguousarray(rgba, dtype=np.float32)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("write_exr expects an [H, W, 4] float image")
    lib = _capi.load_library()
    _check(lib.nbvh_write_exr_host(os.fsencode(path), a.shape[1], a.shape[0], a.ctypes.data_as(C.c_void_p), float(rgb_renormalisation)),
           "nbvh_write_exr_host")


def read_exr(path):
    """Reads an uncompressed scanline EXR with 32-bit float channels (what write_exr / tinyexr's NONE mode write) -> [H, W, 4] RGBA."""
    b = open(path, "rb").read()
    if b[:4] != b"\x76\x2f\x31\x01":
        raise ValueError("not an OpenEXR file")
    pos, attrs = 8, {}
    while b[pos] != 0:
        e = b.index(b"\0", pos)
        name = b[pos:e].decode()
        e2 = b.index(b"\0", e + 1)
        size = struct.unpack_from("<i", b, e2 + 1)[0]
        attrs[name] = (b[e + 1:e2].decode(), b[e2 + 5:e2 + 5 + size])
        pos = e2 + 5 + size
    pos += 1
    if attrs["compression"][1] != b"\0":
        raise ValueError("compressed EXR")
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    names, ch = [], attrs["channels"][1]
    p = 0
    while ch[p] != 0:
        e = ch.index(b"\0", p)
        names.append(ch[p:e].decode())
        if struct.unpack_from("<i", ch, e + 1)[0] != 2:
            raise ValueError("non-float channel")
        p = e + 17
    offsets = struct.unpack_from("<%dQ" % h, b, pos)
    planes = np.zeros((h, len(names), w), np.float32)
    for off in offsets:
        y, size = struct.unpack_from("<2i", b, off)
        planes[y - y0] = np.frombuffer(b, "<f4", len(names) * w, off + 8).reshape(len(names), w)
    out = np.zeros((h, w, 4), np.float32)
    for i, n in enumerate(names):
        out[:, :, "RGBA".index(n)] = planes[:, i, :]
    return out


def result_path(scene_dir, base_config_name, variant_config_name, kind, accumulated_spp):
    """kind: "image" | "learning" | "inference"."""
    lib = _capi.load_library()
    buf = C.create_string_buffer(4096)
    _check(lib.nbvh_result_path(os.fsencode(scene_dir), base_config_name.encode(), variant_config_name.encode(),
                                {"image": 0, "learning": 1, "inference": 2}[kind], int(accumulated_spp), buf, 4096), "nbvh_result_path")
    return os.fsdecode(buf.value)


class FrameTimer:
    """start_frame(frame_idx) / end_frame() / write_to_file(filename, losses=None): the reference's timing CSV."""

    def __init__(self):
        self.frame_times = []
        self._t = None

    def start_frame(self, frame_idx):
        if frame_idx == 0:
            self.frame_times = []
        self._t = time.perf_counter()

    def end_frame(self):
        now = time.perf_counter()
        self.frame_times.append((now - self._t) * 1e3)
        self._t = now

    def write_to_file(self, filename, losses=None):
        lib = _capi.load_library()
        t = np.asarray(self.frame_times, np.float64)
        lo = None
        if losses is not None:
            lo = np.ascontiguousarray(losses, np.float32)
            if len(lo) < len(t):
                raise ValueError("fewer losses than frames")
        _check(lib.nbvh_write_frame_times_csv_host(os.fsencode(filename), t.ctypes.data_as(C.c_void_p),
                                                   lo.ctypes.data_as(C.c_void_p) if lo is not None else None, len(t)),
               "nbvh_write_frame_times_csv_host")
