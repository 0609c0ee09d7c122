"""CPU tests: C-ABI library loads and exports every declared symbol, POD layouts match the header, the product path
fails loudly without a GPU, and the multi-GPU host logic (world_size 2, gloo)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "nbvh_b200.h")


def test_library_exports_every_declared_symbol():
    from nbvh_b200 import _capi
    lib = _capi.load_library()
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(nbvh_[a-z0-9_]+)\s*\(", text))
    assert len(declared) >= 39
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    assert lib.nbvh_abi_version() == 1


def test_pod_layouts_match_header(tmp_path):
    """sizeof/offsetof of every POD block as gcc sees the header == the ctypes mirrors."""
    from nbvh_b200 import _capi
    structs = {"nbvh_renderer_constants": _capi.RendererConstants, "nbvh_module_constants": _capi.ModuleConstants,
               "nbvh_pt_output_constants": _capi.PtOutputConstants, "nbvh_path_constants": _capi.PathConstants,
               "nbvh_scene_constants": _capi.SceneConstants, "nbvh_camera_params": _capi.CameraParams,
               "nbvh_bbox3f": _capi.Bbox3f, "nbvh_network_desc": _capi.NetworkDesc,
               "nbvh_optimizer_desc": _capi.OptimizerDesc, "nbvh_cuda_blas": _capi.CudaBlas,
               "nbvh_scene_data": _capi.SceneData}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "nbvh_b200.h"', "int main(void){"]
    for cname, ct in structs.items():
        src.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for f in ct._fields_:
            if f[0].startswith("_"):
                continue
            src.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f[0], cname, f[0]))
    src.append("return 0;}")
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = str(tmp_path / "layout")
    subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), str(c), "-o", exe])
    out = dict(l.split() for l in subprocess.check_output([exe], text=True).splitlines())
    for cname, ct in structs.items():
        assert int(out[cname]) == C.sizeof(ct), cname
        for f in ct._fields_:
            if not f[0].startswith("_"):
                assert int(out["%s.%s" % (cname, f[0])]) == getattr(ct, f[0]).offset, (cname, f[0])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import nbvh_b200
    with pytest.raises(nbvh_b200.NbvhError) as e:
        nbvh_b200.NeuralPathModule()
    assert e.value.code == -2


def test_product_package_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "nbvh_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "oracle" not in text.replace("oracle/ref_tools", "").lower() or f == "scene_io.py" and "from oracle" not in text, f


def test_shard_pixel_range_partitions_frame():
    from nbvh_b200.parallel import shard_batch_slots, shard_pixel_range
    for n in (1920 * 1080, 3840 * 2160, 1000, 33, 32, 1):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            prev_end = 0
            for r in range(world):
                first, count = shard_pixel_range(r, world, n)
                assert first % 32 == 0
                assert first == prev_end or count == 0
                prev_end = first + count if count else prev_end
                covered += count
            assert covered == n
    for b in (1 << 18, 1000):
        for world in (1, 2, 8):
            assert sum(shard_batch_slots(r, world, b)[1] for r in range(world)) == b


_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from nbvh_b200.parallel import allreduce_gradients, shard_batch_slots, shard_pixel_range
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
g = torch.full((1000,), float(r + 1))
allreduce_gradients(g)
assert torch.all(g == sum(range(1, w + 1))), g[:4]
first, count = shard_pixel_range(r, w, 1920 * 1080)
t = torch.tensor([count], dtype=torch.int64)
dist.all_reduce(t)
assert int(t) == 1920 * 1080
dist.barrier()
dist.destroy_process_group()
print("ok", r)
"""


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % REPO)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)], env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("ok") == 2
