"""Builds libnbvh_b200.so (hand-written CUDA for sm_100a + the C-ABI host side) in-tree with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnbvh_b200.so")
SOURCES = ["context.cu", "config.cu", "io.cu", "abi_render.cu", "abi_train.cu", "comm.cu", "kernels.cu", "render.cu", "path.cu", "treecut.cu", "build.cu", "training.cu", "sampler.cu", "flip.cu"]
HEADERS = ["host_context.hpp", "device_core.cuh", "mlp_tile.cuh", "kernels.cuh", "render_common.cuh", "sampler.cuh", "shading.cuh", "path.cuh", "treecut.cuh", "json_min.hpp", "../../include/nbvh_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # bit-exact traversal contract: only the explicit __fmaf_rn calls may fuse (see device_core.cuh)
    "-fmad=false",
    "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared",
]
LINK_LIBS = ["-ldl"]


OBJ_DIR = os.path.join(HERE, "_obj")


def _deps():
    return [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile_and_link(out, defines=(), tag="default", force=False, verbose=False):
    """One nvcc -c per translation unit (in parallel; only the units whose sources or headers changed), then one link."""
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    odir = os.path.join(OBJ_DIR, tag)
    os.makedirs(odir, exist_ok=True)
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    hdr_t = max(hdr_t, os.path.getmtime(os.path.abspath(__file__)))
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else [])
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(odir, s.replace(".cu", ".o"))
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr_t, os.path.getmtime(src)):
            jobs.append([nvcc] + flags + ["-c", src, "-o", obj])

    def run(cmd):
        print("[nbvh_b200.build]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(odir, s.replace(".cu", ".o")) for s in SOURCES]
    run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-shared", "-o", out] + objs + LINK_LIBS)
    return out


def build(force=False, verbose=False):
    if not force and not needs_build():
        build_host()
        return LIB
    _compile_and_link(LIB, force=force, verbose=verbose)
    build_host()
    return LIB


def build_host():
    """The C++ host side above the C ABI: include/nbvh_b200.hpp (header-only wrapper) + the command line hosts nbvh_render_cli
    (inference) and nbvh_train_cli (validation-mode run: optimisation with tree-cut learning, then frames and result files)."""
    outs = []
    for name in ("nbvh_render_cli", "nbvh_train_cli"):
        src = os.path.join(HERE, "host", name + ".cpp")
        out = os.path.join(HERE, "host", name)
        deps = [src, os.path.join(HERE, "..", "include", "nbvh_b200.hpp"), os.path.join(HERE, "..", "include", "nbvh_b200.h"),
                os.path.join(CSRC, "json_min.hpp"), LIB]
        outs.append(out)
        if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
            continue
        cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(HERE, "..", "include"), "-I/usr/local/cuda/include", src, "-L" + HERE,
               "-lnbvh_b200", "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,$ORIGIN/..", "-o", out]
        print("[nbvh_b200.build]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return outs


def build_variant(name, defines):
    """A/B experiments on the GPU box: the same sources with extra -D switches -> libnbvh_b200_<name>.so, selected at
    load time with NBVH_B200_LIB=<path> (development only; the product library is libnbvh_b200.so)."""
    import shutil
    out = os.path.join(HERE, "libnbvh_b200_%s.so" % name)
    _compile_and_link(out, defines=defines, tag="variant_" + name, force=True)
    shutil.rmtree(os.path.join(OBJ_DIR, "variant_" + name), ignore_errors=True)  # the repo snapshot sent to the GPU box is size-limited
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        build_variant(sys.argv[i + 1], sys.argv[i + 2:])
    else:
        build(force="--force" in sys.argv, verbose="-v" in sys.argv)
