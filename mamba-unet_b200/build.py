"""Build libselscan_b200.so (sm_100a only) in-tree with nvcc.  No GPU is needed to compile.

    python mamba-unet_b200/build.py [--force] [--verbose]

The library lands in mamba-unet_b200/lib/ (git-ignored, but it travels to the GPU box with gpurun).
"""
import argparse
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libselscan_b200.so")
SOURCES = ["selscan_api.cu", "selscan_fwd.cu", "selscan_fwd_tma.cu", "selscan_bwd.cu", "selscan_bwd_ws.cu", "selscan_cross.cu", "selscan_ss2d.cu", "selscan_ln.cu", "selscan_tcgemm.cu"]
HEADERS = ["selscan_common.cuh", "selscan_kernels.h", "selscan_ptx.cuh", "selscan_tma_host.h", os.path.join("..", "..", "include", "selscan_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: build a variant of the library (extra -D macros) to another path, for A/B timing via SELSCAN_B200_LIB."""
    if out is not None:
        return _build_variant(list(defines), out)
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(LIBDIR, s.replace(".cu", ".o"))
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(out)
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    # cudart is linked statically (nvcc default): the .so depends only on libcuda/libstdc++
    link = [nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    subprocess.check_call(link)
    return LIB


def _build_variant(defines, out):
    import tempfile
    tmp = tempfile.mkdtemp(prefix="selscan_variant_")
    objs, procs = [], []
    for s in SOURCES:
        o = os.path.join(tmp, s.replace(".cu", ".o"))
        procs.append((s, subprocess.Popen([nvcc_path()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-c", os.path.join(CSRC, s), "-o", o],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, pr in procs:
        o, _ = pr.communicate()
        if pr.returncode != 0:
            sys.stderr.write(o)
            raise RuntimeError(f"nvcc failed on {s}")
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    subprocess.check_call([nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs)
    shutil.rmtree(tmp, ignore_errors=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("-D", dest="defines", action="append", default=[])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose, defines=a.defines, out=a.out))
