"""Build the REFERENCE's own selective_scan_cuda extension for sm_100a into oracle/_ref/ (build container only).

TEST / BASELINE INFRASTRUCTURE: this is the unmodified reference kernel code, compiled from the sources where they
lie under /root/reference/mamba/csrc/selective_scan/ (nothing is copied into the repo).  It is the GPU-side second
oracle and the "reference on B200" baseline of BASELINE.md row B2; the product never loads it.

Only the TUs Mamba-UNet's call pattern needs are compiled (selective_scan.cpp, selective_scan_fwd_fp32.cu,
selective_scan_bwd_fp32_real.cu).  The host dispatcher (selective_scan.cpp:14-51,328-332,483-487) also references
the half/bf16/complex instantiations; oracle/ref_stub.cu -- our own scaffolding, no reference code -- defines those
as TORCH_CHECK(false) so that the link succeeds.  Flags follow mamba/setup.py:141-153; the arch list is replaced by
sm_100a (the reference ships sm_70/80/90 SASS only, setup.py:108-114).
"""
import os
import sys

REF = "/root/reference/mamba/csrc/selective_scan"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SO = os.path.join(OUT, "selective_scan_cuda.so")


def main():
    if not os.path.isdir(REF):
        print("oracle/build_ref.py: /root/reference not present, nothing to do")
        return 0
    srcs = [os.path.join(REF, f) for f in ("selective_scan.cpp", "selective_scan_fwd_fp32.cu", "selective_scan_bwd_fp32_real.cu")]
    stub = os.path.join(HERE, "ref_stub.cu")
    deps = srcs + [stub, os.path.abspath(__file__)]
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        print("oracle/_ref/selective_scan_cuda.so is up to date")
        return 0
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load

    load(name="selective_scan_cuda", sources=srcs + [stub], extra_include_paths=[REF],
         extra_cflags=["-O3", "-std=c++17"],
         extra_cuda_cflags=["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                            "-U__CUDA_NO_BFLOAT16_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
                            "-U__CUDA_NO_BFLOAT162_OPERATORS__", "-U__CUDA_NO_BFLOAT162_CONVERSIONS__",
                            "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math", "-lineinfo"],
         build_directory=OUT, verbose=False, is_python_module=False)
    assert os.path.exists(SO), SO
    print("built", SO, f"{os.path.getsize(SO) / 1e6:.1f} MB")
    return 0


if __name__ == "__main__":
    sys.exit(main())
