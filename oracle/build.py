"""Build recipe for the oracle's native pieces (test infrastructure, see oracle/__init__.py).

  build_c()   gcc  oracle/nms_c.c                          -> oracle/_build/liboracle_nms.so
  build_ref() nvcc /root/reference/lib/nms/nms_kernel.cu   -> oracle/_ref/libref_gpu_nms.so

build_ref compiles the reference's own CUDA NMS *where it lies* (no copy of the source
enters this repo), for sm_100a, with -fmad=false so that its devIoU is the plain IEEE
sequence the C restatement follows.  It only exists for differential tests on the GPU box
(tests/test_nms_gpu.py); /root/reference is absent there, so the prebuilt .so travels.
The reference's Cython NMS (cpu_nms.pyx) does not compile against Cython 3 / numpy 2
(`np.int_t`, `np.int`), hence the C restatement is the only CPU form of it.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CU = "/root/reference/lib/nms/nms_kernel.cu"


def _newer(dst, srcs):
    return os.path.exists(dst) and all(os.path.getmtime(dst) >= os.path.getmtime(s) for s in srcs)


def build_c(force=False):
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(HERE, "nms_c.c")
    dst = os.path.join(out_dir, "liboracle_nms.so")
    if force or not _newer(dst, [src]):
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off",
                               "-o", dst, src])
    return dst


def build_ref(force=False):
    """Returns the path of the compiled reference kernel, or None when /root/reference is absent
    (GPU box) and no prebuilt copy travelled."""
    out_dir = os.path.join(HERE, "_ref")
    dst = os.path.join(out_dir, "libref_gpu_nms.so")
    if not os.path.exists(REF_CU):
        return dst if os.path.exists(dst) else None
    os.makedirs(out_dir, exist_ok=True)
    if force or not _newer(dst, [REF_CU]):
        subprocess.check_call(["nvcc", "-O2", "-shared", "-Xcompiler", "-fPIC", "-fmad=false",
                               "-gencode", "arch=compute_100a,code=sm_100a",
                               "-I", os.path.dirname(REF_CU), "-o", dst, REF_CU])
    return dst


if __name__ == "__main__":
    print(build_c(force=True))
    print(build_ref(force=True))
