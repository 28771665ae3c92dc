"""Build the UNMODIFIED reference CUDA extensions into oracle/_ref/ (test infrastructure only).

This compiles the reference's own sources *where they lie* under /root/reference
(pointnet2/src/* -> pointnet2_cuda, src/projectPN/fused_conv_select/* ->
fused_conv_select_k_cuda) with the reference's own flags (nvcc -O2, see
pointnet2/setup.py:19-20 and fused_conv_select/setup.py:10-11) for sm_100.
No reference source is copied into this repository; only the resulting .so files
land in oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).

The .so files are the *checker* on the GPU box: tests compare our sm_100a kernels
against these legacy kernels bit for bit.  Nothing in the product path imports them.

Usage:  python oracle/build_ref.py            (no-op when /root/reference is absent)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("I2P_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

EXTS = {
    "pointnet2_cuda": [
        "pointnet2/src/pointnet2_api.cpp",
        "pointnet2/src/ball_query.cpp", "pointnet2/src/ball_query_gpu.cu",
        "pointnet2/src/group_points.cpp", "pointnet2/src/group_points_gpu.cu",
        "pointnet2/src/interpolate.cpp", "pointnet2/src/interpolate_gpu.cu",
        "pointnet2/src/sampling.cpp", "pointnet2/src/sampling_gpu.cu",
    ],
    "fused_conv_select_k_cuda": [
        "src/projectPN/fused_conv_select/fused_conv_g.cpp",
        "src/projectPN/fused_conv_select/fused_conv_go.cu",
    ],
}


def build(force=False):
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} absent: keeping prebuilt oracle/_ref as is")
        return False
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load
    for name, srcs in EXTS.items():
        target = os.path.join(OUT, name + ".so")
        if os.path.exists(target) and not force:
            print(f"[build_ref] {target} exists")
            continue
        bdir = os.path.join(OUT, "build_" + name)
        os.makedirs(bdir, exist_ok=True)
        load(name=name, sources=[os.path.join(REF, s) for s in srcs],
             extra_cflags=["-g"], extra_cuda_cflags=["-O2"],
             build_directory=bdir, verbose=False, is_python_module=False)
        shutil.copy(os.path.join(bdir, name + ".so"), target)
        shutil.rmtree(bdir, ignore_errors=True)
        print(f"[build_ref] built {target}")
    return True


if __name__ == "__main__":
    build(force="--force" in sys.argv)
