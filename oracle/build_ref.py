"""Build the UNMODIFIED reference CUDA extensions into oracle/_ref/ (test infrastructure only).

This compiles the reference's own sources *where they lie* under /root/reference
(pointnet2/src/* -> pointnet2_cuda, src/projectPN/fused_conv_select/* ->
fused_conv_select_k_cuda) with the reference's own flags (nvcc -O2, see
pointnet2/setup.py:19-20 and fused_conv_select/setup.py:10-11) for sm_100.
No reference source is copied into this repository; only the resulting .so files
land in oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).

The .so files are the *checker* on the GPU box: tests compare our sm_100a kernels
against these legacy kernels bit for bit.  Nothing in the product path imports them.

The reference's *Python* (model assembly, modules, operator wrappers, loss) is staged next to them,
byte for byte, under oracle/_ref/py/ (git-ignored like the .so files, shipped by gpurun): /root/reference does
not exist on the GPU box, and the live model-level oracle / legacy-kernel baseline (oracle/ref_live.py) imports
the unchanged files from there.

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


# the reference files oracle/ref_live.py imports (directly or transitively), relative to the checkout
PY_FILES = [
    "pointnet_util.py", "compute_loss.py",
    "pointnet2/pointnet2_utils.py", "pointnet2/pointnet2_modules.py", "pointnet2/pytorch_utils.py",
    "src/config_proj_lidarcenter.py", "src/config_proj_lidarcenter_nus.py", "src/config_lidarcenter.py",
    "src/modellearn_proj_center.py", "src/modellearn_proj_center_iter.py", "src/modellearn.py",
    "src/utils.py", "src/deterministic.py",
    "src/modules/__init__.py", "src/modules/MainModules.py", "src/modules/basicConv.py", "src/modules/point_utils.py",
    "src/modules/pointnet2_module.py", "src/modules/warp_utils.py",
    "src/projectPN/PPBackbone_center.py", "src/projectPN/utils.py",
    "src/projectPN/fused_conv_select/fused_conv_select_k.py",
    "src/util/__init__.py", "src/util/tracker.py",
]


def stage_python():
    """Copy the reference's own Python, unmodified, into oracle/_ref/py/ (same relative paths)."""
    if not os.path.isdir(REF):
        return False
    dst_root = os.path.join(OUT, "py")
    for rel in PY_FILES:
        dst = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    print(f"[build_ref] staged {len(PY_FILES)} reference .py files under {dst_root}")
    return True


def build(force=False):
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} absent: keeping prebuilt oracle/_ref as is")
        return False
    os.makedirs(OUT, exist_ok=True)
    stage_python()
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
