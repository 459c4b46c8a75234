"""oracle/build_ref_ext.py -- TEST INFRASTRUCTURE.  Container-only build recipe.

Compiles the REFERENCE's own kernels and host wrappers -- the nine files under
/root/reference/.../pointnet2_ops/_ext-src/src, unmodified and where they lie -- for sm_100a with
plain nvcc/g++ (via torch.utils.cpp_extension.load, i.e. ninja + direct compiler calls; the
reference's setup.py is NOT run and would target sm_37..sm_75 only, pointnet2_ops_lib/setup.py:19).
Output: oracle/_ref/pn2_ref_ext.so (git-ignored, travels to the GPU box with the snapshot).

On the B200 this module IS the reference: tests/test_gpu_ref_ext.py checks oracle/pn2_oracle.c and
libsg4d.so against it bit for bit, and bench.py times it as the "reference kernels on the same GPU"
row.  No reference source is copied into the repository.
"""
import glob
import os
import sys

REF_SRC = "/root/reference/scene_graph_prediction/pointnet2_dir/pointnet2_ops_lib/pointnet2_ops/_ext-src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
NAME = "pn2_ref_ext"


def so_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
    if not os.path.isdir(REF_SRC):
        return None
    if os.path.exists(so_path()):
        return so_path()
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load
    srcs = sorted(glob.glob(os.path.join(REF_SRC, "src", "*.cpp")) + glob.glob(os.path.join(REF_SRC, "src", "*.cu")))
    load(NAME, sources=srcs, extra_include_paths=[os.path.join(REF_SRC, "include")], extra_cflags=["-O3"],
         extra_cuda_cflags=["-O3", "-gencode", "arch=compute_100a,code=sm_100a"], build_directory=OUT,
         with_cuda=True, is_python_module=False, verbose=verbose)
    return so_path()


REF_PKG = "/root/reference/scene_graph_prediction"
STAGE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "reference")


def stage_reference_python():
    """Container-only: "installs" the reference's Python (the *.py / *.json files of scene_graph_prediction/, 236 KB,
    unmodified, same tree) into the git-ignored baseline/_ref/reference/ so that the reference arm can run the
    reference's OWN model code on the GPU box, where /root/reference does not exist (the reference has no setup.py
    for this package; `pip install /root/reference` has nothing to install).  Nothing here enters the repository."""
    import shutil
    if not os.path.isdir(REF_PKG):
        return None
    dst_root = os.path.join(STAGE, "scene_graph_prediction")
    for root, dirs, files in os.walk(REF_PKG):
        for d in dirs:      # keep the package tree (some packages are bare directories = namespace packages)
            os.makedirs(os.path.join(dst_root, os.path.relpath(os.path.join(root, d), REF_PKG)), exist_ok=True)
        for f in files:
            if f.endswith((".py", ".json")):
                src = os.path.join(root, f)
                dst = os.path.join(dst_root, os.path.relpath(src, REF_PKG))
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
                    shutil.copyfile(src, dst)
    return STAGE


def load_module():
    """Import the prebuilt extension (GPU box or container); None when it was never built."""
    p = so_path()
    if not os.path.exists(p):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
