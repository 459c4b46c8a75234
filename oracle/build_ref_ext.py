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
