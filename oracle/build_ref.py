"""Build the REFERENCE's own silhouette rasteriser (external/neural_renderer, the only native code in the reference tree) as a
torch extension into oracle/_ref/ -- TEST INFRASTRUCTURE: the GPU parity tests of chore_b200's rasteriser compare against it.

The sources are compiled from where they lie under /root/reference (nothing is copied); the output directory is git-ignored but
travels to the GPU box with the snapshot.  The reference targets torch 1.6; its `AT_DISPATCH_*(tensor.type(), ...)` spelling no longer
compiles under torch 2.x, which oracle/ref_torch_compat.h (a forced-include shim) bridges without touching the source.

    python -m oracle.build_ref            # needs nvcc + ninja; ~2-3 minutes
"""
import os
import sys

REF = os.environ.get("CHORE_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "external", "neural_renderer", "neural_renderer", "cuda")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
NAME = "nmr_rasterize_ref"


def so_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
    from torch.utils.cpp_extension import load
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    return load(name=NAME, sources=[os.path.join(SRC, "rasterize_cuda.cpp"), os.path.join(SRC, "rasterize_cuda_kernel.cu")],
                build_directory=OUT, extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-w", "-include",
                                   os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_torch_compat.h")],
                extra_cflags=["-O2", "-w"], verbose=verbose, is_python_module=True)


def load_built():
    """Import the prebuilt extension (GPU box: /root/reference is absent there, only oracle/_ref/*.so exists)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    if not os.path.exists(so_path()):
        return None
    spec = importlib.util.spec_from_file_location(NAME, so_path())
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    if not os.path.isdir(SRC):
        sys.exit(f"{SRC} not found: the reference rasteriser can only be built where /root/reference exists")
    m = build(verbose="-v" in sys.argv)
    print("built", so_path(), [n for n in dir(m) if not n.startswith("_")])
