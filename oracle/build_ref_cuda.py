"""Builds the REFERENCE's own CUDA op, unmodified, for sm_100a (TEST INFRASTRUCTURE).

Sources are compiled where they lie under /root/reference/models/ops/src (vision.cpp,
cpu/ms_deform_attn_cpu.cpp, cuda/ms_deform_attn_cuda.cu); nothing is copied into the repo.  The only
accommodation for torch 2.11 is the force-included oracle/ref_compat.h (see there).  Output:
oracle/_ref/MultiScaleDeformableAttention_ref.so, a torch extension exposing the reference's
ms_deform_attn_forward / ms_deform_attn_backward (vision.cpp:13-16).  It is git-ignored but travels to
the GPU box, where tests/test_parity_gpu.py uses it as the second oracle and bench.py's `gpu_baseline` leg
(time_reference_cuda_op) times it as the "kernel to beat".  It cannot run here (no GPU) and /root/reference does not exist on the GPU box, so
build here, load there.
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/models/ops/src"
NAME = "MultiScaleDeformableAttention_ref"
SO = os.path.join(OUT, NAME + ".so")


def build(force: bool = False) -> str:
    if os.path.exists(SO) and not force:
        return SO
    if not os.path.isdir(REF_SRC):
        raise RuntimeError(f"{REF_SRC} not present (only the build container has the reference)")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils import cpp_extension
    build_dir = os.path.join(OUT, "build_ref_cuda")
    os.makedirs(build_dir, exist_ok=True)
    compat = os.path.join(HERE, "ref_compat.h")
    sources = [os.path.join(REF_SRC, "vision.cpp"), os.path.join(REF_SRC, "cpu", "ms_deform_attn_cpu.cpp"),
               os.path.join(REF_SRC, "cuda", "ms_deform_attn_cuda.cu")]
    cpp_extension.load(
        name=NAME, sources=sources, extra_include_paths=[REF_SRC],
        extra_cflags=["-DWITH_CUDA", "-include", compat, "-w"],
        extra_cuda_cflags=["-DWITH_CUDA", "-include", compat, "-w", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                           "-D__CUDA_NO_HALF2_OPERATORS__"],           # the reference's own flags, setup.py:33-38
        build_directory=build_dir, is_python_module=False, verbose=False)
    shutil.copy(os.path.join(build_dir, NAME + ".so"), SO)
    return SO


def load():
    """Import the built extension as a Python module (GPU box)."""
    if not os.path.exists(SO):
        raise RuntimeError(f"{SO} not built")
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force=True))
