"""Loader for libmsda_sm100.so, the C-ABI CUDA library (include/msda_sm100.h).

The library is built IN-TREE by ``build()`` (``nvcc -gencode arch=compute_100a,code=sm_100a``) into
``ocpg_b200/lib/``; there is no JIT cache and no fallback: if the library is missing or does not load,
every operator raises (the product path never routes through a CPU implementation).
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from ctypes import c_char_p, c_int, c_uint64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
SRC = os.path.join(_PKG, "csrc", "msda_sm100.cu")
INCLUDE = os.path.join(_ROOT, "include")
LIB_PATH = os.environ.get("MSDA_LIB") or os.path.join(_PKG, "lib", "libmsda_sm100.so")   # MSDA_LIB: experiments

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

# every symbol include/msda_sm100.h declares (tests/test_boundary.py checks header <-> this list <-> the .so)
SYMBOLS = [
    "msda_abi_version", "msda_last_error", "msda_forward_f32", "msda_backward_f32", "msda_forward_f64",
    "msda_backward_f64", "msda_forward_bf16", "msda_backward_bf16", "msda_kernel_plan", "msda_launch_count",
    "msda_set_option", "msda_fused_forward_f32", "msda_fused_backward_f32", "msda_fused_forward_bf16",
    "msda_fused_backward_bf16", "msda_epilogue_ln_forward_f32", "msda_epilogue_ln_backward_f32", "msda_column_sum_f32",
    "msda_relu_backward_column_sum_f32", "msda_epilogue_ln_dropout_forward_f32", "msda_epilogue_ln_dropout_backward_f32",
    "msda_dropout_inplace_f32", "msda_relu_dropout_backward_column_sum_f32", "msda_dropout_mask_u8",
    "msda_decoder_select_samples_f32", "msda_decoder_reference_points_f32", "msda_flatten_levels_f32",
    "msda_unflatten_levels_f32",
]

_lib = None


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libmsda_sm100.so")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/msda_sm100.cu for sm_100a into ocpg_b200/lib/libmsda_sm100.so (cross-compiles
    without a GPU).  Rebuilds when the source or header is newer than the library."""
    csrc = os.path.join(_PKG, "csrc")
    deps = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(INCLUDE, "msda_sm100.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps)
    if force or stale:
        os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
        cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-o", LIB_PATH, SRC]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
        if verbose:
            print(proc.stderr)
    return LIB_PATH


def _code_only(text: str) -> str:
    """C/C++ source with comments and all whitespace removed (string literals kept verbatim)."""
    out, i, n = [], 0, len(text)
    while i < n:
        c = text[i]
        if c == '"' or c == "'":
            j = i + 1
            while j < n and text[j] != c:
                j += 2 if text[j] == "\\" else 1
            out.append(text[i:j + 1]); i = j + 1
        elif text.startswith("//", i):
            j = text.find("\n", i); i = n if j < 0 else j
        elif text.startswith("/*", i):
            j = text.find("*/", i + 2); i = n if j < 0 else j + 2
        elif c.isspace():
            i += 1
        else:
            out.append(c); i += 1
    return "".join(out)


def source_fingerprint() -> str:
    """First 16 hex digits of the SHA-256 of the kernel sources' CODE (csrc/*, include/msda_sm100.h; comments and
    whitespace do not count): ties profiles (ncu captures, profiles/traffic.json) to the kernels they were taken from."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(_PKG, "csrc")
    for f in sorted(os.listdir(csrc)) + [os.path.join(INCLUDE, "msda_sm100.h")]:
        path = f if os.path.isabs(f) else os.path.join(csrc, f)
        if path.endswith((".cu", ".cuh", ".h")):
            with open(path, "r", encoding="utf-8") as fh:
                h.update(os.path.basename(path).encode() + b"\0" + _code_only(fh.read()).encode())
    return h.hexdigest()[:16]


def lib() -> ctypes.CDLL:
    """The loaded library; raises RuntimeError (never falls back) if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (or ocpg_b200.build()). "
            "There is no CPU fallback for MSDeformAttn (the reference has none either: "
            "models/ops/src/cpu/ms_deform_attn_cpu.cpp:17-41).")
    L = ctypes.CDLL(LIB_PATH)
    missing = [s for s in SYMBOLS if not hasattr(L, s)]
    if missing:
        raise RuntimeError(f"{LIB_PATH} lacks symbols {missing}; rebuild it")
    L.msda_abi_version.restype = c_int
    L.msda_last_error.restype = c_char_p
    L.msda_launch_count.restype = c_uint64
    L.msda_kernel_plan.restype = c_int
    L.msda_kernel_plan.argtypes = [c_int] * 5
    L.msda_set_option.restype = c_int
    L.msda_set_option.argtypes = [c_char_p, c_int]
    dims = [c_int] * 7
    for sfx in ("f32", "f64", "bf16"):
        f = getattr(L, f"msda_forward_{sfx}")
        f.restype = c_int
        f.argtypes = [c_void_p] * 5 + dims + [c_void_p, c_void_p]
        b = getattr(L, f"msda_backward_{sfx}")
        b.restype = c_int
        n_out = 4 if sfx == "bf16" else 3
        b.argtypes = [c_void_p] * 6 + dims + [c_void_p] * n_out + [c_void_p]
    for sfx in ("f32", "bf16"):
        f = getattr(L, f"msda_fused_forward_{sfx}")
        f.restype = c_int
        f.argtypes = [c_void_p] * 6 + [c_int] + dims + [c_void_p] * 3 + [c_void_p]
        b = getattr(L, f"msda_fused_backward_{sfx}")
        b.restype = c_int
        n_out = 5 if sfx == "bf16" else 4
        b.argtypes = [c_void_p] * 7 + [c_int] + dims + [c_void_p] * n_out + [c_void_p]
    from ctypes import c_float, c_int64
    L.msda_epilogue_ln_forward_f32.restype = c_int
    L.msda_epilogue_ln_forward_f32.argtypes = [c_void_p] * 5 + [c_float, c_int64, c_int] + [c_void_p] * 4 + [c_void_p]
    L.msda_epilogue_ln_backward_f32.restype = c_int
    L.msda_epilogue_ln_backward_f32.argtypes = [c_void_p] * 5 + [c_int64, c_int] + [c_void_p] * 4 + [c_void_p]
    L.msda_column_sum_f32.restype = c_int
    L.msda_column_sum_f32.argtypes = [c_void_p, c_int64, c_int, c_void_p, c_void_p]
    L.msda_relu_backward_column_sum_f32.restype = c_int
    L.msda_relu_backward_column_sum_f32.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]
    from ctypes import c_uint32
    L.msda_epilogue_ln_dropout_forward_f32.restype = c_int
    L.msda_epilogue_ln_dropout_forward_f32.argtypes = ([c_void_p] * 5 + [c_float, c_int64, c_int] + [c_void_p, c_uint32, c_float]
                                                       + [c_void_p] * 4 + [c_void_p])
    L.msda_epilogue_ln_dropout_backward_f32.restype = c_int
    L.msda_epilogue_ln_dropout_backward_f32.argtypes = ([c_void_p] * 5 + [c_int64, c_int] + [c_void_p, c_uint32, c_float]
                                                        + [c_void_p] * 5 + [c_void_p])
    L.msda_dropout_inplace_f32.restype = c_int
    L.msda_dropout_inplace_f32.argtypes = [c_void_p, c_int64, c_void_p, c_uint32, c_float, c_void_p]
    L.msda_relu_dropout_backward_column_sum_f32.restype = c_int
    L.msda_relu_dropout_backward_column_sum_f32.argtypes = [c_void_p, c_void_p, c_float, c_int64, c_int, c_void_p, c_void_p, c_void_p]
    L.msda_dropout_mask_u8.restype = c_int
    L.msda_dropout_mask_u8.argtypes = [c_void_p, c_uint32, c_float, c_int64, c_void_p, c_void_p]
    L.msda_decoder_select_samples_f32.restype = c_int
    L.msda_decoder_select_samples_f32.argtypes = [c_void_p] * 3 + [c_int] * 6 + [c_void_p] * 3 + [c_void_p]
    L.msda_decoder_reference_points_f32.restype = c_int
    L.msda_decoder_reference_points_f32.argtypes = [c_void_p] * 2 + [c_int] * 4 + [c_void_p, c_void_p]
    L.msda_flatten_levels_f32.restype = c_int
    L.msda_flatten_levels_f32.argtypes = [c_int] + [c_void_p] * 5 + [c_int, c_int] + [c_void_p] * 3
    L.msda_unflatten_levels_f32.restype = c_int
    L.msda_unflatten_levels_f32.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
    if L.msda_abi_version() != 4:
        raise RuntimeError("libmsda_sm100.so ABI version mismatch")
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    """Turn a non-zero return code into RuntimeError (the reference raises c10::Error -> RuntimeError,
    ms_deform_attn_cuda.cu:28-52)."""
    if rc != 0:
        msg = lib().msda_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def raw_stream(device) -> int:
    """cudaStream_t of torch's CURRENT stream on ``device`` (ms_deform_attn_cuda.cu:65 uses the current stream too).
    ``torch._C._cuda_getCurrentRawStream`` is the entry point torch's own compiled-kernel launchers use; it skips the
    Stream object the public API builds (8 us per call, ~0.5 ms of host time per eager encoder step)."""
    import torch
    try:
        return torch._C._cuda_getCurrentRawStream(device.index)
    except AttributeError:                                    # older / newer torch without the private hook
        return torch.cuda.current_stream(device).cuda_stream


class on_device:
    """``with on_device(t.device):`` -- make the tensors' device current for the library call (the C ABI launches on the
    calling thread's current device), like ``torch.cuda.device`` but without its per-use index normalisation."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index
        self.prev = -1

    def __enter__(self):
        import torch
        self.prev = torch.cuda._exchange_device(self.idx)
        return self

    def __exit__(self, *exc):
        import torch
        self.prev = torch.cuda._maybe_exchange_device(self.prev)
        return False


def launch_count() -> int:
    return int(lib().msda_launch_count())


def set_option(key: str, value: int) -> None:
    check(lib().msda_set_option(key.encode(), int(value)), f"msda_set_option({key})")
