"""ctypes binding of ``liba2v_sm100.so`` (the C-ABI declared in ``include/a2v_capi.h``).

There is deliberately no fallback: if the shared library is missing, or a kernel is
asked to run on a device that is not sm_100, the call raises. PyTorch is used only for
device memory and streams; every pointer handed to the library comes from a torch tensor
the caller owns.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liba2v_sm100.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

F32 = 0
BF16 = 1


class A2VError(RuntimeError):
    pass


def build(verbose: bool = False, jobs: int = 8) -> str:
    """Compile every CUDA source for sm_100a into the in-tree shared library."""
    cmd = ["make", "-C", CSRC_DIR, f"-j{jobs}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise A2VError("building liba2v_sm100.so failed:\n" + res.stderr[-4000:])
    return LIB_PATH


class Operand(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p),
        ("dim0", C.c_int64),
        ("dim1", C.c_int64),
        ("dim2", C.c_int64),
        ("stride1", C.c_int64),
        ("stride2", C.c_int64),
    ]


class GemmDesc(C.Structure):
    _fields_ = [
        ("mode", C.c_int),
        ("block_n", C.c_int),
        ("a", Operand),
        ("b", Operand),
        ("M", C.c_int),
        ("N", C.c_int),
        ("k_per_tap", C.c_int),
        ("taps", C.c_int),
        ("batch", C.c_int),
        ("groups", C.c_int),
        ("a_group_stride", C.c_int),
        ("a_row_off", C.c_int),
        ("a_tap_rows", C.c_int),
        ("b_group_stride", C.c_int),
        ("b_row_off", C.c_int),
        ("b_tap_rows", C.c_int),
        ("red_rows", C.c_int),
        ("k_splits", C.c_int),
        ("c", C.c_void_p),
        ("c_dtype", C.c_int),
        ("out_atomic", C.c_int),
        ("out_accumulate", C.c_int),
        ("ldc", C.c_int64),
        ("c_batch_stride", C.c_int64),
        ("c_row_off", C.c_int64),
        ("c_group_stride", C.c_int),
        ("c_tap_stride", C.c_int),
        ("alpha", C.c_float),
        ("bias", C.c_void_p),
        ("act", C.c_int),
        ("preact", C.c_void_p),
        ("residual", C.c_void_p),
        ("dgelu_u", C.c_void_p),
        ("a_tap_cols", C.c_int),
        ("a_tap_wrap", C.c_int),
        ("a_grow_add", C.c_int),
        ("a_grow_div", C.c_int),
        ("a_tap_col_stride", C.c_int),
        ("colsum", C.c_void_p),
    ]


class ConvDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p),
        ("w", C.c_void_p),
        ("y", C.c_void_p),
        ("batch", C.c_int),
        ("T", C.c_int),
        ("groups", C.c_int),
        ("taps", C.c_int),
        ("pad", C.c_int),
        ("ng", C.c_int),
        ("x_group_cols", C.c_int),
        ("w_group_rows", C.c_int),
        ("y_group_cols", C.c_int),
        ("ldx", C.c_int64),
        ("ldw", C.c_int64),
        ("ldy", C.c_int64),
        ("y_dtype", C.c_int),
        ("bias", C.c_void_p),
        ("x_real_cols", C.c_int),
    ]


_lib = None
_lock = threading.Lock()

# Every C-ABI compute call launches exactly one kernel; the counter is bench.py's "gpu_launches" evidence.
launch_count = 0
# Optional per-GEMM timing (bench.py roofline leg): list of (start_event, end_event, flops) when enabled.
gemm_timeline = None
# Optional per-call timing of EVERY C-ABI call (tools/profile_step.py): list of (name, start, end) events.
op_timeline = None


def timed_call(name, fn):
    """Run ``fn`` (one C-ABI launch) bracketed by CUDA events when op_timeline is enabled."""
    tl = op_timeline
    if tl is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    tl.append((name, e0, e1))
    return r


def exported_symbols() -> list[str]:
    """Every ``a2v_*`` function name declared in include/a2v_capi.h."""
    import re

    hdr = os.path.join(os.path.dirname(_HERE), "include", "a2v_capi.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"\b(a2v_[a-z0-9_]+)\s*\(", txt)))


def load() -> C.CDLL:
    """Load the shared library (no compute is triggered). Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise A2VError(
                    f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(there is no CPU or PyTorch fallback for the kernels)"
                )
            lib = C.CDLL(LIB_PATH)
            lib.a2v_last_error.restype = C.c_char_p
            _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().a2v_last_error().decode("utf-8", "replace")
        raise A2VError(f"{what} failed (status {rc}): {msg}")


_dev_ok: dict[int, bool] = {}


def require_device(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise A2VError("animal2vec_b200 kernels need CUDA tensors on an sm_100 device; there is no CPU path")
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    ok = _dev_ok.get(idx)
    if ok is None:
        with torch.cuda.device(idx):
            ok = bool(load().a2v_device_supported())
        _dev_ok[idx] = ok
    if not ok:
        raise A2VError("device is not sm_100 (B200): the kernels are sm_100a-only")


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise A2VError(f"unsupported dtype {t.dtype}")
