"""ctypes binding of the C-ABI in include/statecatcher_b200.h.

This is the reference-side stub a maintainer would add: raw device
pointers (``tensor.data_ptr()``), sizes and the current CUDA stream go straight through
``extern "C"`` entry points of ``csrc/libstatecatcher_b200.so``.  There is no CPU fallback:
if the library is missing the import fails loudly, and every non-zero return code raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SC_B200_LIB") or os.path.join(_HERE, "csrc", "libstatecatcher_b200.so")   # SC_B200_LIB: A/B builds

SC_F32, SC_BF16 = 0, 1
SC_SCAN_CKPT = 8

P, I64, I32, F32 = c_void_p, c_int64, c_int, c_float

# name -> argtypes, exactly as declared in include/statecatcher_b200.h
SIGNATURES = {
    "sc_version": [],
    "sc_error_string": [I32],
    "sc_build_info": [c_char_p, I64],
    "sc_gemm_fwd": [P, I64, P, I64, P, P, I64, I64, I64, I64, I32, I32, I32, P],
    "sc_gemm_dgrad": [P, I64, P, I64, P, I64, I64, I64, I64, I32, I32, I32, P],
    "sc_gemm_wgrad": [P, I64, P, I64, P, I64, I64, I64, I64, I32, I32, I32, P],
    "sc_gemm_workspace_bytes": [I64, I64, I64],
    "sc_cast": [P, I64, I32, P, I64, I32, I64, I64, P],
    "sc_split_bf16": [P, I64, P, I64, I64, I64, P],
    "sc_split6_bf16": [P, I64, P, I64, I64, I64, I32, I64, P],
    "sc_colsum": [P, I64, I32, P, I64, I64, I32, P],
    "sc_layernorm_fwd": [P, I64, P, P, P, I64, P, P, I64, I64, I32, P],
    "sc_layernorm_bwd": [P, I64, P, I64, P, P, P, P, I64, P, P, P, I64, I64, I32, P],
    "sc_lucy_scan_fwd": [P, I64, P, P, P, I64, P, P, P, I64, I64, I64, I32, I32, P],
    "sc_lucy_scan_chunked_work_bytes": [I64, I64, I64],
    "sc_lucy_scan_fwd_chunked": [P, I64, P, P, P, I64, P, P, P, P, I64, I64, I64, I32, I32, P],
    "sc_lucy_scan_bwd": [P, I64, P, I64, P, P, P, P, I64, P, I64, P, I64, I64, I64, I32, I32, P],
    "sc_lucy_sscan_fwd": [P, P, P, I64, P, I64, P, P, I64, P, P, I64, I64, I64, I32, I32, I32, F32, P],
    "sc_lucy_sscan_bwd": [P, P, P, I64, P, P, P, I64, P, P, P, I64, P, I64, I64, I64, I32, I32, I32, F32, P],
    "sc_lucy_hscan_fwd": [P, I64, P, I64, P, P, I64, P, I64, I64, I64, I32, P],
    "sc_lucy_hscan_bwd": [P, I64, P, I64, P, I64, P, P, I64, P, I64, P, I64, I64, I64, I64, I32, P],
    "sc_ctc_workspace_bytes": [I64, I64, I64],
    "sc_ctc_lplat_pitch": [I64],
    "sc_ctc_fwd": [P, I64, I64, I32, P, I64, P, P, I64, I64, I64, I64, I64, P, P, P, P, P, P, P, I32, P, P],
    "sc_ctc_emissions": [P, I64, I64, I32, P, I64, P, P, I64, I64, I64, I64, I64, P, P, P, P],
    "sc_ctc_lattice": [P, P, P, I64, P, P, I64, I64, I64, I64, P, P, P, P, I32, P, P],
    "sc_ctc_bwd": [P, I64, I64, I32, P, I64, P, P, I64, I64, I64, I64, I64, P, P, P, P, P, I32,
                   P, I64, I64, I32, P, P],
    "sc_rnnt_fwd": [P, P, I64, P, P, I64, I64, I64, I64, I64, P, P, P, P, P, P, P],
    "sc_rnnt_bwd": [P, I64, P, P, I64, I64, I64, I64, I64, P, I64, P, P, P, P, P, P, P, P],
    "sc_joint_fwd": [P, I64, I64, P, I64, I64, P, I64, I64, I64, I64, I32, P],
    "sc_joint_bwd": [P, P, I64, I64, P, I64, I64, P, I64, I64, P, I64, I64, I64, I64, I32, P],
    "sc_rnnt_lse_gather": [P, I32, P, I64, P, P, I64, I64, I64, I64, I64, I64, I64, P, P, P, P],
    "sc_rnnt_lattice": [P, P, I64, I64, I64, P, P, P, P, P, P],
    "sc_rnnt_node_grads": [P, P, I64, I64, I64, P, P, P, P, P, P, P, P, P],
    "sc_rnnt_dlogits": [P, I32, P, P, P, P, I64, P, I64, I64, I64, I64, I64, I64, I64, P, P, P],
    "sc_ctc_head_phases": [I64, I64, I64],
    "sc_ctc_head": [P, I64, I64, I32, P, I64, P, P, I64, I64, I64, I64, I64, P, P, P, P, P, P, P, I32, P,
                    P, I64, I64, I32, I64, P, P, P, P],
    "sc_ctc_scale_grad": [P, I32, I64, P, P],
    "sc_ctc_greedy_decode": [P, I64, I64, I32, P, I64, I64, I64, I64, P, P, P, P],
    "sc_sumsq_accum": [P, I64, P, P],
    "sc_scale_grads": [P, I64, P, F32, P],
    "sc_adam_step": [P, P, P, P, I64, F32, F32, F32, F32, F32, I64, P, F32, I32, P],
    "sc_lion_step": [P, P, P, I64, F32, F32, F32, F32, P, F32, P],
    "sc_sumsq_accum_multi": [P, P, I64, P, P],
    "sc_adam_step_multi": [P, P, P, P, P, I64, F32, F32, F32, F32, F32, I64, P, F32, I32, P],
    "sc_lion_step_multi": [P, P, P, P, I64, F32, F32, F32, F32, P, F32, P],
    "sc_grads_pack_multi": [P, P, P, I64, I32, P],
    "sc_grads_unpack_multi": [P, P, P, I64, I32, F32, P],
    "sc_frontend_tables_len": [],
    "sc_frontend_tables": [P, I64, I32],
    "sc_frontend": [P, I64, I64, I64, P, I32, F32, P, I64, P, I64, P, P],
    "sc_mask_rows": [P, I64, I32, P, P, I64, I64, I64, P],
    "sc_frame_mask": [P, I64, I64, I64, I64, I64, F32, I64, P, P, P],
}
_RESTYPES = {"sc_error_string": c_char_p, "sc_gemm_workspace_bytes": I64, "sc_lucy_scan_chunked_work_bytes": I64,
             "sc_ctc_workspace_bytes": I64, "sc_ctc_lplat_pitch": I64, "sc_ctc_head_phases": I64,
             "sc_frontend_tables_len": I64}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m statecatcher_b200.build` "
            "(statecatcher_b200 has no CPU or PyTorch fallback path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if an export is missing
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, I32)
    _lib = lib
    return lib


class StatecatcherError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        msg = load().sc_error_string(rc).decode()
        raise StatecatcherError(f"{what} failed with code {rc}: {msg}")


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return SC_F32
    if t.dtype == torch.bfloat16:
        return SC_BF16
    raise TypeError(f"statecatcher_b200 kernels take float32 or bfloat16 tensors, got {t.dtype}")


def ptr(t):
    return 0 if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def device_ctx(t):
    """Context manager that makes ``t``'s CUDA device current (nothing for non-CUDA tensors)."""
    import contextlib
    if isinstance(t, torch.Tensor) and t.is_cuda:
        return torch.cuda.device(t.device)
    return contextlib.nullcontext()


def on_tensor_device(fn):
    """Decorator for the public entry points: run ``fn`` with the CUDA runtime's current device set
    to the device of its first CUDA tensor argument.  The C-ABI launches on the CURRENT device and
    ``stream()`` returns the current device's stream, so a model on cuda:1 called while cuda:0 is
    current would otherwise launch on the wrong device with foreign pointers (autograd's backward
    thread already switches to the tensors' device by itself)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kw):
        for a in list(args) + list(kw.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kw)
                break
        return fn(*args, **kw)
    return wrapper


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: statecatcher_b200 runs on sm_100a CUDA devices only "
            "(no CPU fallback; the CPU oracle lives in oracle/ and is test-only)")


# ---- instrumentation read by bench.py ------------------------------------------------
# launches = C-ABI calls made; kernels = CUDA kernels those calls enqueue (each entry point
# launches a fixed, documented number of kernels); profile = None, or a list that receives
# (name, algorithmic work, start event, end event) per call for the roofline accounting.
launches = 0
kernels = 0
profile = None

KERNELS_PER_CALL = {
    "sc_gemm_fwd": 1, "sc_gemm_dgrad": 1, "sc_gemm_wgrad": 1, "sc_cast": 1, "sc_colsum": 2,
    "sc_layernorm_fwd": 1, "sc_layernorm_bwd": 1, "sc_lucy_scan_fwd": 1, "sc_lucy_scan_fwd_chunked": 3, "sc_lucy_scan_bwd": 1,
    "sc_lucy_sscan_fwd": 1, "sc_lucy_sscan_bwd": 1, "sc_lucy_hscan_fwd": 1, "sc_lucy_hscan_bwd": 1,
    "sc_ctc_fwd": 5, "sc_ctc_emissions": 1, "sc_ctc_lattice": 4, "sc_ctc_bwd": 1, "sc_ctc_scale_grad": 1, "sc_rnnt_fwd": 2, "sc_rnnt_bwd": 2, "sc_split_bf16": 1, "sc_joint_fwd": 1, "sc_joint_bwd": 2, "sc_rnnt_lse_gather": 1, "sc_rnnt_lattice": 1,
    "sc_rnnt_node_grads": 1, "sc_rnnt_dlogits": 1, "sc_ctc_greedy_decode": 2, "sc_frontend": 1, "sc_frame_mask": 1,
}
_ESZ = {SC_F32: 4, SC_BF16: 2}


def _work(name, a):
    """Algorithmic work of one call: FLOPs for the projections, HBM bytes for the scans
    (fwd 6H*e, bwd 12H*e bytes per frame and layer)."""
    if name == "sc_gemm_fwd":
        return 2.0 * a[7] * a[8] * a[9]
    if name in ("sc_gemm_dgrad", "sc_gemm_wgrad"):
        return 2.0 * a[6] * a[7] * a[8]
    if name == "sc_lucy_scan_fwd":
        return 6.0 * a[11] * _ESZ[a[12]] * a[9] * a[10]
    if name == "sc_lucy_scan_bwd":
        return 12.0 * a[14] * _ESZ[a[15]] * a[12] * a[13]
    return 0.0


def _shape(name, a):
    if name == "sc_gemm_fwd":
        return (a[7], a[8], a[9])
    if name in ("sc_gemm_dgrad", "sc_gemm_wgrad"):
        return (a[6], a[7], a[8])
    return ()


def call(name: str, *args):
    global launches, kernels
    launches += 1
    kernels += KERNELS_PER_CALL.get(name, 1)
    if name == "sc_ctc_head":        # emission, recursion and gradient launch per phase + check, recomputation, reduction, fix-up
        ph = load().sc_ctc_head_phases(args[9], args[11], args[26]) if args[29] and args[30] else 1
        kernels += 3 * ph + 4 - 1 if ph >= 2 else 6 - 1
    if name == "sc_gemm_wgrad" and args[9] == SC_BF16 and not args[10]:
        kernels += 1                 # the tcgen05 split-R wgrad zero-fills dW first (zero2d_kernel) when not accumulating
    fn = getattr(load(), name)
    if profile is None:
        check(fn(*args), name)
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(*args)
    e1.record()
    check(rc, name)
    profile.append((name, _work(name, args), e0, e1, _shape(name, args)))
