"""ctypes binding of the C-ABI library (include/otpose_b200.h).

The CUDA library is the product: there is no PyTorch / CPU fallback.  If the
shared object is missing the import of any compute entry point fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libotpose_b200.so")

PREC_FP32, PREC_BF16, PREC_FP16 = 0, 1, 2
_PREC = {"fp32": PREC_FP32, "bf16": PREC_BF16, "fp16": PREC_FP16}

vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t


class BlockParams(C.Structure):
    """otp_block_params (include/otpose_b200.h)."""
    _fields_ = [(n, vp) for n in (
        "ln1_w", "ln1_b", "ln2_w", "ln2_b", "q_conv_w", "k_conv_w", "v_conv_w",
        "q_norm_w", "q_norm_b", "k_norm_w", "k_norm_b", "v_norm_w", "v_norm_b",
        "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "proj_w", "proj_b",
        "mlp0_w", "mlp0_b", "mlp3_w", "mlp3_b", "scale_attn", "scale_mlp")]


# name -> (restype, argtypes); every symbol include/otpose_b200.h declares
SIGNATURES = {
    "otp_version": (C.c_char_p, []),
    "otp_last_error": (C.c_char_p, []),
    "otp_device_is_sm100": (i32, []),
    "otp_has_tensor_core_path": (i32, []),
    "otp_launch_count": (C.c_ulonglong, []),
    "otp_profile_num_kernels": (i32, []),
    "otp_profile_kernel_name": (C.c_char_p, [i32]),
    "otp_profile_enable": (i32, [i32]),
    "otp_profile_read": (i32, [C.POINTER(C.c_float), C.POINTER(C.c_int), i32]),
    "otp_final_preds": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
    "otp_mdcn_forward": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32,
                               i32, i32, f32, i32, vp]),
    "otp_mdcn_backward_workspace_bytes": (sz, [i32] * 10),
    "otp_mdcn_backward": (i32, [vp] * 10 + [i32] * 12 + [vp, sz, vp]),
    "otp_fusion_sum": (i32, [vp, i32, i32, i32, vp, vp, vp]),
    "otp_fusion_stack": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
    "otp_fusion_sum_frames": (i32, [vp, i32, i32, i32, i32, vp, vp, vp]),
    "otp_fusion_stack_frames": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
    "otp_block_packed_bytes": (sz, [i32, i32]),
    "otp_block_pack": (i32, [C.POINTER(BlockParams), i32, i32, vp, sz, vp]),
    "otp_block_workspace_bytes": (sz, [i32, i32, i32, i32, i32, i32]),
    "otp_block_forward": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, sz, vp]),
    "otp_flow_encoder_supported": (i32, [i32, i32, i32, i32]),
    "otp_flow_encoder_workspace_bytes": (sz, [i32, i32]),
    "otp_flow_encoder_forward": (i32, [vp, i32, vp, vp, i32, vp, i32, i32, vp, sz, vp]),
    "otp_add_pos_embd": (i32, [vp, vp, i32, vp, i32, i32, i32, vp]),
    "otp_upsample_linear": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "otp_pyramid_conv1x1": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, i32, vp, i64, vp]),
    "otp_pyramid_conv1x1_tc_supported": (i32, [i32, i32, i32]),
    "otp_pyramid_conv1x1_tc_pack_bytes": (sz, [i32]),
    "otp_pyramid_conv1x1_tc_pack": (i32, [vp, i32, i32, i32, vp, sz, vp]),
    "otp_pyramid_conv1x1_tc": (i32, [vp, vp, vp, i32, i32, i32, vp, vp, i32, vp, i64, i32, vp]),
    "otp_final_layer_fusion_sum": (i32, [vp, i32, i32, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
    "otp_conv2d": (i32, [vp, i64, vp, i64, vp, vp, vp, i64, vp, i64, i32, i32, i32, i32, i32, i32, i32,
                         i32, vp]),
    "otp_conv2d_wgrad_workspace_bytes": (sz, [i32, i32, i32, i32, i32, i32]),
    "otp_conv2d_wgrad": (i32, [vp, i64, vp, i64, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, sz, vp]),
    "otp_conv_bn_fold": (i32, [vp, vp, vp, vp, vp, vp, f32, i32, i32, i32, vp, vp, vp]),
    "otp_window_assemble": (i32, [vp, i32, i32, i32, i64, vp, i32, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
    "otp_rsb_block_supported": (i32, [i32, i32, i32, i32]),
    "otp_rsb_block_pack_bytes": (sz, [i32, i32, i32]),
    "otp_rsb_block_pack": (i32, [vp, vp, i32, i32, i32, vp, sz, vp]),
    "otp_rsb_block_workspace_bytes": (sz, [i32, i32, i32, i32, i32]),
    "otp_rsb_block_forward": (i32, [vp, vp, i64, vp, i64, i32, i32, i32, i32, i32, i32, vp, sz, vp]),
    "otp_conv2d_tc_supported": (i32, [i32, i32, i32, i32, i32]),
    "otp_conv2d_tc_pack_bytes": (sz, [i32, i32, i32]),
    "otp_conv2d_tc_pack": (i32, [vp, i32, i32, i32, i32, vp, sz, vp]),
    "otp_conv2d_tc": (i32, [vp, i64, vp, i64, vp, vp, vp, i64, vp, i64, i32, i32, i32, i32, i32, i32, i32, i32,
                            vp]),
    "otp_offset_mask_pack_bytes": (sz, []),
    "otp_offset_mask_pack": (i32, [vp, vp, i32, i32, vp, sz, vp]),
    "otp_offset_mask_dcn_forward": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, i32, i32, vp]),
    "otp_debug_trace": (i32, [i32]),
    "otp_debug_umma_rate": (i32, [i32, i32, i32, i32, C.POINTER(C.c_longlong)]),
    "otp_debug_trace_read": (i32, [C.POINTER(C.c_ulonglong), i32]),
    "otp_debug_umma_gemm": (i32, [vp, i32, vp, i32, vp, i32, i32] + [C.c_uint] * 8 + [i32, i32, i32, vp]),
}

_lib = None


def load():
    """Load (once) and return the ctypes handle; raises if the library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension is the only implementation of this "
                "package (no CPU / PyTorch fallback).  Build it with `python -m otpose_b200.build` "
                "or `python -c 'import __graft_entry__ as g; g.build()'`.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().otp_last_error().decode(errors="replace")
        kind = {1: "bad argument", 2: "unsupported", 3: "workspace", 4: "CUDA error"}.get(status, "error")
        if status == 2:
            raise NotImplementedError(f"otpose_b200 {what}: {msg}")
        raise RuntimeError(f"otpose_b200 {what}: {kind} ({status}): {msg}")


def precision_code(p) -> int:
    if isinstance(p, int):
        return p
    try:
        return _PREC[p]
    except KeyError:
        raise ValueError(f"precision must be 'fp32', 'bf16' or 'fp16', got {p!r}") from None


def dptr(t, dtype=torch.float32, allow_none=False):
    """Device pointer of a contiguous CUDA tensor (None -> NULL when allowed)."""
    if t is None:
        if allow_none:
            return None
        raise ValueError("tensor required")
    if not t.is_cuda:
        # the reference's DCN op raises NotImplementedError for non-CUDA input
        # (thirdparty/deform_conv/functions/deform_conv.py:131-132); same contract here
        raise NotImplementedError("otpose_b200 kernels need CUDA tensors (no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("tensor has to be contiguous")
    return t.data_ptr()


def require_cuda(t):
    """Same contract as the reference's DCN op (functions/deform_conv.py:131-132)."""
    if not t.is_cuda:
        raise NotImplementedError("otpose_b200 kernels need CUDA tensors (no CPU path)")


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def profile_read():
    """{kernel name: (total device ms, recorded launches)} since otp_profile_enable(1)."""
    lib = load()
    n = lib.otp_profile_num_kernels()
    ms = (C.c_float * n)()
    cnt = (C.c_int * n)()
    check(lib.otp_profile_read(ms, cnt, n), "otp_profile_read")
    return {lib.otp_profile_kernel_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i]}


class _WorkspacePool:
    """Scratch buffers for the kernels' workspaces.

    Eager calls: one growing buffer per (device, stream, tag); kernels are stream-ordered, so consecutive
    calls on the same stream share it and branches running concurrently on different streams (the two
    temporal encoders) never do.  A regrown buffer's predecessor goes back to the caching allocator, which
    re-issues it only to later work on the same stream.

    During CUDA-graph capture the pool is bypassed: the buffer is a plain ``torch.empty`` made inside the
    capture, i.e. it lives in the capturing graph's private memory pool for as long as that graph exists.
    A captured graph therefore never holds a pointer into a pooled buffer that a later, larger eager call
    (or the capture of another shape) could regrow and free.
    """

    def __init__(self):
        self._bufs = {}

    def get(self, nbytes: int, device, tag: str = "ws") -> torch.Tensor:
        if torch.cuda.is_current_stream_capturing():
            return torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        key = (torch.device(device).index, torch.cuda.current_stream(device).cuda_stream, tag)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf

    def drop_stream(self, device, stream) -> None:
        """Forget the buffers keyed by a stream that is going away (a graph warm-up stream)."""
        dev, ptr = torch.device(device).index, stream.cuda_stream
        for key in [k for k in self._bufs if k[0] == dev and k[1] == ptr]:
            del self._bufs[key]

    def clear(self):
        self._bufs.clear()


workspace = _WorkspacePool()
