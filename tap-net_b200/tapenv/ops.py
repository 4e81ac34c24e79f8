"""pack.update_dynamic / pack.update_mask with the reference's signatures (pack.py:333-376, :276-331),
executed by the sm_100a kernels behind the C ABI.  CUDA tensors only -- there is no CPU path."""
import ctypes as C

import torch

from . import _capi
from .config import tensor_config


# The reference's gather / scatter raise IndexError for a pointer outside [0, S) (pack.py:347, :318-321); the kernels cannot
# raise, and silently clamping would corrupt the precedence state.  With CHECK_POINTERS (default) update_dynamic / update_mask
# test the range first -- one tiny reduction and a host sync, negligible beside the syncs an unmodified model.py performs per
# step.  The fused paths (BatchedContainers.step, DecodeLoop) never sync: they set sticky flag 4 instead (check_flags()).
CHECK_POINTERS = True


def _check_ptr(ptr, S, who):
    if CHECK_POINTERS and ptr.numel() and bool(((ptr < 0) | (ptr >= S)).any()):
        raise IndexError("tapenv.%s: chosen_idx outside [0, %d)" % (who, S))


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(t, name, dtype):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("tapenv: %s must be a CUDA tensor (no CPU fallback exists)" % name)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def update_dynamic(dynamic, static, chosen_idx, input_type, allow_rot):
    """Out-of-place: returns a new [B, rows, S] tensor with the rows of the chosen block zeroed in
    every band (pack.py:370-374).  The block id is read from static[:,0,ptr] (pack.py:347)."""
    dynamic = _dev(dynamic, "dynamic", torch.float32)
    static = _dev(static, "static", torch.float32)
    ptr = _dev(chosen_idx, "chosen_idx", torch.int64)
    B, rows, S = dynamic.shape
    cfg = tensor_config(B, static.shape[1], rows, S, input_type, allow_rot)
    _check_ptr(ptr, S, "update_dynamic")
    out = torch.empty_like(dynamic)
    with torch.cuda.device(dynamic.device):
        _capi.check(_capi.lib.tapenv_update_dynamic(C.byref(cfg), _p(dynamic), _p(static), _p(ptr), _p(out), _stream()),
                    "update_dynamic")
    return out


def update_mask(mask, dynamic, static, chosen_idx, input_type, allow_rot):
    """Returns (new_mask, chosen_mask) (pack.py:276-331): chosen_mask = mask with every rotation of the
    chosen block cleared; new_mask = chosen_mask restricted to the accessible candidates of the
    (already updated) `dynamic`."""
    mask = _dev(mask, "mask", torch.float32)
    dynamic = _dev(dynamic, "dynamic", torch.float32)
    ptr = _dev(chosen_idx, "chosen_idx", torch.int64)
    B, rows, S = dynamic.shape
    cfg = tensor_config(B, static.shape[1], rows, S, input_type, allow_rot)
    _check_ptr(ptr, S, "update_mask")
    new_mask = torch.empty_like(mask)
    chosen = torch.empty_like(mask)
    with torch.cuda.device(dynamic.device):
        _capi.check(_capi.lib.tapenv_update_mask(C.byref(cfg), _p(mask), _p(dynamic), _p(ptr), _p(new_mask), _p(chosen),
                                                 _stream()), "update_mask")
    return new_mask, chosen
