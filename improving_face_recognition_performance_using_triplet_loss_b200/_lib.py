"""ctypes binding of libefmtl.so (the C ABI declared in include/efmtl.h).

There is no CPU fallback: importing this module without a loadable libefmtl.so raises, and every
operator refuses non-CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libefmtl.so")

F32, BF16, F16 = 0, 1, 2
ACT_EFM3, ACT_MFM2 = 3, 2
BH_NORMALIZE, BH_DETACH_GATHERED, BH_MINE_BF16 = 1, 2, 4
GAL_POSITIVE_ONLY = 1
GAL_MAX_K = 16

_DTYPES = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}

_vp, _l, _i, _f, _sz = C.c_void_p, C.c_long, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/efmtl.h declares
SIGNATURES = {
    "efm_last_error": (C.c_char_p, []),
    "efm_abi_version": (_i, []),
    "efm_device_arch": (_i, []),
    "efm_act_fwd": (_i, [_i, _vp, _vp, _l, _l, _l, _i, _vp]),
    "efm_act_bwd": (_i, [_i, _vp, _vp, _vp, _l, _l, _l, _i, _vp]),
    "efm_l2norm_fwd": (_i, [_vp, _vp, _vp, _l, _l, _i, _vp]),
    "efm_l2norm_bwd": (_i, [_vp, _vp, _vp, _l, _l, _i, _vp]),
    "efm_sumsq": (_i, [_vp, _l, _i, _vp, _vp]),
    "efm_frob_fwd": (_i, [_vp, _vp, _vp, _l, _i, _vp]),
    "efm_frob_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _l, _i, _vp]),
    "efm_dot": (_i, [_vp, _vp, _l, _i, _vp, _vp]),
    "efm_triplet_fwd": (_i, [_vp, _vp, _vp, _l, _l, _i, _f, _vp, _vp]),
    "efm_triplet_bwd": (_i, [_vp, _vp, _vp, _vp, _l, _l, _i, _f, _vp, _vp, _vp, _vp]),
    "efm_cosine_pairs": (_i, [_vp, _vp, _l, _l, _i, _vp, _vp]),
    "efm_cosine_pairs_indexed": (_i, [_vp, _vp, _vp, _l, _l, _i, _vp, _vp]),
    "efm_cosine_allpairs_workspace_bytes": (_sz, [_l, _l, _l, _i]),
    "efm_cosine_allpairs": (_i, [_vp, _vp, _l, _l, _l, _i, _vp, _l, _vp, _sz, _vp]),
    "efm_bh_workspace_bytes": (_sz, [_l, _l, _l, _i, _i]),
    "efm_bh_triplet_fwd_bwd": (_i, [_vp, _vp, _l, _l, _l, _l, _i, _f, _vp, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "efm_bh_mine": (_i, [_vp, _l, _l, _l, _l, _i, _i, _vp, _sz, _vp]),
    "efm_bh_backward_workspace_bytes": (_sz, [_l]),
    "efm_bh_backward": (_i, [_vp, _l, _l, _l, _l, _i, _f, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "efm_bh_grad_rows_workspace_bytes": (_sz, [_l, _l]),
    "efm_bh_grad_rows": (_i, [_vp, _l, _l, _i, _i, _vp, _vp, _vp, _l, _l, _vp, _vp, _sz, _vp]),
    "efm_bh_xchg_bytes": (_sz, [_l, _i, _l, _i]),
    "efm_bh_dist_workspace_bytes": (_sz, [_l, _i]),
    "efm_xchg_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "efm_xchg_free": (_i, [_vp]),
    "efm_ipc_export": (_i, [_vp, C.c_char_p]),
    "efm_ipc_import": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "efm_ipc_release": (_i, [_vp]),
    "efm_bh_dist_fwd_bwd": (_i, [_vp, _vp, _l, _l, _i, _f, _vp, _i, _i, _i, C.POINTER(_vp), C.c_ulonglong, _vp, _vp, _vp, _vp,
                                 _vp, _sz, _vp]),
    "efm_bh_profile_events": (_i, [C.POINTER(_vp), _i]),
    "efm_gallery_scores_f32": (_i, [_vp, _vp, _l, _l, _vp, _l, _vp, _vp]),
    "efm_simd_dot_rows": (_i, [_vp, _l, _l, _vp, _l, _vp, _vp]),
    "efm_simd_norm_rows": (_i, [_vp, _l, _l, _vp, _vp]),
    "efm_gallery_topk_workspace_bytes": (_sz, [_l, _l, _l, _i, _i]),
    "efm_gallery_topk": (_i, [_vp, _vp, _l, _l, _vp, _l, _i, _f, _i, _i, _l, _vp, _vp, _vp, _sz, _vp]),
    "efm_topk_merge": (_i, [_vp, _vp, _i, _l, _i, _vp, _vp, _vp]),
    "efm_gather_pid": (_i, [_vp, _vp, _l, _vp, _vp]),
    "efm_histogram": (_i, [_vp, _l, _vp, _i, _vp, _vp]),
}


def _load():
    if _build._stale():
        # (re)build in-tree when a toolchain is present; otherwise fail loudly (no CPU fallback exists)
        try:
            _build.build_library()
        except Exception as e:  # noqa: BLE001
            if os.path.exists(LIB_PATH):
                raise ImportError("libefmtl.so is older than its sources and the rebuild failed: %s" % e)
            raise ImportError(
                "libefmtl.so is missing and could not be built (%s). This package has no CPU fallback; "
                "run `python -m improving_face_recognition_performance_using_triplet_loss_b200.build`." % e)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class EfmError(RuntimeError):
    pass


def last_error() -> str:
    msg = lib.efm_last_error()
    return msg.decode() if msg else ""


def check(rc: int):
    if rc == 0:
        return
    msg = "libefmtl error %d: %s" % (rc, last_error())
    if rc in (-1, -2):
        raise ValueError(msg)
    raise EfmError(msg)


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise ValueError("unsupported dtype %s (float32, bfloat16, float16)" % t.dtype) from None


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ValueError("libefmtl operators run on CUDA tensors only (got a %s tensor); there is no CPU path" % t.device)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def workspace(nbytes: int, device) -> torch.Tensor:
    """256-byte aligned scratch (torch's caching allocator aligns to 512 B)."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
