"""Python front-ends of the hand-written sm_100a kernels (``edl_b200/csrc``).

Dispatch rule: CUDA tensors ALWAYS go to the native extension (and fail loudly if it is missing --
a silent eager fallback on a GPU box would hide that the product path is not running); CPU tensors
use small pure-PyTorch reference implementations so the whole stack is testable without a GPU.
"""
from __future__ import annotations

import importlib
import threading

_lock = threading.Lock()
_C = None
_launches = 0  # number of native kernel launches issued through this module (bench bookkeeping)


def native():
    """Return the compiled extension module, importing it on first use."""
    global _C
    if _C is None:
        with _lock:
            if _C is None:
                try:
                    import torch  # noqa: F401  (must be loaded before the extension)

                    _C = importlib.import_module("edl_b200._C")
                except ImportError as e:  # pragma: no cover - exercised only on broken installs
                    raise RuntimeError(
                        "edl_b200 native extension is not built; run `python -m edl_b200.build_ext` "
                        "(or __graft_entry__.build()). Original error: %r" % (e,)
                    )
    return _C


def native_available() -> bool:
    try:
        native()
        return True
    except RuntimeError:
        return False


def count_launch(n: int = 1) -> None:
    global _launches
    _launches += n


def launches() -> int:
    return _launches


def reset_launches() -> None:
    global _launches
    _launches = 0


from .bn import batch_norm_act, BatchNormAct2d, scale_shift_act, bn_stats_into, set_fused_bn, drop_bn_hook  # noqa: E402
from .loss import soft_cross_entropy, topk_accuracy  # noqa: E402
from .pool import max_pool_3x3_s2, avg_pool_2x2, global_avg_pool  # noqa: E402
from .optim import FlatSGDMomentum, FlatAdam  # noqa: E402
from .gemm import (gemm_bf16, linear_bf16, conv1x1, conv_lib, conv3x3, conv3x3_supported, conv3x3_infer,  # noqa: E402
                   conv3x3_infer_supported, conv3x3_wgrad, conv3x3_wgrad_supported, conv3x3_s2, conv3x3_s2_infer,
                   conv3x3_s2_supported, stem_conv, stem_conv_supported)
from .misc import rope, rope_tables, embedding_bag_mean, normalize_u8, DynamicLossScaler  # noqa: E402

__all__ = [
    "native", "native_available", "launches", "reset_launches",
    "batch_norm_act", "BatchNormAct2d", "scale_shift_act", "bn_stats_into",
    "soft_cross_entropy", "topk_accuracy",
    "max_pool_3x3_s2", "avg_pool_2x2", "global_avg_pool",
    "FlatSGDMomentum", "FlatAdam", "gemm_bf16", "linear_bf16", "conv1x1", "conv_lib", "conv3x3",
    "conv3x3_supported", "conv3x3_infer", "conv3x3_infer_supported",
    "conv3x3_wgrad", "conv3x3_wgrad_supported", "conv3x3_s2", "conv3x3_s2_infer", "conv3x3_s2_supported",
    "stem_conv", "stem_conv_supported",
    "rope", "rope_tables", "embedding_bag_mean", "normalize_u8", "DynamicLossScaler", "set_fused_bn", "drop_bn_hook",
]
