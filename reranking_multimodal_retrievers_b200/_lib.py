"""ctypes binding of libplaid_b200.so (the C ABI declared in include/plaid_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc (sm_100a cross-compile);
if that is impossible the import fails loudly.  Every call checks the int return code and raises
``PlaidError`` with ``plaid_last_error()``.
"""
from __future__ import annotations

import ctypes
import os
import re

from . import build as _build

_P = ctypes.c_void_p
_I = ctypes.c_int
_I64 = ctypes.c_int64
_F = ctypes.c_float

# name -> argtypes (all functions return int unless listed in _RESTYPES)
_SIGNATURES = {
    "plaid_abi_version": [],
    "plaid_last_error": [],
    "plaid_arch": [],
    "plaid_merge_cells": [_P, _P, _P, _I, _I, _I, _P, _P],
    "plaid_compress_residuals": [_P, _P, _P, _P, ctypes.c_int64, _I, _I, _P, _P, _P],
    "plaid_prepare_queries": [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "plaid_f32_to_bf16": [_P, _P, _I64, _P],
    "plaid_centroid_scores": [_P, _I, _P, _P, _I, _I, _F, _I, _I, _P, _I, _P, _P, _P, _P, _P],
    "plaid_candidates": [_P, _P, _P, _I, _I, _I, _P, _P, _I, _I, _P, _P, _P, _P, _I, _P, _P, _P],
    "plaid_approx_scores": [_P, _P, _I, _I, _P, _I, _P, _P, _I, _P, _P, _P, _P],
    "plaid_filter_stage1_ivf": [_P, _P, _I, _I, _P, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _I, _P, _I, _P, _P, _P, _I, _P, _P,
                                _P],
    "plaid_set_ivf_range_slots": [_I],
    "plaid_select_top": [_P, _P, _P, _I, _I, _I, _P, _P, _P, _I, _P, _P],
    "plaid_filter_pids": [_P, _P, _I, _I, _P, _I, _P, _P, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "plaid_build_weight_table": [_P, _P, _P, _I, _P, _P],
    "plaid_decompress_residuals": [_P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P],
    "plaid_unpack_residual_codes": [_P, _I64, _I, _P, _P, _P, _P],
    "plaid_decompress_tokens_f16": [_P, _P, _I64, _P, _P, _I, _I, _I, _P, _P],
    "plaid_packbits": [_P, _I64, _P, _P],
    "plaid_doc_token_offsets": [_P, _P, _I, _I, _P, _I, _P, _P],
    "plaid_decompress_normalize_bf16": [_P, _P, _I, _I, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P],
    "plaid_decompress_normalize_f16": [_P, _P, _I, _I, _P, _I, _P, _P, _P, _P, _P, _I, _I, _P, _P],
    "plaid_maxsim_packed": [_P, _P, _I, _I, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    "plaid_maxsim_fused": [_P, _P, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P],
    "plaid_token_inv_norms": [_P, _P, _I64, _P, _P, _I, _I, _P, _P],
    "plaid_segmented_maxsim": [_P, _I, _P, _P, _I, _P, _P],
    "plaid_colbert_score_padded": [_P, _P, _I, _I, _I, _P, _P, _I64, _I, _I, _P, _P, _I, _P, _P],
    "plaid_colbert_score_backward": [_P, _P, _I, _I, _P, _P, _I64, _I, _I, _I, _P, _P, _P, _P, _P],
    "plaid_colbert_score_reduce": [_P, _P, _I64, _I, _I, _P, _P],
    "plaid_merge_topk": [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P],
    "plaid_merge_topk_msg": [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "plaid_prefix_share": [_P, _I, _I, _I, _I, _I, _P, _I, _P, _P],
    "plaid_segmented_lookup": [_P, _I64, _P, _P, _P, _I, _P, _P],
}
_RESTYPES = {"plaid_last_error": ctypes.c_char_p, "plaid_arch": ctypes.c_char_p}


class PlaidError(RuntimeError):
    """A libplaid_b200 call returned a negative status."""


def header_path() -> str:
    return _build.HEADER


def declared_symbols() -> list[str]:
    """Function names declared in include/plaid_b200.h."""
    with open(header_path()) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(plaid_[a-z0-9_]+)\s*\(", text)))


_LIB = None


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = os.environ.get("PLAID_B200_LIB") or _build.LIB_PATH     # override: A/B timing of kernel variants
        if not os.path.exists(path):
            path = _build.build_library()  # raises if nvcc is unavailable: no silent fallback
        handle = ctypes.CDLL(path)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        _LIB = handle
    return _LIB


def call(name: str, *args):
    fn = getattr(lib(), name)
    rc = fn(*args)
    if rc != 0:
        msg = lib().plaid_last_error()
        raise PlaidError(f"{name} failed with status {rc}: {msg.decode() if msg else ''}")
    return rc
