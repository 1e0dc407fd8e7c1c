"""ctypes binding of include/clairs_to_b200.h.  There is no CPU fallback: a missing library is an error."""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcto_b200.so")

_lib = None
ABI_VERSION = 4       # include/clairs_to_b200.h CTO_ABI_VERSION


class HostStream(C.Structure):
    _fields_ = [("planes", C.c_void_p), ("grp_off", C.c_void_p), ("ref_code", C.c_void_p), ("ind_off", C.c_void_p),
                ("ind_entry", C.c_void_p), ("win_pos", C.c_void_p), ("n_groups", C.c_int64), ("n_rows", C.c_int64),
                ("n_ind", C.c_int64)]


P, I32, I64, INT = C.c_void_p, C.c_int32, C.c_int64, C.c_int

SIGNATURES = {
    "cto_abi_version": (INT, []),
    "cto_last_error": (C.c_char_p, []),
    "cto_device_check": (INT, [P]),
    "cto_encode_pileup": (INT, [P, P, P, P, P, P, I64, I64, P, P, P]),
    "cto_pack_reads": (INT, [P, P, P, P, I64, INT, P, P, INT]),
    "cto_engine_create": (INT, [P, I64, P, INT, P, I64, P, INT, I64, P]),
    "cto_engine_destroy": (None, [P]),
    "cto_engine_heads": (INT, [P]),
    "cto_engine_set_likelihood": (INT, [P, P, INT]),
    "cto_rescale": (INT, [P, P, I64, P, P]),
    "cto_forward_aff": (INT, [P, P, I64, P, P]),
    "cto_forward_neg": (INT, [P, P, I64, P, P]),
    "cto_softmax_posterior": (INT, [P, P, P, I64, P, P, P, P, P, P]),
    "cto_engine_set_qual_thresholds": (INT, [P, C.c_double, C.c_double, C.c_double]),
    "cto_engine_set_tensor_cores": (INT, [P, INT]),
    "cto_engine_set_overlap": (INT, [P, INT]),
    "cto_engine_fused_status": (INT, [P, P]),
    "cto_aff_stage_layers": (INT, [P, INT, P, I64, P]),
    "cto_neg_recurrence": (INT, [P, P, I64, P, P, INT, P]),
    "cto_engine_workspace": (INT, [P, INT, P, P]),
    "cto_gemm_nt": (INT, [P, I64, P, P, P, I64, P, I64, I64, INT, INT, INT, INT, P]),
    "cto_posterior_from_probs": (INT, [P, INT, P, P, I64, P, P, P]),
    "cto_launch_count": (I64, []),
    "cto_debug_set": (None, [INT]),
    "cto_debug_timing": (None, [P]),
    "cto_debug_timing_fused": (None, [P]),
    "cto_engine_profile": (INT, [P, INT]),
    "cto_engine_profile_kinds": (INT, []),
    "cto_engine_profile_name": (C.c_char_p, [INT]),
    "cto_engine_profile_read": (INT, [P, P, P, P]),
    "cto_strand_counts": (INT, [P, I64, P, P, P]),
    "cto_predict": (INT, [P, P, P, P, P, I64, P, P, P, P, P, P, P, P, P, P]),
    "cto_run_sites_host": (INT, [P, P, P, I64, P, P, P, P, P, P, P, P]),
    "cto_tokenize_mpileup": (INT, [C.c_char_p, I64, C.c_char_p, I64, I64, P, I64, INT, INT, P]),
    "cto_tokens_sizes": (INT, [P, P, P, P, P]),
    "cto_tokens_export": (INT, [P, P, P, P, P, P, P, P, P, P, P]),
    "cto_tokens_destroy": (None, [P]),
    "cto_render_mpileup": (I64, [P, P, P, P, I64, P, P, P, C.c_char_p, I64, P, I64]),
    "cto_format_tensor_row": (I64, [P, P, I64]),
    "cto_format_prob_fields": (I64, [P, INT, P, I64]),
    "cto_parse_tensor_row": (INT, [C.c_char_p, I64, P]),
    "cto_parse_tensor_file": (INT, [C.c_char_p, I64, I64, P, P, P, P]),
    "cto_format_tensor_can_rows": (I64, [C.c_char_p, I64, I64, P, C.c_char_p, P, C.c_char_p, P, P, P, I64]),
    "cto_format_predict_rows": (I64, [C.c_char_p, P, I64, P, P, P, INT, P, I64]),
    "cto_parse_predict_file": (INT, [C.c_char_p, I64, INT, I64, P, P, P, P]),
    "cto_index_rows": (INT, [P, I64, P, I64, P, P]),
    "cto_scan_candidates": (INT, [P, I64, P, I64, P, I64, I64, C.c_double, C.c_double, C.c_double, INT, INT, P, P, P, P, P]),
    "cto_tokenize_count": (INT, [P, I64, P, I64, P, I64, I64, P, P, P, P, P, P, P]),
    "cto_tokenize_write": (INT, [P, I64, P, I64, P, I64, I64, INT, INT, P, P, P, P, P]),
    "cto_window_table": (INT, [P, I64, P, I64, P, P]),
    "cto_hf_parse": (INT, [C.c_char_p, I64, INT, C.c_char_p, I64, I64, P]),
    "cto_hf_parse_mt": (INT, [C.c_char_p, I64, INT, C.c_char_p, I64, I64, INT, P]),
    "cto_hf_sizes": (INT, [P, P]),
    "cto_hf_export": (INT, [P] * 15),
    "cto_hf_free": (None, [P]),
    "cto_hard_filter_sites": (INT, [P, P, INT, INT, INT, INT, P, C.c_double, C.c_double, P, P, P, P, P]),
    "cto_scan_candidates_host": (INT, [P, I64, P, I64, I64, C.c_double, C.c_double, C.c_double, INT, INT, I64, P, P, P, P, P, P]),
}


class CtoError(RuntimeError):
    pass


def lib():
    """Load libcto_b200.so (built in-tree by clairs_to_b200.build); raise if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CtoError("%s not found: run `python -m clairs_to_b200.build` (nvcc, sm_100a). "
                           "There is no CPU fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.cto_abi_version() != ABI_VERSION:
            raise CtoError("%s has ABI version %d, this package binds version %d: rebuild with "
                           "`python -m clairs_to_b200.build --force`" % (LIB_PATH, handle.cto_abi_version(), ABI_VERSION))
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().cto_last_error()
        raise CtoError("%s failed (%d): %s" % (what or "clairs_to_b200 call", rc, msg.decode() if msg else "?"))
