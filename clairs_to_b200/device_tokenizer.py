"""mpileup text -> the encoder's packed input on the GPU: Python side of ``cto_index_rows`` +
``cto_tokenize_count`` + ``cto_tokenize_write`` + ``cto_window_table`` (csrc/tokenize_dev.cu).

Same arrays as the host pair ``host.tokenize_mpileup`` + ``pileup_format.pack_stream`` (which restate
src/create_tensor_pileup_calling.py:472-497, 120-144), byte for byte -- but the host only copies the text.  The alt_info
strings (ibid. 158-209) are not produced: this is the path for callers that want tensors / probabilities, not tensor_can files.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .pileup_format import PackedStream


def _p(t):
    return C.c_void_p(t.data_ptr())


def text_to_device(text, device, stream=None) -> tuple:
    """bytes / pinned uint8 tensor -> (device uint8 tensor allocated 32 bytes beyond the text, length)."""
    if isinstance(text, (bytes, bytearray, memoryview)):
        text = torch.frombuffer(bytearray(text), dtype=torch.uint8) if len(text) else torch.zeros(0, dtype=torch.uint8)
    n = int(text.numel())
    buf = torch.empty(n + 32, dtype=torch.uint8, device=device)
    buf[n:].zero_()
    if n:
        buf[:n].copy_(text, non_blocking=True)
    return buf, n


def tokenize_text_device(text_dev: torch.Tensor, n_bytes: int, ref_dev: torch.Tensor, ref_start: int, low_bq_cut: int,
                         cand_pos_dev: torch.Tensor = None, max_indel_length: int = 60):
    """Text in HBM -> (device PackedStream, row positions int32).  ``ref_dev``: uint8 reference bases of positions
    ``ref_start ..``; ``cand_pos_dev``: int64 candidate positions (window table; None = no windows yet)."""
    lib = _lib.lib()
    dev = text_dev.device
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    n = C.c_int64()
    _lib.check(lib.cto_index_rows(_p(text_dev), n_bytes, None, 0, C.byref(n), stream), "cto_index_rows")
    n_rows = n.value
    row_off = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
    _lib.check(lib.cto_index_rows(_p(text_dev), n_bytes, _p(row_off), n_rows, C.byref(n), stream), "cto_index_rows")
    row_pos = torch.empty(max(n_rows, 1), dtype=torch.int32, device=dev)
    ref_code = torch.empty(max(n_rows, 1), dtype=torch.uint8, device=dev)
    grp_off = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
    ind_off = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
    n_groups, n_ind = C.c_int64(), C.c_int64()
    _lib.check(lib.cto_tokenize_count(_p(text_dev), n_bytes, _p(row_off), n_rows, _p(ref_dev), int(ref_start), int(ref_dev.numel()),
                                      _p(row_pos), _p(ref_code), _p(grp_off), _p(ind_off), C.byref(n_groups), C.byref(n_ind), stream),
               "cto_tokenize_count")
    plane_bytes = ((n_groups.value * 8 + 15) & ~15) + 16
    planes = torch.empty(plane_bytes, dtype=torch.uint8, device=dev)
    planes[n_groups.value * 8:].zero_()
    ind_entry = torch.empty(max(n_ind.value, 1), dtype=torch.int32, device=dev)
    _lib.check(lib.cto_tokenize_write(_p(text_dev), n_bytes, _p(row_off), n_rows, _p(ref_dev), int(ref_start), int(ref_dev.numel()),
                                      int(low_bq_cut), int(max_indel_length), _p(grp_off), _p(ind_off), _p(planes), _p(ind_entry), stream),
               "cto_tokenize_write")
    if cand_pos_dev is not None:
        n_cand = int(cand_pos_dev.numel())
        win_pos = torch.empty(n_cand * 33, dtype=torch.int32, device=dev)
        _lib.check(lib.cto_window_table(_p(row_pos), n_rows, _p(cand_pos_dev), n_cand, _p(win_pos), stream), "cto_window_table")
    else:
        win_pos = torch.empty(0, dtype=torch.int32, device=dev)
    ps = PackedStream(planes, grp_off, ref_code[:n_rows], ind_off, ind_entry[:n_ind.value], win_pos, n_groups.value, int(low_bq_cut), 0)
    return ps, row_pos[:n_rows]
