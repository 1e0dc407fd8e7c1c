"""Host-side (no GPU) entry points of the C ABI: mpileup tokenizer and chunk-file text codec.

These sit on either side of the CUDA hot path: ``tokenize_mpileup`` turns ``samtools mpileup`` text
into the encoder's struct-of-arrays input (src/create_tensor_pileup_calling.py:120-149, 472-497)
and builds the ``alt_info`` strings (ibid. 158-209); the codec reads/writes the tensor_can and
predict chunk files byte-compatibly (ibid. 551; clairs/predict.py:121-132).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .pileup_format import N_CH, N_POS, PileupStream


class Tokens:
    """stream: PileupStream (win_pos empty: the caller maps candidates to rows); row_pos: genomic position of every
    pileup row; alt_info: per row the alt_info string for candidate rows, '' otherwise (sliced lazily from one blob)."""

    def __init__(self, stream, row_pos, alt_blob, alt_off):
        self.stream, self.row_pos, self._blob, self._off, self._list = stream, row_pos, alt_blob, alt_off, None

    @property
    def alt_info(self):
        if self._list is None:
            blob, off = self._blob.decode(), self._off
            self._list = [blob[off[i]:off[i + 1]] for i in range(len(off) - 1)]
        return self._list

    def alt_info_of(self, row):
        return self._blob[self._off[row]:self._off[row + 1]].decode()


def _p(a):
    return C.c_void_p(a.ctypes.data)


def tokenize_mpileup(text, ref_seq, ref_start, candidate_pos, max_indel_length=60, n_threads=0) -> Tokens:
    lib = _lib.lib()
    if isinstance(text, str):
        text = text.encode()
    if isinstance(ref_seq, str):
        ref_seq = ref_seq.encode()
    cand = np.ascontiguousarray(np.asarray(candidate_pos if isinstance(candidate_pos, np.ndarray) else list(candidate_pos), dtype=np.int64))
    handle = C.c_void_p()
    _lib.check(lib.cto_tokenize_mpileup(text, len(text), ref_seq, len(ref_seq), int(ref_start), _p(cand), cand.size,
                                        int(max_indel_length), int(n_threads), C.byref(handle)), "cto_tokenize_mpileup")
    try:
        sizes = [C.c_int64() for _ in range(4)]
        _lib.check(lib.cto_tokens_sizes(handle, *[C.byref(s) for s in sizes]), "cto_tokens_sizes")
        n_reads, n_rows, n_ind, n_alt = [s.value for s in sizes]
        code = np.empty(n_reads, np.uint8)
        bq = np.empty(n_reads, np.uint8)
        mq = np.empty(n_reads, np.uint8)
        pos_off = np.empty(n_rows + 1, np.int32)
        ref_code = np.empty(n_rows, np.uint8)
        ind_off = np.empty(n_rows + 1, np.int32)
        ind_entry = np.empty(n_ind, np.uint32)
        row_pos = np.empty(n_rows, np.int64)
        alt = np.empty(max(n_alt, 1), np.uint8)
        alt_off = np.empty(n_rows + 1, np.int64)
        _lib.check(lib.cto_tokens_export(handle, _p(code), _p(bq), _p(mq), _p(pos_off), _p(ref_code), _p(ind_off),
                                         _p(ind_entry), _p(row_pos), _p(alt), _p(alt_off)), "cto_tokens_export")
    finally:
        lib.cto_tokens_destroy(handle)
    stream = PileupStream(code, bq, mq, pos_off, ref_code, ind_off, ind_entry, np.empty(0, np.int32))
    return Tokens(stream, row_pos, alt.tobytes()[:n_alt], alt_off)


def format_tensor_rows(tensors) -> list:
    """int16 [n,33,34] -> n strings of 1122 space-separated ints (the 4th tensor_can column)."""
    lib = _lib.lib()
    t = np.ascontiguousarray(tensors, dtype=np.int16).reshape(-1, N_POS * N_CH)
    cap = N_POS * N_CH * 7 + 8
    buf = C.create_string_buffer(cap)
    out = []
    for k in range(t.shape[0]):
        n = lib.cto_format_tensor_row(_p(t[k]), buf, cap)
        if n < 0:
            raise _lib.CtoError("cto_format_tensor_row: buffer too small")
        out.append(buf.raw[:n].decode())
    return out


def parse_tensor_row(text) -> np.ndarray:
    lib = _lib.lib()
    if isinstance(text, str):
        text = text.encode()
    row = np.empty(N_POS * N_CH, np.int16)
    _lib.check(lib.cto_parse_tensor_row(text, len(text), _p(row)), "cto_parse_tensor_row")
    return row.reshape(N_POS, N_CH)


def format_prob_fields(probs) -> str:
    """float32 [k,2] -> 'p0 p1<TAB>p0 p1...' with 8 decimals, exactly like "{:0.8f}".format."""
    lib = _lib.lib()
    p = np.ascontiguousarray(probs, dtype=np.float32).reshape(-1, 2)
    cap = 64 * p.shape[0] + 64
    buf = C.create_string_buffer(cap)
    n = lib.cto_format_prob_fields(_p(p), p.shape[0], buf, cap)
    if n < 0:
        raise _lib.CtoError("cto_format_prob_fields: buffer too small")
    return buf.raw[:n].decode()


class TensorFile:
    """A whole tensor_can chunk file parsed by one native call (``cto_parse_tensor_file``): ``tensor`` int16
    [n,33,34], ``depth`` int32 [n], and the text fields of row r as slices of ``text`` via ``field(r, k)``
    (k: 0 contig, 1 position, 2 ref33, 4 alt_info, 5 variant_type, 6 ref_centre).  Rows whose centre reference
    base is not ACGT are already dropped (clairs/predict.py:219-220)."""

    def __init__(self, text: bytes):
        lib = _lib.lib()
        self.text = text
        max_rows = text.count(b"\n") + 1
        self.tensor = np.empty((max_rows, N_POS, N_CH), np.int16)
        self.depth = np.empty(max_rows, np.int32)
        self.fields = np.empty((max_rows, 7, 2), np.int64)
        n = C.c_int64(0)
        _lib.check(lib.cto_parse_tensor_file(text, len(text), max_rows, _p(self.tensor), _p(self.depth), _p(self.fields),
                                             C.byref(n)), "cto_parse_tensor_file")
        self.n = int(n.value)
        self.tensor, self.depth, self.fields = self.tensor[:self.n], self.depth[:self.n], self.fields[:self.n]

    def field(self, r, k) -> str:
        o, ln = self.fields[r, k]
        return self.text[o:o + ln].decode()


def format_predict_rows(tf: TensorFile, start, n, fwd, rev, probs, n_heads) -> bytes:
    """The predict-file rows (clairs/predict.py:114-152) of rows [start, start+n) of a parsed tensor file."""
    lib = _lib.lib()
    fwd = np.ascontiguousarray(fwd, dtype=np.int32).reshape(n, 4)
    rev = np.ascontiguousarray(rev, dtype=np.int32).reshape(n, 4)
    probs = np.ascontiguousarray(probs, dtype=np.float32).reshape(n, 2 * n_heads, 2)
    fields = np.ascontiguousarray(tf.fields[start:start + n])
    cap = int(fields[:, (0, 1, 4), 1].sum()) + n * (2 * 64 + n_heads * 64 + 64) + 64
    buf = C.create_string_buffer(cap)
    w = lib.cto_format_predict_rows(tf.text, _p(fields), n, _p(fwd), _p(rev), _p(probs), n_heads, buf, cap)
    if w < 0:
        raise _lib.CtoError("cto_format_predict_rows failed (%d)" % w)
    return buf.raw[:w]


def format_tensor_can_rows(ctg, pos, ref33, tensors, alt_infos, types) -> bytes:
    """All rows of a tensor_can chunk file (src/create_tensor_pileup_calling.py:561-568) with one native call.
    pos: n genomic positions; ref33: n strings of 33 reference bases; tensors int16 [n,33,34]; alt_infos / types: n strings."""
    lib = _lib.lib()
    n = len(pos)
    if n == 0:
        return b""
    pos = np.ascontiguousarray(pos, dtype=np.int64)
    t = np.ascontiguousarray(tensors, dtype=np.int16).reshape(n, N_POS * N_CH)
    ref = "".join(ref33).encode()
    assert len(ref) == n * N_POS, "every reference context must be %d bases" % N_POS
    parts, alt_off, type_off, o = [], np.empty((n, 2), np.int64), np.empty((n, 2), np.int64), 0
    for k in range(n):
        a, ty = alt_infos[k].encode(), types[k].encode()
        alt_off[k] = (o, len(a)); o += len(a)
        type_off[k] = (o, len(ty)); o += len(ty)
        parts.append(a); parts.append(ty)
    blob = b"".join(parts)
    ctg_b = ctg.encode()
    cap = n * (len(ctg_b) + 24 + N_POS + N_POS * N_CH * 7 + 16) + len(blob) + 64
    buf = C.create_string_buffer(cap)
    w = lib.cto_format_tensor_can_rows(ctg_b, len(ctg_b), n, _p(pos), ref, _p(t), blob, _p(alt_off), _p(type_off), buf, cap)
    if w < 0:
        raise _lib.CtoError("cto_format_tensor_can_rows failed (%d)" % w)
    return buf.raw[:w]


class PredictFile:
    """A whole predict chunk file parsed by one native call (``cto_parse_predict_file``): ``p_aff`` / ``p_neg`` double
    [n, H] (P(positive class) per head, the doubles the reference gets from float()), and the text fields of row r as
    slices of ``text`` via ``field(r, k)`` (k: 0 chrom, 1 pos, 2 ref, 3 alt_info, 4 forward counts, 5 reverse counts)."""

    def __init__(self, text: bytes, n_heads: int):
        lib = _lib.lib()
        self.text = text
        max_rows = text.count(b"\n") + 1
        self.p_aff = np.empty((max_rows, n_heads), np.float64)
        self.p_neg = np.empty((max_rows, n_heads), np.float64)
        self.fields = np.empty((max_rows, 6, 2), np.int64)
        n = C.c_int64(0)
        _lib.check(lib.cto_parse_predict_file(text, len(text), int(n_heads), max_rows, _p(self.p_aff), _p(self.p_neg), _p(self.fields),
                                              C.byref(n)), "cto_parse_predict_file")
        self.n = int(n.value)
        self.p_aff, self.p_neg, self.fields = self.p_aff[:self.n], self.p_neg[:self.n], self.fields[:self.n]

    def field(self, r, k) -> str:
        o, ln = self.fields[r, k]
        return self.text[o:o + ln].decode()
