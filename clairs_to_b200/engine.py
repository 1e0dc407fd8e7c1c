"""Python face of the CUDA engine: torch tensors are containers for device memory, all arithmetic
happens in ``libcto_b200.so`` (include/clairs_to_b200.h).  No CPU fallback.

The method names mirror the reference call sites they replace:
  ``encode``      decode_pileup_bases + window assembly (src/create_tensor_pileup_calling.py:95-229, 537-570)
  ``forward_aff`` model_aff(x)   (clairs/predict.py:646, clairs/model.py:231)
  ``forward_neg`` model_neg(x)   (clairs/predict.py:648, clairs/model.py:440)
  ``predict``     one pass of the predict() mini-batch loop (clairs/predict.py:610-699) + the Bayes
                  combine of output_vcf_from_probability (clairs/call_variants.py:154-304)
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .pileup_format import N_CH, N_POS, PackedStream, PileupStream, pack_stream
from .weights import export_aff, export_neg, likelihood_tables, state_dict_from_checkpoint


def low_bq_cut_for(platform: str) -> int:
    """The literal of decode_pileup_bases (src/create_tensor_pileup_calling.py:149).  NOTE: create_tensor()
    never forwards its --platform to that function (ibid. 499-511), so inside the pipeline the callee's
    default 'ont' applies and the cut is always ``PIPELINE_LOW_BQ_CUT`` (pinned by tests/golden/create_tensor)."""
    return 30 if platform == 'ont' else 10


PIPELINE_LOW_BQ_CUT = low_bq_cut_for('ont')


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        assert t.is_contiguous()
        return C.c_void_p(t.data_ptr())
    if isinstance(t, np.ndarray):
        assert t.flags['C_CONTIGUOUS']
        return C.c_void_p(t.ctypes.data)
    raise TypeError(type(t))


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_TORCH_VIEW = {np.dtype(np.uint32): np.int32}


def packed_to_device(ps: PackedStream, device) -> PackedStream:
    """Host PackedStream -> the same arrays in HBM (uint32 carried as int32 bits)."""
    out = []
    for a in ps.arrays():
        if isinstance(a, torch.Tensor):
            out.append(a.to(device, non_blocking=True))
            continue
        a = np.ascontiguousarray(a)
        if a.dtype in _TORCH_VIEW:
            a = a.view(_TORCH_VIEW[a.dtype])
        out.append(torch.from_numpy(a).to(device, non_blocking=True))
    return PackedStream(*out, ps.n_groups, ps.low_bq_cut, ps.n_reads)


class DeviceStream:
    """A PileupStream on its way to the GPU.  The low-BQ literal of decode_pileup_bases
    (src/create_tensor_pileup_calling.py:149) is applied when the reads are packed to one byte, so the packed device
    copy is made (and cached) per cut on first use."""

    def __init__(self, stream: PileupStream, device):
        self.host, self.device, self._packed = stream, device, {}

    def packed(self, low_bq_cut: int) -> PackedStream:
        cut = int(low_bq_cut)
        if cut not in self._packed:
            self._packed[cut] = packed_to_device(pack_stream(self.host, cut), self.device)
        return self._packed[cut]

    @property
    def win_pos(self):
        return self.host.win_pos


def stream_to_device(stream: PileupStream, device) -> DeviceStream:
    """numpy PileupStream -> handle whose packed device copy is built at the first encode."""
    return DeviceStream(stream, device)


def encode_pileup(s, low_bq_cut: int = None, device=None):
    """DeviceStream / device PackedStream -> (int16 [N,33,34], int32 centre depth [N]) through ``cto_encode_pileup``.
    Needs no weights, so the create_tensor sub-command uses it without building an Engine."""
    lib = _lib.lib()
    if isinstance(s, DeviceStream):
        s = s.packed(low_bq_cut)
    elif low_bq_cut is not None and int(low_bq_cut) != s.low_bq_cut:
        raise ValueError("stream was packed with low_bq_cut=%d, encode asked for %d" % (s.low_bq_cut, low_bq_cut))
    device = s.win_pos.device if device is None else device
    n = s.win_pos.numel() // N_POS
    tensor = torch.empty((n, N_POS, N_CH), dtype=torch.int16, device=device)
    depth = torch.empty((n,), dtype=torch.int32, device=device)
    _lib.check(lib.cto_encode_pileup(_ptr(s.planes), _ptr(s.grp_off), _ptr(s.ref_code), _ptr(s.ind_off), _ptr(s.ind_entry),
                                     _ptr(s.win_pos), n, s.n_groups, _ptr(tensor), _ptr(depth), _stream_ptr()),
               "cto_encode_pileup")
    return tensor, depth


class Engine:
    def __init__(self, aff_state_dict, neg_state_dict, max_batch=8192, device=None, likelihood=None):
        if not torch.cuda.is_available():
            raise _lib.CtoError("clairs_to_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _lib.lib()
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        aff_blob, aff_cfg = export_aff(aff_state_dict)
        neg_blob, neg_cfg = export_neg(neg_state_dict)
        handle = C.c_void_p()
        _lib.check(self.lib.cto_engine_create(_ptr(aff_blob), aff_blob.size, _ptr(aff_cfg), aff_cfg.size,
                                              _ptr(neg_blob), neg_blob.size, _ptr(neg_cfg), neg_cfg.size,
                                              int(max_batch), C.byref(handle)), "cto_engine_create")
        self.handle = handle
        self.n_heads = int(self.lib.cto_engine_heads(handle))
        self.max_batch = int(max_batch)
        self.has_likelihood = False
        if likelihood is not None:
            self.set_likelihood(likelihood)

    @classmethod
    def from_checkpoints(cls, chkpnt_fn_acgt, chkpnt_fn_nacgt, **kw):
        return cls(state_dict_from_checkpoint(chkpnt_fn_acgt, 'model_acgt'),
                   state_dict_from_checkpoint(chkpnt_fn_nacgt, 'model_nacgt'), **kw)

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.cto_engine_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tensor_cores(self, mode):
        """0 / False: the exact fp32 CUDA-core engine; 1 / True (default): bf16x3 tcgen05 contractions with the fused
        AFF transformer kernel; 2: tcgen05 contractions with one kernel per op (round-1 layers, kept for A/B tests)."""
        _lib.check(self.lib.cto_engine_set_tensor_cores(self.handle, int(mode)), "set_tensor_cores")

    def set_overlap(self, enable=True):
        """AFF on a second stream beside NEG inside predict / run_sites (default on; identical results)."""
        _lib.check(self.lib.cto_engine_set_overlap(self.handle, int(bool(enable))), "set_overlap")

    def neg_recurrence(self, xproj, n, two_chains=True):
        """Layer-2 GRU recurrence alone (kernel-level hook, cto_neg_recurrence): xproj fp32 [6H, 33 * bp] on the device ->
        (out_hi, out_mid) int16 views of the bf16 planes [n, 33, 2H]."""
        bp = (n + 127) // 128 * 128
        assert xproj.dtype == torch.float32 and xproj.is_contiguous() and xproj.shape[1] == N_POS * bp, xproj.shape
        h2 = xproj.shape[0] // 6
        hi = torch.zeros((n, N_POS, 2 * h2), dtype=torch.int16, device=self.device)
        mid = torch.zeros_like(hi)
        _lib.check(self.lib.cto_neg_recurrence(self.handle, _ptr(xproj), n, _ptr(hi), _ptr(mid), int(bool(two_chains)), _stream_ptr()),
                   "neg_recurrence")
        return hi, mid

    def workspace(self, which):
        """Test hook (cto_engine_workspace): a COPY of one NEG workspace tensor as raw bytes (uint8 device tensor)."""
        import ctypes as C
        nbytes = C.c_int64()
        _lib.check(self.lib.cto_engine_workspace(self.handle, int(which), None, C.byref(nbytes)), "workspace")
        out = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.cto_engine_workspace(self.handle, int(which), _ptr(out), C.byref(nbytes)), "workspace")
        return out

    def fused_status(self):
        """Watchdog record of the fused AFF kernel (synchronises): all zeros unless an in-kernel barrier wait timed out."""
        out = np.zeros(8, np.int32)
        _lib.check(self.lib.cto_engine_fused_status(self.handle, _ptr(out)), "fused_status")
        return out

    def set_qual_thresholds(self, qual_pass, qual_phaseable=None, qual_unphaseable=None):
        """QUAL -> FILTER thresholds evaluated on the device (clairs/call_variants.py:67-76 ``--qual``;
        src/postprocess_vcf.py:61-82; platform defaults in shared/param.py:35-40)."""
        qp = qual_pass if qual_phaseable is None else qual_phaseable
        qu = qual_pass if qual_unphaseable is None else qual_unphaseable
        _lib.check(self.lib.cto_engine_set_qual_thresholds(self.handle, float(qual_pass), float(qp), float(qu)), "set_qual_thresholds")

    def set_likelihood(self, path_or_array):
        tables = np.ascontiguousarray(likelihood_tables(path_or_array, self.n_heads))
        _lib.check(self.lib.cto_engine_set_likelihood(self.handle, _ptr(tables), self.n_heads), "set_likelihood")
        self.has_likelihood = True

    # ---- encoder ----------------------------------------------------------------------------
    def encode(self, s, low_bq_cut: int = None):
        """DeviceStream / device PackedStream -> (int16 [N,33,34], int32 depth [N])."""
        return encode_pileup(s, low_bq_cut, self.device)

    # ---- networks ---------------------------------------------------------------------------
    def rescale(self, x_i16, depth):
        out = torch.empty(x_i16.shape, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.cto_rescale(_ptr(x_i16), _ptr(depth), x_i16.shape[0], _ptr(out), _stream_ptr()), "cto_rescale")
        return out

    def _forward(self, fn, x):
        x = x.to(self.device, torch.float32).contiguous()
        assert x.shape[1:] == (N_POS, N_CH), x.shape
        logits = torch.empty((x.shape[0], self.n_heads, 2), dtype=torch.float32, device=self.device)
        _lib.check(fn(self.handle, _ptr(x), x.shape[0], _ptr(logits), _stream_ptr()), "forward")
        return logits

    def forward_aff(self, x):
        return self._forward(self.lib.cto_forward_aff, x)

    def forward_neg(self, x):
        return self._forward(self.lib.cto_forward_neg, x)

    def predict(self, x_aff, depth_aff, x_neg=None, depth_neg=None, posterior=None):
        """int16 tensors [N,33,34] (+ centre depths) -> dict of device tensors."""
        if x_neg is None:
            x_neg, depth_neg = x_aff, depth_aff
        n = x_aff.shape[0]
        posterior = self.has_likelihood if posterior is None else posterior
        dev, h = self.device, self.n_heads
        out = dict(
            logits_aff=torch.empty((n, h, 2), dtype=torch.float32, device=dev),
            logits_neg=torch.empty((n, h, 2), dtype=torch.float32, device=dev),
            probs=torch.empty((n, 2 * h, 2), dtype=torch.float32, device=dev),
            fwd=torch.empty((n, 4), dtype=torch.int32, device=dev),
            rev=torch.empty((n, 4), dtype=torch.int32, device=dev),
            post=torch.empty((n, h), dtype=torch.float64, device=dev) if posterior else None,
            call=torch.empty((n,), dtype=torch.int32, device=dev) if posterior else None,
            qual=torch.empty((n,), dtype=torch.float64, device=dev) if posterior else None,
            filter=torch.empty((n,), dtype=torch.int32, device=dev) if posterior else None,
        )
        _lib.check(self.lib.cto_predict(self.handle, _ptr(x_aff), _ptr(depth_aff), _ptr(x_neg), _ptr(depth_neg), n,
                                        _ptr(out['logits_aff']), _ptr(out['logits_neg']), _ptr(out['probs']),
                                        _ptr(out['post']), _ptr(out['call']), _ptr(out['fwd']), _ptr(out['rev']),
                                        _ptr(out['qual']), _ptr(out['filter']), _stream_ptr()), "cto_predict")
        return out

    def run_sites(self, aff, neg, low_bq_cut: int = None, posterior=None):
        """Device-resident hot path: encode both streams, then predict."""
        xa, da = self.encode(aff, low_bq_cut)
        if neg is None:
            xn, dn = xa, da
        else:
            xn, dn = self.encode(neg, low_bq_cut)
        out = self.predict(xa, da, xn, dn, posterior=posterior)
        out.update(tensor_aff=xa, tensor_neg=xn, depth_aff=da, depth_neg=dn)
        return out

    # ---- end to end from host memory ----------------------------------------------------------
    @staticmethod
    def _host_struct(s: PackedStream):
        hs = _lib.HostStream()
        keep = []
        for name in PackedStream.ARRAYS:
            a = getattr(s, name)
            if isinstance(a, torch.Tensor):
                ptr = a.data_ptr()
            else:
                a = np.ascontiguousarray(a)
                ptr = a.ctypes.data
            keep.append(a)
            setattr(hs, name, ptr)
        hs.n_groups = s.n_groups
        hs.n_rows = len(s.ref_code)
        hs.n_ind = len(s.ind_entry)
        return hs, keep

    def run_sites_text(self, text_aff, text_neg, ref, ref_start, cand_pos, low_bq_cut, out=None, pieces=None):
        """mpileup text (pinned host tensors) -> host results; tokenized on the device (see ``run_sites_text`` below)."""
        return run_sites_text(self, text_aff, text_neg, ref, ref_start, cand_pos, low_bq_cut, out, pieces)

    def run_sites_host(self, aff, neg, low_bq_cut: int = None, out=None, want_tensors=False):
        """Host arrays in -> host results out through ``cto_run_sites_host`` (H2D + D2H inside).  ``aff`` / ``neg``:
        PackedStream (host, ideally pinned), or PileupStream, which is packed here first (``low_bq_cut`` required).
        ``out`` may carry preallocated (pinned) host tensors 'probs', 'post', 'call'."""
        if isinstance(aff, PileupStream):
            aff = pack_stream(aff, low_bq_cut)
        if isinstance(neg, PileupStream):
            neg = pack_stream(neg, low_bq_cut)
        n = len(aff.win_pos) // N_POS
        h = self.n_heads
        out = dict(out or {})
        if 'probs' not in out:
            out['probs'] = torch.empty((n, 2 * h, 2), dtype=torch.float32)
        if self.has_likelihood:
            out.setdefault('post', torch.empty((n, h), dtype=torch.float64))
            out.setdefault('call', torch.empty((n,), dtype=torch.int32))
            out.setdefault('qual', torch.empty((n,), dtype=torch.float64))
            out.setdefault('filter', torch.empty((n,), dtype=torch.int32))
        if want_tensors:
            out['tensor_aff'] = torch.empty((n, N_POS, N_CH), dtype=torch.int16)
            out['tensor_neg'] = torch.empty((n, N_POS, N_CH), dtype=torch.int16)
        ha, keep_a = self._host_struct(aff)
        hn, keep_n = (None, None) if neg is None else self._host_struct(neg)
        _lib.check(self.lib.cto_run_sites_host(self.handle, C.byref(ha), C.byref(hn) if hn is not None else None, n,
                                               _ptr(out['probs']), _ptr(out.get('post')),
                                               _ptr(out.get('call')), _ptr(out.get('qual')), _ptr(out.get('filter')),
                                               _ptr(out.get('tensor_aff')),
                                               _ptr(out.get('tensor_neg')), _stream_ptr()), "cto_run_sites_host")
        return out


def _row_start_at_or_after(text: torch.Tensor, n: int, pos: int) -> int:
    """Byte offset of the first mpileup row whose position (column 2) is >= ``pos``; rows are position sorted.  Bisection on
    byte offsets: jump to the next row start behind the probe and read its position."""
    view = text.numpy()

    def row_after(b):                                           # start of the first row that begins at or behind byte b
        if b <= 0:
            return 0
        k, step = b - 1, 512
        while k < n:
            hit = np.flatnonzero(view[k:k + step] == 10)
            if len(hit):
                return min(k + int(hit[0]) + 1, n)
            k += step
            step = min(step * 4, 1 << 20)
        return n

    def position(start):
        cols = view[start:start + 256].tobytes().split(b"\t", 2)
        digits = cols[1] if len(cols) > 1 else b""
        k = 0
        while k < len(digits) and 48 <= digits[k] <= 57:
            k += 1
        return int(digits[:k] or b"0")

    lo, hi = 0, n                                               # invariant: rows starting before lo are < pos; rows from hi on are >= pos
    while lo < hi:
        mid = (lo + hi) // 2
        start = row_after(mid)
        if start >= n or position(start) >= pos:
            hi = mid
        else:
            lo = start + 1
    return row_after(lo)


def run_sites_text(eng: "Engine", text_aff, text_neg, ref: bytes, ref_start: int, cand_pos, low_bq_cut: int, out=None, pieces: int = None,
                   max_indel_length: int = 60):
    """mpileup TEXT in host memory (pinned uint8 tensors, one per stream, rows position sorted) -> probabilities (and posterior,
    QUAL, FILTER when the engine has likelihood tables) in host memory.  The host only copies: rows are indexed and tokenized on
    the device (device_tokenizer), then encoded and run through both networks.  The candidate list is cut into ``pieces`` (default: one per engine chunk): while
    the networks of piece k run on the caller's stream, piece k + 1 is tokenized on a second stream and the text of piece k + 2
    is copied on a third.  ``text_neg`` None: one stream feeds both
    networks (Illumina, run_clairs_to:1248-1252)."""
    from .device_tokenizer import tokenize_text_device
    dev = eng.device
    cand = np.ascontiguousarray(cand_pos, dtype=np.int64)
    n = len(cand)
    if n > 1 and bool(np.any(cand[1:] < cand[:-1])):
        raise ValueError("run_sites_text: candidate positions must be in ascending order (the text is cut by position)")
    h = eng.n_heads
    out = dict(out or {})
    if 'probs' not in out:
        out['probs'] = torch.empty((n, 2 * h, 2), dtype=torch.float32).pin_memory()
    if eng.has_likelihood:
        for key, shape, dt in (('post', (n, h), torch.float64), ('call', (n,), torch.int32), ('qual', (n,), torch.float64), ('filter', (n,), torch.int32)):
            if key not in out:
                out[key] = torch.empty(shape, dtype=dt).pin_memory()
    if n == 0:
        return out
    if not hasattr(eng, "_text_copy_stream"):
        eng._text_copy_stream = torch.cuda.Stream(device=dev)
    cs, main = eng._text_copy_stream, torch.cuda.current_stream()
    ref_dev = torch.frombuffer(bytearray(ref), dtype=torch.uint8).to(dev, non_blocking=True)
    texts = [t for t in (text_aff, text_neg) if t is not None]
    if pieces is None:                                          # one piece per engine chunk: every network pass runs full waves
        step = int(eng.max_batch)
        bounds = list(range(0, n, step)) + [n]
        pieces = len(bounds) - 1
    else:
        pieces = max(1, min(int(pieces), n))
        bounds = [n * k // pieces for k in range(pieces + 1)]
    spans = []                                                  # per piece and stream: byte range of the rows its windows touch
    for k in range(pieces):
        lo_pos, hi_pos = int(cand[bounds[k]]) - 16, int(cand[bounds[k + 1] - 1]) + 16
        spans.append([(_row_start_at_or_after(t, t.numel(), lo_pos), _row_start_at_or_after(t, t.numel(), hi_pos + 1)) for t in texts])

    # two text slots per stream, kept on the engine between calls (a fresh allocation per piece made the caching allocator
    # fall back to cudaMalloc / cudaFree with their implicit synchronisation: 320 ms instead of 55 ms per 100 k sites)
    need = max(b - a for sp in spans for a, b in sp) + 32
    slots = getattr(eng, "_text_slots", None)
    if slots is None or len(slots) != len(texts) or slots[0][0].numel() < need:
        slots = eng._text_slots = [[torch.empty(need + need // 8, dtype=torch.uint8, device=dev) for _ in range(2)] for _ in texts]
        main.synchronize()
    consumed = [None, None]                                    # per slot: event recorded once its text has been tokenized

    def issue_copy(k):
        slot = k & 1
        bufs = []
        with torch.cuda.stream(cs):
            if consumed[slot] is not None:
                cs.wait_event(consumed[slot])
            for t, pair, (a, b) in zip(texts, slots, spans[k]):
                buf = pair[slot]
                buf[b - a:b - a + 32].zero_()
                if b > a:
                    buf[:b - a].copy_(t[a:b], non_blocking=True)
                bufs.append((buf, b - a))
            ev = torch.cuda.Event()
            ev.record(cs)
        return bufs, ev

    if not hasattr(eng, "_text_tok_stream"):
        eng._text_tok_stream = torch.cuda.Stream(device=dev)
    ts = eng._text_tok_stream
    ready = torch.cuda.Event()
    ready.record(main)                                          # ref_dev and the slots are ready for the side streams after this
    cs.wait_event(ready)
    ts.wait_event(ready)

    def tokenize(k, copied):
        """Piece k on the tokenizer stream: the host waits for ITS row / group counts only, the networks of the previous piece
        keep running on the caller's stream meanwhile."""
        bufs, ev = copied
        with torch.cuda.stream(ts):
            ts.wait_event(ev)
            cand_dev = torch.from_numpy(cand[bounds[k]:bounds[k + 1]]).to(dev, non_blocking=True)
            packed = [tokenize_text_device(buf, nb, ref_dev, ref_start, low_bq_cut, cand_dev, max_indel_length)[0] for buf, nb in bufs]
            done = torch.cuda.Event()
            done.record(ts)
        consumed[k & 1] = done
        for ps in packed:
            for a in ps.arrays():
                a.record_stream(main)
        return packed, done

    copies = {0: issue_copy(0)}
    if pieces > 1:
        copies[1] = issue_copy(1)
    tok = tokenize(0, copies.pop(0))
    for k in range(pieces):
        packed, done = tok
        main.wait_event(done)
        c0, c1 = bounds[k], bounds[k + 1]
        res = eng.run_sites(packed[0], packed[1] if len(packed) > 1 else None, low_bq_cut)
        if k + 2 < pieces:                                      # slot k & 1 is free again: piece k has been tokenized
            copies[k + 2] = issue_copy(k + 2)
        if k + 1 < pieces:                                      # tokenize the next piece beside the networks of this one
            tok = tokenize(k + 1, copies.pop(k + 1))
        out['probs'][c0:c1].copy_(res['probs'], non_blocking=True)
        for key in ('post', 'call', 'qual', 'filter'):
            if key in out and res.get(key) is not None:
                out[key][c0:c1].copy_(res[key], non_blocking=True)
    main.synchronize()
    return out


def gemm_nt(a, w, bias=None, residual=None, act=0, tensor_cores=True):
    """C = act(A @ W^T + bias) (+ residual) through ``cto_gemm_nt``: the dense-contraction building
    block of both networks, on tcgen05 (bf16x3) or on the fp32 CUDA-core kernel.  ``tensor_cores`` may
    also be the integer mode mask of ``cto_gemm_nt`` (1 | 2 pre-split A | 4 split C | 8 bias per row | 16 wide tiles)."""
    lib = _lib.lib()
    assert a.is_cuda and a.dtype == torch.float32 and w.dtype == torch.float32
    m, k = a.shape
    n = w.shape[0]
    c = torch.empty((m, n), dtype=torch.float32, device=a.device)
    assert a.stride(1) == 1 and w.is_contiguous()
    _lib.check(lib.cto_gemm_nt(C.c_void_p(a.data_ptr()), a.stride(0), _ptr(w), _ptr(bias), _ptr(residual),
                               residual.stride(0) if residual is not None else 0, _ptr(c), n, m, n, k, int(act),
                               int(tensor_cores), _stream_ptr()), "cto_gemm_nt")
    return c
