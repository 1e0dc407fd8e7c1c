"""Layouts of per-site read arrays (the encoder's input), shared by the host tokenizer, the synthetic
generator and the CUDA encoder.  Mirrors include/clairs_to_b200.h.

Two layouts.  ``PileupStream`` (below) is what the tokenizer and the generator produce: three byte arrays per
read.  ``PackedStream`` is what travels to the GPU: ``pack_stream`` folds the three bytes of a read into ONE byte
holding exactly what decode_pileup_bases() looks at, and stores every group of eight reads bit-sliced:

``planes[8*G]``   uint8   group g = bytes [8g, 8g+8); byte j = bit j of the group's eight packed read bytes (bit i =
                          read i).  Packed read byte: bits 0-3 symbol (``PACKED_SYMBOLS``), bit 4 plain (no indel
                          suffix), bit 5 MQ >= 20, bit 6 MQ < 20, bit 7 BQ < low-BQ cut; 0 = null read (padding).
``grp_off[P+1]``  int32   CSR offsets of rows into groups: ceil(depth / 8) groups per row.
``ref_code, ind_off, ind_entry, win_pos``: as in ``PileupStream``.

One *stream* (AFF: ``--min-BQ <platform>``, NEG: ``--min-BQ 0``; run_clairs_to:1230-1271)
is a struct-of-arrays over R reads grouped into P pileup rows (genomic positions):

``code[R]``      uint8   low nibble = symbol index in ``SYMBOLS`` (the read-opening tokens of
                         src/create_tensor_pileup_calling.py:140); bit 4 (``HAS_INDEL``) = the
                         entry carries a ``+N<seq>`` / ``-N<seq>`` suffix and therefore counts only
                         toward I/i/D/d (ibid. 160-204).
``bq[R], mq[R]`` uint8   phred values (``ord(c) - 33``, ibid. 489-490); ``QUAL_ABSENT`` marks a
                         read the reference's positional ``zip`` would have truncated away.
``pos_off[P+1]`` int32   CSR offsets of rows into the read arrays.
``ref_code[P]``  uint8   0..3 = A,C,G,T after ``evc_base_from`` coercion (ibid. 82-92, 485).
``ind_off[P+1]`` int32   CSR offsets into ``ind_entry``.
``ind_entry[K]`` uint32  one per indel-carrying read, in read order: bits 0-15 allele id (distinct
                         per (symbol, sign, sequence) key inside the row), 16-23 mq, bit 24 deletion,
                         bit 25 reverse strand (symbol not in ``ACGTN*``, ibid. 182, 199), bit 26
                         longer than ``max_indel_length`` (ibid. 174-176, 189-191).
``win_pos[N*33]`` int32  row index feeding each (candidate, flank slot), -1 = no pileup row
                         (all-zero row, ibid. 461).
"""

import numpy as np

SYMBOLS = "ACGTNacgtn*#"
SYM_INDEX = {c: i for i, c in enumerate(SYMBOLS)}
HAS_INDEL = 0x10
QUAL_ABSENT = 254

IND_DEL = 1 << 24
IND_REV = 1 << 25
IND_LONG = 1 << 26

N_POS = 33          # shared/param.py:60
N_CH = 34           # shared/param.py:56
CENTER = 16         # shared/param.py:59
MIN_RESCALE_COV = 50  # shared/param.py:26


class PileupStream:
    """Plain container of the arrays above (numpy on host, torch tensors on device)."""

    __slots__ = ("code", "bq", "mq", "pos_off", "ref_code", "ind_off", "ind_entry", "win_pos")

    def __init__(self, code, bq, mq, pos_off, ref_code, ind_off, ind_entry, win_pos):
        self.code, self.bq, self.mq = code, bq, mq
        self.pos_off, self.ref_code = pos_off, ref_code
        self.ind_off, self.ind_entry, self.win_pos = ind_off, ind_entry, win_pos

    @property
    def n_candidates(self):
        return len(self.win_pos) // N_POS

    @property
    def n_rows(self):
        return len(self.ref_code)

    @property
    def n_reads(self):
        return len(self.code)

    def arrays(self):
        return [getattr(self, k) for k in self.__slots__]

    def nbytes(self):
        return int(sum(np.asarray(a).nbytes if isinstance(a, np.ndarray) else a.numel() * a.element_size()
                       for a in self.arrays()))

    def algorithmic_bytes(self):
        """3 B per read + 5 B per window slot + 2244 B out per candidate."""
        n = self.n_candidates
        reads = int(self.pos_off[-1])
        return 3 * reads + 5 * N_POS * n + 2 * N_POS * N_CH * n


PACKED_SYMBOLS = "ACGTacgt*#Nn"


class PackedStream:
    """The encoder's input: bit-plane packed reads (1 byte per read) + the per-row arrays of ``PileupStream``.
    ``planes`` is allocated up to a multiple of 16 bytes (the CUDA encoder's bulk copies read 16-byte blocks)."""

    __slots__ = ("planes", "grp_off", "ref_code", "ind_off", "ind_entry", "win_pos", "n_groups", "low_bq_cut", "n_reads")

    def __init__(self, planes, grp_off, ref_code, ind_off, ind_entry, win_pos, n_groups, low_bq_cut, n_reads):
        self.planes, self.grp_off, self.ref_code = planes, grp_off, ref_code
        self.ind_off, self.ind_entry, self.win_pos = ind_off, ind_entry, win_pos
        self.n_groups, self.low_bq_cut, self.n_reads = int(n_groups), int(low_bq_cut), int(n_reads)

    ARRAYS = ("planes", "grp_off", "ref_code", "ind_off", "ind_entry", "win_pos")

    @property
    def n_candidates(self):
        return len(self.win_pos) // N_POS

    @property
    def n_rows(self):
        return len(self.ref_code)

    def arrays(self):
        return [getattr(self, k) for k in self.ARRAYS]

    def nbytes(self):
        return int(sum(np.asarray(a).nbytes if isinstance(a, np.ndarray) else a.numel() * a.element_size()
                       for a in self.arrays()))

    def algorithmic_bytes(self):
        """3 B per read + 5 B per window slot + 2244 B out per candidate (the survey's
        figure for byte-per-field read arrays; the packed layout moves about half of it)."""
        n = self.n_candidates
        return 3 * self.n_reads + 5 * N_POS * n + 2 * N_POS * N_CH * n


def group_offsets(pos_off):
    """CSR offsets of rows into groups of eight reads."""
    depth = np.diff(np.asarray(pos_off, dtype=np.int64))
    out = np.zeros(len(pos_off), dtype=np.int64)
    np.cumsum((depth + 7) // 8, out=out[1:])
    if out[-1] >= 2 ** 31:
        raise ValueError("one batch holds at most 2^31 read groups")
    return out.astype(np.int32)


def pack_stream(stream: PileupStream, low_bq_cut: int, n_threads: int = 0, out_planes=None) -> PackedStream:
    """PileupStream -> PackedStream through the native packer ``cto_pack_reads`` (multi-threaded).
    ``out_planes``: optional preallocated uint8 buffer (e.g. pinned) of at least the padded size."""
    import ctypes as C
    from . import _lib
    lib = _lib.lib()
    pos_off = np.ascontiguousarray(stream.pos_off, dtype=np.int32)
    grp_off = group_offsets(pos_off)
    n_groups = int(grp_off[-1])
    padded = (8 * n_groups + 15) // 16 * 16
    planes = np.empty(max(padded, 16), dtype=np.uint8) if out_planes is None else out_planes
    assert planes.nbytes >= padded and planes.dtype == np.uint8
    code, bq, mq = (np.ascontiguousarray(a, dtype=np.uint8) for a in (stream.code, stream.bq, stream.mq))
    p = lambda a: C.c_void_p(a.ctypes.data)
    _lib.check(lib.cto_pack_reads(p(code), p(bq), p(mq), p(pos_off), len(pos_off) - 1, int(low_bq_cut), p(grp_off), p(planes),
                                  int(n_threads)), "cto_pack_reads")
    return PackedStream(planes, grp_off, np.ascontiguousarray(stream.ref_code, dtype=np.uint8),
                        np.ascontiguousarray(stream.ind_off, dtype=np.int32),
                        np.ascontiguousarray(stream.ind_entry, dtype=np.uint32),
                        np.ascontiguousarray(stream.win_pos, dtype=np.int32), n_groups, low_bq_cut, len(code))


def pack_stream_numpy(stream: PileupStream, low_bq_cut: int) -> PackedStream:
    """Pure-numpy statement of the packed layout (tests check the native packer against it)."""
    old_to_new = np.array([0, 1, 2, 3, 10, 4, 5, 6, 7, 11, 8, 9, 15, 15, 15, 15], dtype=np.uint8)
    code, bq, mq = (np.asarray(a, dtype=np.uint8) for a in (stream.code, stream.bq, stream.mq))
    b = old_to_new[code & 0xF].copy()
    b |= np.where((code & HAS_INDEL) == 0, 0x10, 0).astype(np.uint8)
    b |= np.where(mq != QUAL_ABSENT, np.where(mq >= 20, 0x20, 0x40), 0).astype(np.uint8)
    b |= np.where((bq != QUAL_ABSENT) & (bq.astype(np.int32) < low_bq_cut), 0x80, 0).astype(np.uint8)
    pos_off = np.asarray(stream.pos_off, dtype=np.int64)
    grp_off = group_offsets(pos_off)
    n_groups = int(grp_off[-1])
    padded = np.zeros((n_groups, 8), dtype=np.uint8)                  # [group, read in group]
    depth = np.diff(pos_off)
    row_of = np.repeat(np.arange(len(depth)), depth)
    within = np.arange(len(code)) - pos_off[row_of]
    padded.reshape(-1)[grp_off[row_of].astype(np.int64) * 8 + within] = b
    bits = (padded[:, :, None] >> np.arange(8)[None, None, :]) & 1    # [group, read, bit]
    planes = (bits.transpose(0, 2, 1) << np.arange(8)[None, None, :]).sum(axis=2).astype(np.uint8)   # [group, plane]
    flat = np.zeros(max((8 * n_groups + 15) // 16 * 16, 16), dtype=np.uint8)
    flat[:8 * n_groups] = planes.reshape(-1)
    return PackedStream(flat, grp_off, np.asarray(stream.ref_code, dtype=np.uint8), np.asarray(stream.ind_off, dtype=np.int32),
                        np.asarray(stream.ind_entry, dtype=np.uint32), np.asarray(stream.win_pos, dtype=np.int32), n_groups,
                        low_bq_cut, len(code))
