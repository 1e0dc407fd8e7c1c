"""Device-side layout of per-site read arrays (the encoder's input), shared by the host
tokenizer, the synthetic generator and the CUDA encoder.  Mirrors include/clairs_to_b200.h.

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
