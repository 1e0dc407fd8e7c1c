/*
 * clairs_to_b200.h -- C ABI of the B200-native drop-in for the ClairS-TO per-candidate hot path
 * (pileup tensor encoder -> AFF/NEG forward -> posterior combine).
 *
 * The reference (HKU-BAL/ClairS-TO v0.4.4) has NO FFI for this path: it is Python calling
 * PyTorch/NumPy.  Each entry point below therefore cites the reference
 * Python function it replaces; INTEGRATION.md shows the ctypes binding a maintainer would add
 * inside src/create_tensor_pileup_calling.py, clairs/predict.py and clairs/call_variants.py.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; "dev" pointers are CUDA device memory on the
 *     current device, "host" pointers are ordinary (preferably pinned) host memory;
 *   - the caller owns every input/output buffer; the library owns weights + workspace only;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are
 *     stream-ordered and return without synchronising unless stated;
 *   - return 0 on success, non-zero on error with a message in cto_last_error() (thread-local).
 *     The Python shims turn that into sys.exit(msg) like the reference's own error style
 *     (e.g. src/create_tensor_pileup_calling.py:424).  There is no CPU fallback.
 */
#ifndef CLAIRS_TO_B200_H
#define CLAIRS_TO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTO_N_POS 33       /* shared/param.py:60 */
#define CTO_N_CH 34        /* shared/param.py:56 */
#define CTO_ABI_VERSION 4

typedef struct cto_engine cto_engine;

int cto_abi_version(void);
const char* cto_last_error(void);

/* Checks that the current CUDA device is sm_100 (B200).  Returns 0 and writes the SM count. */
int cto_device_check(int* sm_count);

/*
 * Pileup tensor encoder.  Replaces decode_pileup_bases() + the window assembly of create_tensor()
 * (src/create_tensor_pileup_calling.py:146-229, 461, 513-516, 537-543).  Input = one packed stream (see
 * clairs_to_b200/pileup_format.py and cto_pack_reads below): `planes` holds 8 bit-plane bytes per group of eight
 * reads, grp_off[row] .. grp_off[row + 1] are the groups of a pileup row; ref_code / ind_off / ind_entry / win_pos as
 * documented there.  The low-BQ literal of ibid. 149 (30 if platform == 'ont' else 10) is applied when the reads are
 * packed.  Output tensor int16 [n_candidates, 33, 34]; depth_dev (nullable) receives the centre-row depth that leads
 * the reference's alt_info string (ibid. 208).
 * Contract: planes_dev 8-byte aligned (16-byte aligned arrays take the faster bulk-copy path) and ALLOCATED up to the
 * next multiple of 16 bytes beyond 8 * n_groups.
 */
int cto_encode_pileup(const uint8_t* planes_dev, const int32_t* grp_off_dev, const uint8_t* ref_code_dev,
                      const int32_t* ind_off_dev, const uint32_t* ind_entry_dev, const int32_t* win_pos_dev,
                      int64_t n_candidates, int64_t n_groups, int16_t* tensor_dev, int32_t* depth_dev, void* stream);

/*
 * Host packer: per-read byte arrays (code / bq / mq as produced by cto_tokenize_mpileup; pos_off = CSR offsets of the
 * rows) -> the encoder's bit-plane layout.  One packed byte per read: bits 0-3 symbol (0-3 ACGT, 4-7 acgt, 8 '*',
 * 9 '#', 10 N, 11 n), bit 4 "plain" (no indel suffix: src/create_tensor_pileup_calling.py:160-204 counts indel reads
 * only toward I/D), bit 5 MQ >= 20 (ibid. 147), bit 6 MQ < 20 (148), bit 7 BQ < low_bq_cut (149); rows are padded to
 * whole groups of eight reads with null bytes, and each group is stored bit-sliced (byte j = bit j of its eight reads).
 * grp_off: int32 [n_rows + 1], the prefix sum of ceil(row depth / 8) (computed by the caller); planes_out: 8 *
 * grp_off[n_rows] bytes (+ padding to a multiple of 16, which is zero-filled).  n_threads <= 0: all host cores.
 */
int cto_pack_reads(const uint8_t* code, const uint8_t* bq, const uint8_t* mq, const int32_t* pos_off, int64_t n_rows,
                   int low_bq_cut, const int32_t* grp_off, uint8_t* planes_out, int n_threads);

/*
 * Engine = AFF + NEG weights (flat fp32 blobs in HOST memory, produced by
 * clairs_to_b200/weights.py from the reference checkpoints that clairs/predict.py:512-568 loads)
 * plus workspace for `max_batch` candidates per internal chunk.
 *   aff_cfg = [n_heads, n_stages, (C, heads, depth) per stage]   neg_cfg = [n_heads, 34, H1, H2]
 */
int cto_engine_create(const float* aff_blob_host, int64_t aff_len, const int32_t* aff_cfg, int aff_cfg_len,
                      const float* neg_blob_host, int64_t neg_len, const int32_t* neg_cfg, int neg_cfg_len,
                      int64_t max_batch, cto_engine** out);
void cto_engine_destroy(cto_engine* e);
int cto_engine_heads(const cto_engine* e);   /* 4 (SNV) or 6 (indel) */

/*
 * Likelihood tables for the posterior (clairs/call_variants.py:655-796): per head 100 matrix
 * entries (row = AFF bin), then 11 AFF bin edges, then 11 NEG bin edges, as doubles in host memory.
 */
int cto_engine_set_likelihood(cto_engine* e, const double* tables_host, int n_heads);

/* int16 tensor -> fp32 network input with the >50x depth rescale of clairs/predict.py:179-197. */
int cto_rescale(const int16_t* x_dev, const int32_t* depth_dev, int64_t n, float* out_dev, void* stream);

/*
 * model(x) of clairs/model.py:231 (CvT / CvT_Indel) and :440 (BiGRU_NACGT / _Indel):
 * x fp32 [n, 33, 34] -> post-SELU "logits" fp32 [n, n_heads, 2].  Any n (chunked internally).
 */
int cto_forward_aff(cto_engine* e, const float* x_dev, int64_t n, float* logits_dev, void* stream);
int cto_forward_neg(cto_engine* e, const float* x_dev, int64_t n, float* logits_dev, void* stream);

/*
 * Softmax(dim=1) of clairs/predict.py:574, 659-684 and, when the engine has likelihood tables and
 * post_dev/call_dev are non-NULL, the Bayes combine + argmax of clairs/call_variants.py:154-224
 * (SNV) / 226-304 (indel) in fp64, on probabilities rounded through the reference's 8-decimal
 * text round trip.  probs_dev fp32 [n, 2*n_heads, 2] in predict-file order (a c g t [i d] na ...);
 * post_dev double [n, n_heads]; call_dev int32 [n]: bits 0-7 argmax, bit 8 = bin index clamped.
 */
int cto_softmax_posterior(cto_engine* e, const float* logits_aff_dev, const float* logits_neg_dev, int64_t n,
                          float* probs_dev, double* post_dev, int32_t* call_dev, double* qual_dev, int32_t* filter_dev,
                          void* stream);
/*
 * QUAL and the QUAL -> FILTER thresholds on the device (same kernel): qual_dev double [n] = quality_score_from of the
 * winning posterior (clairs/call_variants.py:81-88: Phred_Trans * log(((1-p)+1e-10)/(p+1e-10)) + 2, floored at 0,
 * 4 decimals); filter_dev int32 [n]: bit 0 QUAL >= qual_pass (`--qual`, call_variants.py:67-76), bit 1 QUAL >=
 * qual_phaseable, bit 2 QUAL >= qual_unphaseable (src/postprocess_vcf.py:61-82; defaults shared/param.py:35-40).
 * The RefCall / variant decision needs the alt_info strings (call_variants.py:306-358) and stays on the host.
 * Thresholds default to 0 (every QUAL passes).
 */
int cto_engine_set_qual_thresholds(cto_engine* e, double qual_pass, double qual_phaseable, double qual_unphaseable);

/*
 * The Bayes combine alone, for the call_variants sub-command whose input is a predict FILE
 * (clairs/call_variants.py:798-853): p_aff / p_neg = P(positive class) per head as doubles [n, n_heads]
 * (the python floats parsed from the 8-decimal text), tables as in cto_engine_set_likelihood (host).
 */
int cto_posterior_from_probs(const double* tables_host, int n_heads, const double* p_aff_dev, const double* p_neg_dev,
                             int64_t n, double* post_dev, int32_t* call_dev, void* stream);

/*
 * Dense contractions and the GRU recurrence run on tcgen05 tensor cores by default, as "bf16x3"
 * (every fp32 operand split into bf16 hi + mid, three kind::f16 MMAs per k-step, fp32 accumulate in
 * TMEM) because a single reduced-precision pass breaks the 1e-3 logit contract; enable = 0 forces the
 * fp32 CUDA-core kernels everywhere (the parity tests run both).  cto_gemm_nt exposes the building
 * block itself: C[m,n] = act(A[m,k] * W[n,k]^T + bias) (+ residual), act 0 none / 1 GELU(erf) / 2 SELU.
 * use_tensor_cores: 0 = fp32 CUDA cores; bit 0 = tensor cores, plus (the modes the engine uses between its
 * own kernels) bit 1 = A handed over as pre-split bf16 planes, bit 2 = C produced as bf16 planes, bit 3 =
 * bias indexed by the output row and n-major tile order (the transposed input projections of the GRU), bit 4 (with
 * bit 1) = 128 x 256 output tiles when the problem has enough of them (ignored otherwise), bit 5 (with bits 1 and 3,
 * k <= 256) = CTA pairs with A resident in shared memory (csrc/gemm_pair.cu; ignored when the shape does not fit).
 */
int cto_engine_set_tensor_cores(cto_engine* e, int mode);
/* cto_predict / cto_run_sites_host run the AFF network on a second stream beside NEG (default on; always off while
 * cto_engine_profile is on).  enable = 0: both on the caller's stream, back to back.  Results are identical. */
int cto_engine_set_overlap(cto_engine* e, int enable);
/*
 * mode 1 (default) additionally runs all transformer layers of a CvT stage (clairs/model.py:134-147) in ONE tcgen05
 * kernel that keeps the stage's residual stream in tensor memory (csrc/aff_fused.cu); mode 2 = tensor cores with the
 * round-1 kernel-per-op layers (kept for A/B parity tests).  There is no shape-driven fallback from the tensor-core
 * engine to the CUDA-core kernels: an unsupported network configuration is an error.
 * cto_engine_fused_status: synchronises and copies the fused kernel's watchdog record (int32[8]; [0] != 0 means an
 * in-kernel barrier wait timed out: [0] barrier code, [1] CTA, [2] thread, [3] parity).
 */
int cto_engine_fused_status(cto_engine* e, int32_t* out8);
/*
 * Test hook: size of one NEG workspace tensor and, when dst_dev is not NULL, a device-to-device copy of it as the last
 * forward left it (synchronises).  which: 0 = the transposed
 * input projection of the last GRU layer run (fp32 [6H][33 * bp_max]), 1 / 2 = hi / mid bf16 planes of the layer-1 output
 * (time-major [33][bp][2 H1]), 3 / 4 = hi / mid planes of the layer-2 output (batch-major [n][33][2 H2]).
 */
int cto_engine_workspace(cto_engine* e, int which, void* dst_dev, int64_t* bytes);
/*
 * Kernel-level building block, like cto_gemm_nt: all transformer layers (x += Attention(LN(x)); x += FF(LN(x)),
 * clairs/model.py:143-147) of CvT stage `stage` (0-based) on a residual stream x fp32 [n, W_stage, C_stage]
 * (channels-last, device memory, in place), on whichever engine mode is selected.  n <= max_batch.
 */
int cto_aff_stage_layers(cto_engine* e, int stage, float* x_dev, int64_t n, void* stream);
/*
 * Kernel-level building block: the second GRU layer's recurrence alone (h_t = GRU(x_t, h_{t-1}) in both directions,
 * torch.nn.GRU as used by clairs/model.py:440-470), with the loaded NEG network's recurrent weights.
 * xproj_dev: the layer's input projection W_ih x_t + b_ih (+ b_hr, b_hz), TRANSPOSED: fp32 [6H rows (direction, gate r|z|n,
 * unit)][33 * bp columns (t * bp + candidate)], bp = n rounded up to 128.  out_hi / out_mid: the bf16 split of
 * h_t, [n][33][2H] (forward | backward).  two_chains: 0 = one recurrence chain per CTA pair (csrc/gru_tc3.cu),
 * 1 = two chains per CTA pair (csrc/gru_tc4.cu, H = 192 only; what the engine runs in mode 1).  Tensor-core engine only.
 */
int cto_neg_recurrence(cto_engine* e, const float* xproj_dev, int64_t n, uint16_t* out_hi_dev, uint16_t* out_mid_dev,
                       int two_chains, void* stream);
int cto_gemm_nt(const float* a_dev, int64_t lda, const float* w_dev, const float* bias_dev, const float* residual_dev,
                int64_t ldr, float* c_dev, int64_t ldc, int64_t m, int n, int k, int act, int use_tensor_cores,
                void* stream);

/*
 * Instrumentation for bench.py: kernels launched by this library so far in this process, and
 * optional CUDA-event timing of the kernel families of the forward passes (events are recorded
 * on the launching stream; cto_engine_profile_read synchronises and returns, per family, the
 * summed milliseconds, the number of timed launches and the algorithmic FLOP per candidate).
 */
int64_t cto_launch_count(void);
int cto_engine_profile(cto_engine* e, int enable);
int cto_engine_profile_kinds(void);
const char* cto_engine_profile_name(int kind);
int cto_engine_profile_read(cto_engine* e, double* ms, int64_t* launches, double* flops_per_candidate);

/*
 * Tuning / profiling knobs (not part of the drop-in surface): cto_debug_set(1) + cto_debug_timing(buf) make
 * CTA 0 of the GEMM kernel (slots [0,32)) and cluster 0 of the GRU kernel (slots [32,64)) record per-phase
 * clock64() counters into a device buffer [64 x int64] (profiles/phase_timing_*.py print them).  Results stay correct.
 * Flag bits other than bit 0 select timing-attribution kernel variants that compute WRONG results; they exist only in
 * a debug library built with CTO_DEBUG_KNOBS=1 and are ignored by the release libcto_b200.so.
 */
void cto_debug_set(int flags);
void cto_debug_timing(long long* dev_buf);
void cto_debug_timing_fused(long long* dev_buf);   /* [32] phase counters of CTA 0 of the fused AFF layer kernel */

/* Strand-count recovery of clairs/predict.py:626-642 from the un-rescaled AFF tensor: int32 [n,4] x2. */
int cto_strand_counts(const int16_t* x_aff_dev, int64_t n, int32_t* fwd_dev, int32_t* rev_dev, void* stream);

/*
 * The per-mini-batch body of predict() (clairs/predict.py:610-699) for n candidates at once:
 * rescale both tensors, NEG and AFF forward, softmax, strand counts, posterior, QUAL / FILTER bits.
 * Nullable outputs are skipped.  x_neg_dev == x_aff_dev is allowed (Illumina symlink case,
 * run_clairs_to:1248-1252).
 */
int cto_predict(cto_engine* e, const int16_t* x_aff_dev, const int32_t* depth_aff_dev, const int16_t* x_neg_dev,
                const int32_t* depth_neg_dev, int64_t n, float* logits_aff_dev, float* logits_neg_dev,
                float* probs_dev, double* post_dev, int32_t* call_dev, int32_t* fwd_dev, int32_t* rev_dev,
                double* qual_dev, int32_t* filter_dev, void* stream);

/*
 * One packed stream in HOST memory (see clairs_to_b200/pileup_format.py, cto_pack_reads).
 */
typedef struct cto_host_stream {
    const uint8_t* planes;
    const int32_t* grp_off;
    const uint8_t* ref_code;
    const int32_t* ind_off;
    const uint32_t* ind_entry;
    const int32_t* win_pos;
    int64_t n_groups, n_rows, n_ind;
} cto_host_stream;

/*
 * End-to-end host call: packed host streams in -> host probabilities / posteriors out.  Copies both streams to the
 * device span by span (overlapped with the kernels of the previous chunk), encodes, predicts and copies results back;
 * synchronises before returning.  neg == NULL reuses the AFF stream (Illumina symlink case, run_clairs_to:1248-1252).
 * Outputs (host, nullable): probs fp32 [n, 2H, 2], post double [n, H], call int32 [n], qual double [n], filter int32 [n]
 * (cto_engine_set_qual_thresholds), tensor_aff / tensor_neg int16 [n, 33, 34].
 */
int cto_run_sites_host(cto_engine* e, const cto_host_stream* aff, const cto_host_stream* neg, int64_t n_candidates,
                       float* probs_host, double* post_host, int32_t* call_host, double* qual_host, int32_t* filter_host,
                       int16_t* tensor_aff_host, int16_t* tensor_neg_host, void* stream);

/*
 * Host tokenizer for `samtools mpileup` text (src/create_tensor_pileup_calling.py:120-144, 472-497):
 * parses rows "chr pos ref depth bases BQ MQ" into the read arrays above, and builds the alt_info
 * strings (ibid. 158-209) of candidate rows.  Two-phase: create -> query sizes -> export -> destroy.
 * The text is split at row boundaries over n_threads host threads (<= 0: all cores); rows are independent.
 */
typedef struct cto_tokens cto_tokens;
int cto_tokenize_mpileup(const char* text, int64_t text_len, const char* ref_seq, int64_t ref_len,
                         int64_t ref_start, const int64_t* candidate_pos, int64_t n_candidates,
                         int max_indel_length, int n_threads, cto_tokens** out);
int cto_tokens_sizes(const cto_tokens* t, int64_t* n_reads, int64_t* n_rows, int64_t* n_ind, int64_t* alt_info_bytes);
int cto_tokens_export(const cto_tokens* t, uint8_t* code, uint8_t* bq, uint8_t* mq, int32_t* pos_off,
                      uint8_t* ref_code, int32_t* ind_off, uint32_t* ind_entry, int64_t* row_pos,
                      char* alt_info, int64_t* alt_info_off);
void cto_tokens_destroy(cto_tokens* t);
/*
 * The inverse of the tokenizer, for synthetic workloads (bench.py's text-in leg, tests): read arrays -> mpileup text rows
 * "ctg pos N depth bases BQ MQ" (position = first_pos + row).  ind_len / ind_seq: per indel-carrying read, in read
 * order, the indel length and the inserted bases packed two bits each.  Returns bytes written, -1 if cap is too small.
 */
int64_t cto_render_mpileup(const uint8_t* code, const uint8_t* bq, const uint8_t* mq, const int32_t* pos_off, int64_t n_rows,
                           const uint32_t* ind_entry, const int64_t* ind_len, const int64_t* ind_seq, const char* ctg,
                           int64_t first_pos, char* out, int64_t cap);

/*
 * Text codec for the chunk files: the 1122-int tensor field of a tensor_can row
 * (src/create_tensor_pileup_calling.py:551) and the "%0.8f" probability fields of a predict row
 * (clairs/predict.py:121-132).  Return bytes written (excluding NUL), or -1 if `cap` is too small.
 */
int64_t cto_format_tensor_row(const int16_t* tensor_row, char* out, int64_t cap);
int64_t cto_format_prob_fields(const float* probs, int n_pairs, char* out, int64_t cap);
int cto_parse_tensor_row(const char* text, int64_t len, int16_t* tensor_row);

/*
 * Whole chunk files at once.
 * cto_parse_tensor_file: `text` = a decompressed tensor_can file (rows of src/create_tensor_pileup_calling.py:561-568).
 * Replaces the row loop of tensor_generator_from (clairs/predict.py:172-175, 179, 219-220): rows with fewer than 7
 * tab-separated fields are skipped, rows whose centre reference base is not ACGT are dropped.  For each kept row r:
 * tensor[r] = the 1122 ints, depth[r] = the leading number of alt_info, fields[r][k] = (byte offset, length) in `text`
 * of field k (contig, position, ref33, tensor, alt_info, variant_type, ref_centre; int64 [rows][7][2]).
 * cto_format_predict_rows: the predict-file rows of clairs/predict.py:114-152 for n kept rows (strand counts
 * int32 [n,4] x2 as recovered by cto_strand_counts, probabilities float32 [n, 2*n_heads, 2]); returns the bytes
 * written, -1 if `cap` is too small, -2 on a bad argument.
 */
/*
 * The rows of a tensor_can chunk file (src/create_tensor_pileup_calling.py:561-568) for n candidates of one contig:
 * "ctg \t pos \t ref33 \t 1122 ints \t alt_info \t variant_type \t ref33[16] \n".  ref33: n x 33 characters;
 * alt_off / type_off: int64 [n][2] = (byte offset, length) of the row's alt_info / variant type inside `blob`.
 * Returns the bytes written, -1 if `cap` is too small, -2 on a NULL argument.
 */
int64_t cto_format_tensor_can_rows(const char* ctg, int64_t ctg_len, int64_t n, const int64_t* pos, const char* ref33,
                                   const int16_t* tensor, const char* blob, const int64_t* alt_off, const int64_t* type_off,
                                   char* out, int64_t cap);
int cto_parse_tensor_file(const char* text, int64_t len, int64_t max_rows, int16_t* tensor, int32_t* depth, int64_t* fields,
                          int64_t* n_rows);
int64_t cto_format_predict_rows(const char* text, const int64_t* fields, int64_t n, const int32_t* fwd, const int32_t* rev,
                                const float* probs, int n_heads, char* out, int64_t cap);
/*
 * cto_parse_predict_file: `text` = a decompressed predict file (rows of clairs/predict.py:114-152).  Replaces the row loop
 * of clairs/call_variants.py:798-829: fields[r][k] = (byte offset, length) of chrom, pos, ref, alt_info, forward and
 * reverse strand list-reprs (int64 [rows][6][2]); p_aff / p_neg double [rows][n_heads] = P(positive class) of every head,
 * parsed with strtod (the doubles python's float() gives the reference).
 */
int cto_parse_predict_file(const char* text, int64_t len, int n_heads, int64_t max_rows, double* p_aff, double* p_neg,
                           int64_t* fields, int64_t* n_rows);

/*
 * Candidate scan -- STEP 1 of the pipeline, SURVEY.md section 8 row f3.  Replaces the per-row work of
 * src/extract_candidates_calling.py: the tokenizer and counters of decode_pileup_bases (ibid. 73-120), its allele-frequency /
 * coverage tests (122-144) and the SNV / indel candidate rules of extract_pair_candidates (335-377), for EVERY row of a
 * chunk's `samtools mpileup` text, on the GPU (one thread per row).
 *
 * cto_index_rows: byte offsets of the rows of a '\n'-separated text in device memory.  row_off_dev int64 [cap_rows + 1]
 * receives row_off[0 .. n_rows] (row_off[n_rows] = text_len; a last row without '\n' counts); NULL = count only.
 * Synchronises the stream (the row count comes back to the host).
 *
 * cto_scan_candidates: per row r, pos_dev[r] = column 2, depth_dev[r] = the reference's `depth` (ibid. 104-109), flags_dev[r]:
 *   CTO_CAND_VALID     the reference base of the row is A/C/G/T (ibid. 341-342; other rows are skipped)
 *   CTO_CAND_PASS_AF   pass_af (ibid. 144): the position enters candidates_set / the .bed file
 *   CTO_CAND_SNV       member of snv_candidates_set (ibid. 366-371)
 *   CTO_CAND_INDEL     member of indel_candidates_set (ibid. 372-377; only with select_indel_candidates)
 *   CTO_CAND_MALFORMED fewer than five columns, CTO_CAND_BAD_REF position outside [ref_start, ref_start + ref_len),
 *   CTO_CAND_OVERFLOW  more than 8192 distinct indel alleles in one row (the reference would raise / cannot happen with
 *                      samtools' --max-depth 8000) -- the three are errors for the caller.
 * ref_dev: upper- or lower-case reference bases of positions ref_start .. ref_start + ref_len - 1 (1-based, the
 * `reference_sequence` / `reference_start` of ibid. 286-297).  alternative_base_num < 0 stands for None (ibid. 134-137).
 * Rows with more than 24 distinct indel alleles take a second launch; *n_overflow (nullable) reports how many did.
 * Contract: text_len < 4 GiB per call; cto_index_rows reads 16-byte blocks, so text_dev should be ALLOCATED up to the next
 * multiple of 16 bytes beyond text_len (any alignment works).  Synchronises the stream.
 *
 * cto_scan_candidates_host: the same from HOST memory (text as samtools wrote it, preferably pinned): the text is cut
 * into 32 MB pieces at row ends, piece k + 1 is copied while piece k is indexed and scanned; outputs are host arrays of
 * cap_rows entries, *n_rows the rows found.
 */
#define CTO_CAND_VALID 1
#define CTO_CAND_PASS_AF 2
#define CTO_CAND_SNV 4
#define CTO_CAND_INDEL 8
#define CTO_CAND_MALFORMED 32
#define CTO_CAND_BAD_REF 64
#define CTO_CAND_OVERFLOW 128
int cto_index_rows(const uint8_t* text_dev, int64_t text_len, int64_t* row_off_dev, int64_t cap_rows, int64_t* n_rows, void* stream);
int cto_scan_candidates(const uint8_t* text_dev, int64_t text_len, const int64_t* row_off_dev, int64_t n_rows, const uint8_t* ref_dev,
                        int64_t ref_start, int64_t ref_len, double min_coverage, double snv_min_af, double indel_min_af,
                        int alternative_base_num, int select_indel_candidates, int32_t* pos_dev, int32_t* depth_dev,
                        uint8_t* flags_dev, int32_t* n_overflow, void* stream);
int cto_scan_candidates_host(const char* text, int64_t text_len, const char* ref, int64_t ref_start, int64_t ref_len,
                             double min_coverage, double snv_min_af, double indel_min_af, int alternative_base_num,
                             int select_indel_candidates, int64_t cap_rows, int32_t* pos, int32_t* depth, uint8_t* flags,
                             int64_t* n_rows, int64_t* n_overflow, void* stream);

/*
 * Device tokenizer -- SURVEY.md section 8 row f2: `samtools mpileup` text in DEVICE memory -> the encoder's packed input,
 * without the host tokenizer.  Replaces the row split and the tokenizer of src/create_tensor_pileup_calling.py:472-497,
 * 120-144 like cto_tokenize_mpileup + cto_pack_reads do on host threads, with identical output arrays (tests compare them
 * byte for byte); the alt_info strings (ibid. 158-209) are NOT produced here -- callers that write tensor_can files keep the
 * host tokenizer for the candidate rows.
 *
 * cto_tokenize_count: per row position (row_pos_dev int32 [n_rows]), reference code (ref_code_dev uint8 [n_rows], A for every
 * non-ACGT base like evc_base_from), and the offsets grp_off_dev / ind_off_dev (int32 [n_rows + 1]: groups of eight reads,
 * indel-carrying reads).  *n_groups / *n_ind = the totals the caller sizes planes (8 * n_groups bytes, allocated up to a
 * multiple of 16) and ind_entry (n_ind uint32) with.  row_off_dev from cto_index_rows.  Synchronises the stream.
 * cto_tokenize_write: fills planes_dev and ind_entry_dev (layout: clairs_to_b200/pileup_format.py, PackedStream).
 * Errors (non-zero return, message names the first row): fewer than 7 columns, position outside [ref_start, ref_start +
 * ref_len), an empty line, more than 8192 distinct indel alleles in one row.
 * cto_window_table: win_pos_dev[c * 33 + s] = index of the row at position cand_pos[c] - 16 + s, -1 if samtools printed none
 * (ibid. 461, 513-516); row_pos_dev ascending.
 */
int cto_tokenize_count(const uint8_t* text_dev, int64_t text_len, const int64_t* row_off_dev, int64_t n_rows, const uint8_t* ref_dev,
                       int64_t ref_start, int64_t ref_len, int32_t* row_pos_dev, uint8_t* ref_code_dev, int32_t* grp_off_dev,
                       int32_t* ind_off_dev, int64_t* n_groups, int64_t* n_ind, void* stream);
int cto_tokenize_write(const uint8_t* text_dev, int64_t text_len, const int64_t* row_off_dev, int64_t n_rows, const uint8_t* ref_dev,
                       int64_t ref_start, int64_t ref_len, int low_bq_cut, int max_indel_length, const int32_t* grp_off_dev,
                       const int32_t* ind_off_dev, uint8_t* planes_dev, uint32_t* ind_entry_dev, void* stream);
int cto_window_table(const int32_t* row_pos_dev, int64_t n_rows, const int64_t* cand_pos_dev, int64_t n_candidates,
                     int32_t* win_pos_dev, void* stream);

/*
 * Per-site hard filters -- SURVEY.md section 8 row f4.  Replaces, per called variant ("site"), the reference's
 * _haplotype_build_state_and_line + _haplotype_finalize_line (src/haplotype_filtering.py:570-703, 344-565, cited as HF; long
 * reads, phased) and _postfilter_build_state_and_line + _postfilter_finalize_line (src/postfilter_variants.py:368-446, 278-365,
 * PV; short reads): read start/end, variant cluster ("co-exist"), strand bias (Fisher exact, HF:60-98), sequence entropy
 * (HF:101-151), low alt BQ / MQ, and the haplotype consistency tests against nearby germline variants.
 *
 * Two stages.  (1) cto_hf_parse (host): the chunk's `samtools mpileup --output-MQ --output-QNAME [--output-extra HP]` text
 * (HF:303-311, PV:262-268) -> integer arrays: every string of the reference (read key, upper-cased token, raw indel suffix)
 * becomes a dense id.  Rows must be in increasing position order (samtools' order).  (2) cto_hard_filter_sites (device): one
 * thread block per site walks the rows of the site's window (pos - flanking .. pos + flanking) and evaluates every test with
 * integer counters; the Fisher p-value reproduces the reference's arithmetic (exact big-integer quotient, correctly rounded,
 * then the multiply / divide walk in double precision) and the entropy its order of additions.
 *
 * cto_hf_parse: `ref` = upper-case reference bases of positions region_lo .. region_lo + ref_len - 1 (the chunk_ref of
 * HF:1096-1099).  with_phasing: the text has the ninth (HP) column.  cto_hf_sizes: sizes[8] = rows, entries, start/end
 * entries, reads, tokens, token bytes, suffixes, suffix bytes.  cto_hf_export: copies the arrays out (any pointer may be NULL):
 *   row_pos[rows], row_off[rows + 1] (entry range of a row), row_flags[rows] (CTO_HF_ROW_*), rse_off[rows + 1] / rse_ent[]
 *   (entry indices of the row's read start/end set), per entry rid / tok / sfx (ids), info (CTO_HF_* bits, haplotype tag in
 *   bits 0-1, suffix length incl. its sign in bits 8-23), qual (bq | mq << 8); tok_off[tokens + 1] + tok_blob and
 *   sfx_off[suffixes + 1] + sfx_blob = the interned strings (suffix 0 = none).
 */
#define CTO_HF_HAP_MASK 3u      /* HP tag of the read in this row: 0 (none), 1, 2 */
#define CTO_HF_REV 4u           /* reverse strand (read key ends in '_1') */
#define CTO_HF_STAR 8u          /* token is '*' or '#' */
#define CTO_HF_IS_REF 16u       /* token equals the reference base of the row */
#define CTO_HF_PLUS 32u         /* suffix starts with '+' */
#define CTO_HF_MINUS 64u        /* '-' occurs in the suffix */
#define CTO_HF_SHADOW 128u      /* an earlier entry of a read key that occurs again in the row (dict(zip()) keeps the last) */
#define CTO_HF_LEN_SHIFT 8
#define CTO_HF_ROW_REF_OK 1     /* the row's position lies inside the reference passed to cto_hf_parse */
#define CTO_HF_ROW_RSE 2        /* len(read_start_end_set) >= len(base_list) * 0.2 (HF:627) */
#define CTO_HF_ROW_COUNTER 4    /* the row's Counter is kept for the variant-cluster test (HF:690-696) */
typedef struct cto_hf_chunk cto_hf_chunk;
int cto_hf_parse(const char* text, int64_t len, int with_phasing, const char* ref, int64_t ref_len, int64_t region_lo,
                 cto_hf_chunk** out);
/* the same on n_threads host threads (0 = up to 8, one for texts under 8 MB): row ranges parsed side by side, ids translated when
 * the ranges are appended -- the arrays are identical to the single-threaded ones */
int cto_hf_parse_mt(const char* text, int64_t len, int with_phasing, const char* ref, int64_t ref_len, int64_t region_lo,
                    int n_threads, cto_hf_chunk** out);
int cto_hf_sizes(const cto_hf_chunk* chunk, int64_t* sizes);
int cto_hf_export(const cto_hf_chunk* chunk, int32_t* row_pos, int32_t* row_off, uint8_t* row_flags, int32_t* rse_off,
                  int32_t* rse_ent, int32_t* rid, int32_t* tok, int32_t* sfx, uint32_t* info, uint16_t* qual, int32_t* tok_off,
                  char* tok_blob, int32_t* sfx_off, char* sfx_blob);
void cto_hf_free(cto_hf_chunk* chunk);

/*
 * cto_hard_filter_sites: every pointer is DEVICE memory.  `chunk` = the arrays of cto_hf_export; `sites` = per site:
 *   row_lo / row_hi   rows [row_lo, row_hi) of the chunk lie in the site's window; centre_row = the row of the site (-1: none)
 *   kind              0 SNV, 1 insertion, 2 deletion, 3 other (HF:357-359); alt_tok = token id the alt reads carry (SNV: the
 *                     alt base; insertion: alt[0] + '+' + alt[1:]; -1 = no such token in the chunk); del_len = len(ref_base)
 *   low_af            af < 0.1 (SNV) / 0.3 (indel), HF:380-385, evaluated by the caller on the VCF's AF
 *   rid_min / rid_span  the read ids of the window's rows lie in [rid_min, rid_min + rid_span)
 *   seq_off / seq_len  the site's entropy sequence (HF:147-148: 33 reference bases, IUPAC codes 0-3) inside `seq`
 *   ph_off            rows (ascending) whose HP tags define the reads' haplotypes: the site's row and the heterozygous
 *                     germline positions of the window (HF:621-625): ph_row[ph_off[s] .. ph_off[s + 1])
 *   het_off / hom_off  germline records of the site: het_idx[het_off[s] ..), hom_idx[hom_off[s] ..)
 * germline record g: g_row[g] = its row (-1: position not in the chunk), g_match[g_off[g] + k] for entry k of that row:
 * bit 0 = the read carries the allele by the heterozygous rule (HF:444-451), bit 1 by the homozygous rule (HF:473-481).
 * mode 1 = haplotype filtering (HF), 0 = post filtering (PV).  entropy_tab[35] = e * log(e), e = i / 33 (HF:106-109),
 * entropy_mul = -1 / log(33): computed by the CALLER with the host's libm, so that the sum matches the reference bit for bit.
 * Outputs: out_flags[s] = CTO_HFO_* bits, out_p[s] = the Fisher p-value, out_counts[s][8] = a0, r0, a1, r1 (HF:536-541),
 * match_count, ins_length, depth, |alt read set|.  scratch: uint32 words for sites whose rid_span exceeds the shared-memory table
 * (scratch_off[s] = word offset, -1 = none needed).
 */
#define CTO_HFO_VERDICT 1u
#define CTO_HFO_PHASEABLE 2u
#define CTO_HFO_HETERO 4u
#define CTO_HFO_HOMO 8u
#define CTO_HFO_READ_START_END 16u
#define CTO_HFO_BQ 32u
#define CTO_HFO_MQ 64u
#define CTO_HFO_CO_EXIST 128u
#define CTO_HFO_BOTH_SIDE 256u
#define CTO_HFO_STRAND_BIAS 512u
#define CTO_HFO_ENTROPY 1024u
#define CTO_HF_SMEM_READS 16384  /* reads of one window whose state fits the shared-memory table */
typedef struct {
    int64_t n_rows, n_entries;
    const int32_t *row_pos, *row_off, *rse_off, *rse_ent, *rid, *tok;
    const uint8_t* row_flags;
    const uint32_t* info;
    const uint16_t* qual;
} cto_hf_chunk_arrays;
typedef struct {
    int64_t n_sites;
    const int32_t *row_lo, *row_hi, *centre_row, *alt_tok, *del_len, *rid_min, *rid_span, *seq_off, *ph_off, *ph_row, *het_off,
        *het_idx, *hom_off, *hom_idx;
    const uint8_t *kind, *low_af, *seq_len, *seq;
    const int64_t* scratch_off;
    int64_t n_germline;
    const int32_t *g_row, *g_off;
    const uint8_t* g_match;
} cto_hf_site_arrays;
int cto_hard_filter_sites(const cto_hf_chunk_arrays* chunk, const cto_hf_site_arrays* sites, int mode, int flanking,
                          int disable_read_start_end_filtering, int max_co_exist_read_num, const double* entropy_tab,
                          double entropy_mul, double entropy_threshold, uint32_t* scratch, uint32_t* out_flags, double* out_p,
                          int32_t* out_counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLAIRS_TO_B200_H */
