// Weight layout + forward orchestration for the AFF (CvT) and NEG (BiGRU) networks.
// The flat fp32 weight blobs are produced by clairs_to_b200/weights.py from a reference
// checkpoint's state_dict; both sides walk the SAME segment order below.
#pragma once
#include "nn_kernels.cuh"
#include <vector>

namespace cto {

constexpr int MAX_DEPTH = 16;
constexpr int DIM_HEAD = 64;       // clairs/model.py:103
constexpr int FC_DIM = 128;        // clairs/model.py:214, 419

struct CvtLayer {
    const float *ln1_g, *ln1_b, *q_dw, *q_pw, *q_bias, *kv_dw, *kv_pw, *kv_bias, *out_w, *out_b;
    const float *ln2_g, *ln2_b, *ff1_w, *ff1_b, *ff2_w, *ff2_b;
};

struct CvtStage {
    int c, cin, heads, depth, win, wout, wkv;
    const float *embed_w, *embed_b, *ln_g, *ln_b;
    CvtLayer layers[MAX_DEPTH];
    // fused transformer layers (aff_fused.cu): prepacked weight tile stream, vector table, cumulative biases
    uint8_t* fused_stream = nullptr;
    float* fused_vecs = nullptr;              // per layer a 16 KB block of small vectors (LN, taps, biases, cumulative biases)
    float* fused_cb = nullptr;                // [C] sum of every out-projection / FF2 bias of the stage
    long long fused_layer_bytes = 0;
    int fused_vec_bytes = 0;
};

struct HeadW {
    const float *fc1_w, *fc1_b, *fc2_w, *fc2_b, *fc3_w, *fc3_b;
};

constexpr int NEG_IN_LD = 40;      // NEG input rows are padded 34 -> 40 floats (16-byte row stride for the fp32 x rows
                                   // AND for the bf16 rows of the K-padded W_ih copy)
constexpr int SEG_ALIGN = 8;       // every blob segment starts on a multiple of 8 elements (16 bytes as bf16)

// fp32 weights plus their bf16 hi / mid split (same element offsets in all three arrays) for the bf16x3 tensor-core kernels
struct WeightSet {
    float* blob = nullptr;
    uint16_t *bhi = nullptr, *bmid = nullptr;
    int64_t n = 0;
    bool owns(const float* w) const { return w >= blob && w < blob + n; }
};

struct AffModel {
    int n_heads = 0, n_stages = 0, feat = 0;
    CvtStage st[3];
    HeadW head;
    WeightSet ws;
    // per-candidate workspace sizes (floats)
    int64_t sz_x = 0, sz_kvin = 0, sz_q = 0, sz_kv = 0, sz_ff = 0, sz_col = 0;
};

struct GruLayerW {
    int in_dim, hidden;
    const float *wih, *bih, *whh_t, *bhn;
};

struct NegModel {
    int n_heads = 0;
    GruLayerW l[2];
    HeadW head;
    WeightSet ws;
    WeightSet wih1_pad;        // layer-1 W_ih with K padded 34 -> NEG_IN_LD (zeros), for the tensor-core path
    WeightSet whh_pair[2];     // per layer: W_hh [2 dirs][unit block (32) x half (16) x gate x unit][H], for gru_tc3.cu
    WeightSet win_pair;        // layer 1: W_ih in the same row order, K padded to 64, bias in column in_dim (gru_in_tc.cu)
    bool fuse_l1 = false;      // layer 1 runs with its input projection fused into the recurrence kernel
};

struct Engine {
    AffModel aff;
    NegModel neg;
    int64_t max_batch = 0;
    // workspace (device, fp32)
    float *x_aff = nullptr, *x_neg = nullptr;                  // rescaled inputs: AFF [chunk,33,34], NEG [chunk,33,NEG_IN_LD]
    float *a_t0 = nullptr, *a_xs = nullptr, *a_y = nullptr, *a_dq = nullptr, *a_dkv = nullptr;
    float *a_q = nullptr, *a_kv = nullptr, *a_att = nullptr, *a_ff = nullptr;
    // tensor-core AFF path: bf16 hi / mid planes of every tensor that only feeds a GEMM (index 0 = hi, 1 = mid)
    uint16_t *p_dq[2] = {nullptr, nullptr}, *p_dkv[2] = {nullptr, nullptr}, *p_att[2] = {nullptr, nullptr},
             *p_y[2] = {nullptr, nullptr}, *p_ff[2] = {nullptr, nullptr}, *p_col[2] = {nullptr, nullptr};
    float *n_xp = nullptr, *n_o1 = nullptr, *n_o2 = nullptr;
    // tensor-core NEG path: bf16 hi / mid planes. x and o1 are time-major [33, bp, .] (operands of the transposed
    // input projections), o2 is batch-major [n, 33 * 2H] (A operand of the flattening fc1); bp = chunk rounded up to 128
    uint16_t *nx_hi = nullptr, *nx_mid = nullptr, *o1_hi = nullptr, *o1_mid = nullptr, *o2_hi = nullptr, *o2_mid = nullptr;
    int64_t bp_max = 0;
    float *f1 = nullptr, *f2 = nullptr;                        // head activations (shared sizes)
    float *f1n = nullptr, *f2n = nullptr;
    double* tables = nullptr;                                  // likelihood tables, n_heads * 122
    int table_heads = 0;
    std::vector<void*> allocs;
    cudaStream_t copy_stream = nullptr;                        // H2D spans of cto_run_sites_host overlap the kernels
    cudaMemPool_t pool = nullptr;                              // private stream-ordered pool of cto_run_sites_host
    // AFF and NEG of a chunk only meet at the posterior: cto_predict runs AFF on `aux_stream` beside NEG on the caller's
    // stream (each fills the other's tail waves; profiles/two_stream_overlap.py: 3-7 % per chunk, identical results).  Off
    // while per-kernel profiling is on, so that every timed kernel runs alone.
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    bool overlap_networks = true;
    bool use_tc = true;                                        // dense contractions on tcgen05 (bf16x3)
    bool use_fused = true;                                     // AFF transformer layers in the fused kernel (aff_fused.cu)
    bool use_two_chains = true;                                // GRU layer 2 with two chains per CTA pair (gru_tc4.cu)
    int* fused_dbg = nullptr;                                  // device int[8]: barrier-timeout report of the fused kernel
    // optional per-kernel-family timing with CUDA events on the launching stream (bench.py roofline)
    int profile = 0;           // 0 off, 1 = NEG kernel families + AFF as one block, 2 = every AFF kernel family too
    struct ProfRec { int kind; cudaEvent_t start, stop; };
    std::vector<ProfRec> prof;
    bool prof_open = false;
    std::vector<cudaEvent_t> ev_free;
};

enum ProfKind { PK_AFF = 0, PK_AFF_STAGE1, PK_AFF_GEMM, PK_AFF_EMBED, PK_AFF_LN, PK_AFF_DWCONV, PK_AFF_ATTENTION, PK_AFF_LAYERS, PK_AFF_LAYERS_LAST, PK_AFF_HEADS, PK_NEG_PROJ1, PK_NEG_GRU1, PK_NEG_PROJ2, PK_NEG_GRU2, PK_NEG_FC1, PK_NEG_HEADS, PK_COUNT };
const char* prof_kind_name(int kind);
double prof_kind_flops_per_candidate(const Engine& e, int kind);
int prof_begin(Engine& e, int kind, cudaStream_t s);
int prof_end(Engine& e, cudaStream_t s);
// synchronises, sums elapsed ms and launch counts per kind, clears the records
int prof_collect(Engine& e, double* ms, int64_t* count);

// whole first CvT stage in one kernel (aff_stage1.cu) when the stage has the predict.py shape (C=16, 1 head, depth 1)
// all transformer layers of one stage in one tcgen05 kernel (aff_fused.cu); C = 32 / 64 / 128
bool aff_layers_fused_supported(const CvtStage& st);
int aff_fused_prepare(CvtStage& st, const float* host_blob, const float* dev_blob);
void aff_fused_release(CvtStage& st);
int launch_aff_layers(const CvtStage& st, float* x, int64_t n, int* dbg, cudaStream_t s);
bool aff_stage1_fused_supported(const CvtStage& st);
int launch_aff_stage1(const CvtStage& st, const float* x, float* out, int64_t n, cudaStream_t s);

int aff_load(AffModel& m, const float* host_blob, int64_t n, const int32_t* cfg, int cfg_len);
int neg_load(NegModel& m, const float* host_blob, int64_t n, const int32_t* cfg, int cfg_len);
int engine_alloc(Engine& e, int64_t max_batch);
void engine_free(Engine& e);

// aff: x device fp32 [n, 33, 34]; neg: x device fp32 [n, 33, NEG_IN_LD] (zero padded);
// logits: device fp32 [n, n_heads, 2]; n <= max_batch
int aff_forward(Engine& e, const float* x, int64_t n, float* logits, cudaStream_t s);
int aff_stage_layers_on(Engine& e, int si, float* x, int64_t n, cudaStream_t s);
int neg_recurrence_on(Engine& e, const float* xproj, int64_t n, uint16_t* out_hi, uint16_t* out_mid, int two_chains, cudaStream_t s);
int neg_forward(Engine& e, const float* x, int64_t n, float* logits, cudaStream_t s);
// the same from the encoder's int16 tensor [n, 33, 34] + per-candidate depth (depth rescale of clairs/predict.py:179-197 fused in)
int neg_forward_from_counts(Engine& e, const int16_t* x, const int32_t* depth, int64_t n, float* logits, cudaStream_t s);

}  // namespace cto
