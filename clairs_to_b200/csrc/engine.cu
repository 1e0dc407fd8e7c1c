// Weight blob walking, workspace and layer-by-layer forward of AFF / NEG.
// Reference structure: clairs/model.py:134-147 (Transformer), 194-198 (stage), 231-261 (CvT.forward),
// 440-467 (BiGRU_NACGT.forward); hyper-parameters come from the blob's cfg, not from constants.
#include "engine.cuh"
#include <algorithm>

namespace cto {

namespace {
struct Walker {
    const float* base;
    int64_t off = 0, total;
    const float* take(int64_t n) {
        const float* p = base + off;
        off += (n + SEG_ALIGN - 1) & ~int64_t(SEG_ALIGN - 1);   // segments are padded to 8 floats (bf16 copies stay 16-byte aligned)
        return p;
    }
};

enum SplitKind { SPLIT_NONE = 0, SPLIT_BF16 = 1 };

int upload(const float* host, int64_t n, WeightSet& ws, int split) {
    ws.n = n;
    CTO_CHECK(cudaMalloc(&ws.blob, sizeof(float) * n));
    CTO_CHECK(cudaMemcpy(ws.blob, host, sizeof(float) * n, cudaMemcpyHostToDevice));
    if (split & SPLIT_BF16) {
        CTO_CHECK(cudaMalloc(&ws.bhi, sizeof(uint16_t) * n));
        CTO_CHECK(cudaMalloc(&ws.bmid, sizeof(uint16_t) * n));
        if (int rc = launch_split_bf16(ws.blob, ws.bhi, ws.bmid, n, nullptr)) return rc;
    }
    CTO_CHECK(cudaDeviceSynchronize());
    return 0;
}

void release(WeightSet& ws) {
    if (ws.blob) cudaFree(ws.blob);
    if (ws.bhi) cudaFree(ws.bhi);
    if (ws.bmid) cudaFree(ws.bmid);
    ws = WeightSet();
}

void walk_head(Walker& w, HeadW& h, int feat, int n_heads) {
    h.fc1_w = w.take((int64_t)FC_DIM * feat);
    h.fc1_b = w.take(FC_DIM);
    h.fc2_w = w.take((int64_t)n_heads * FC_DIM * FC_DIM);
    h.fc2_b = w.take((int64_t)n_heads * FC_DIM);
    h.fc3_w = w.take((int64_t)n_heads * 2 * FC_DIM);
    h.fc3_b = w.take((int64_t)n_heads * 2);
}
}  // namespace

int aff_load(AffModel& m, const float* host_blob, int64_t n, const int32_t* cfg, int cfg_len) {
    CTO_REQUIRE(cfg_len >= 2 && cfg_len == 2 + 3 * cfg[1], "aff cfg: expected [n_heads, n_stages, (C, heads, depth)*]");
    m.n_heads = cfg[0];
    m.n_stages = cfg[1];
    CTO_REQUIRE(m.n_stages >= 1 && m.n_stages <= 3, "aff cfg: %d stages unsupported", m.n_stages);
    CTO_REQUIRE(m.n_heads == 4 || m.n_heads == 6, "aff cfg: %d heads (expected 4 or 6)", m.n_heads);
    if (upload(host_blob, n, m.ws, SPLIT_BF16)) return 1;
    Walker w{m.ws.blob, 0, n};
    int cin = N_CH, win = N_POS;
    for (int s = 0; s < m.n_stages; ++s) {
        CvtStage& st = m.st[s];
        st.c = cfg[2 + 3 * s];
        st.heads = cfg[3 + 3 * s];
        st.depth = cfg[4 + 3 * s];
        CTO_REQUIRE(st.depth <= MAX_DEPTH && (st.c == 16 || st.c == 32 || st.c == 64 || st.c == 128), "aff cfg: stage %d C=%d depth=%d unsupported (C must be 16, 32, 64 or 128)",
                    s, st.c, st.depth);
        st.cin = cin;
        st.win = win;
        st.wout = (win + 1) / 2;                 // 3-tap, stride 2, pad 1
        st.wkv = (st.wout + 1) / 2;
        const int c = st.c, inner = st.heads * DIM_HEAD;
        st.embed_w = w.take((int64_t)c * 3 * cin);
        st.embed_b = w.take(c);
        st.ln_g = w.take(c);
        st.ln_b = w.take(c);
        for (int d = 0; d < st.depth; ++d) {
            CvtLayer& L = st.layers[d];
            L.ln1_g = w.take(c);
            L.ln1_b = w.take(c);
            L.q_dw = w.take(3 * c);
            L.q_pw = w.take((int64_t)inner * c);
            L.q_bias = w.take(inner);
            L.kv_dw = w.take(3 * c);
            L.kv_pw = w.take((int64_t)2 * inner * c);
            L.kv_bias = w.take(2 * inner);
            L.out_w = w.take((int64_t)c * inner);
            L.out_b = w.take(c);
            L.ln2_g = w.take(c);
            L.ln2_b = w.take(c);
            L.ff1_w = w.take((int64_t)4 * c * c);
            L.ff1_b = w.take(4 * c);
            L.ff2_w = w.take((int64_t)4 * c * c);
            L.ff2_b = w.take(c);
        }
        m.sz_x = std::max<int64_t>(m.sz_x, (int64_t)st.wout * c);
        m.sz_kvin = std::max<int64_t>(m.sz_kvin, (int64_t)st.wkv * c);
        m.sz_q = std::max<int64_t>(m.sz_q, (int64_t)st.wout * inner);
        m.sz_kv = std::max<int64_t>(m.sz_kv, (int64_t)st.wkv * 2 * inner);
        m.sz_ff = std::max<int64_t>(m.sz_ff, (int64_t)st.wout * 4 * c);
        m.sz_col = std::max<int64_t>(m.sz_col, (int64_t)st.wout * 3 * cin);
        cin = c;
        win = st.wout;
    }
    m.feat = cin * win;
    walk_head(w, m.head, m.feat, m.n_heads);
    CTO_REQUIRE(w.off == n, "aff blob: %lld floats given, layout needs %lld", (long long)n, (long long)w.off);
    for (int s = 0; s < m.n_stages; ++s)
        if (aff_layers_fused_supported(m.st[s]) && !(s == 0 && aff_stage1_fused_supported(m.st[s])))
            if (int rc = aff_fused_prepare(m.st[s], host_blob, m.ws.blob)) return rc;
    return 0;
}

int neg_load(NegModel& m, const float* host_blob, int64_t n, const int32_t* cfg, int cfg_len) {
    CTO_REQUIRE(cfg_len == 4, "neg cfg: expected [n_heads, in_dim, H1, H2]");
    m.n_heads = cfg[0];
    CTO_REQUIRE(m.n_heads == 4 || m.n_heads == 6, "neg cfg: %d heads (expected 4 or 6)", m.n_heads);
    CTO_REQUIRE(cfg[1] == N_CH, "neg cfg: input dim %d != %d", cfg[1], N_CH);
    if (upload(host_blob, n, m.ws, SPLIT_BF16)) return 1;
    Walker w{m.ws.blob, 0, n};
    int in_dim = cfg[1];
    for (int l = 0; l < 2; ++l) {
        GruLayerW& g = m.l[l];
        g.in_dim = in_dim;
        g.hidden = cfg[2 + l];
        const int h = g.hidden;
        g.wih = w.take((int64_t)6 * h * in_dim);
        g.bih = w.take(6 * h);
        g.whh_t = w.take((int64_t)2 * h * 3 * h);
        g.bhn = w.take(2 * h);
        in_dim = 2 * h;
    }
    walk_head(w, m.head, N_POS * in_dim, m.n_heads);
    CTO_REQUIRE(w.off == n, "neg blob: %lld floats given, layout needs %lld", (long long)n, (long long)w.off);
    {   // K-padded copy of the layer-1 input projection
        const int rows = 6 * m.l[0].hidden, k = m.l[0].in_dim;
        std::vector<float> pad((size_t)rows * NEG_IN_LD, 0.0f);
        const float* src = host_blob + (m.l[0].wih - m.ws.blob);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < k; ++c) pad[(size_t)r * NEG_IN_LD + c] = src[(size_t)r * k + c];
        if (upload(pad.data(), (int64_t)pad.size(), m.wih1_pad, SPLIT_BF16)) return 1;
    }
    for (int l = 0; l < 2; ++l) {
        // W_hh rows regrouped for the CTA-pair recurrence (gru_tc3.cu): (32-unit block, 16-unit half, gate, unit), so
        // that each CTA of the pair streams the 48 rows (r,z,n x 16 units) whose accumulator columns land in its
        // half of the TMEM lanes
        const int h = m.l[l].hidden;
        CTO_REQUIRE(h == 128 || h == 192, "neg cfg: GRU hidden size %d has no tensor-core recurrence kernel (128 / 192 built)", h);
        const float* wt = host_blob + (m.l[l].whh_t - m.ws.blob);      // [2][H (k)][3H (gate*H + unit)]
        std::vector<float> blk((size_t)2 * 3 * h * h);
        for (int d = 0; d < 2; ++d)
            for (int b = 0; b < h / 32; ++b)
                for (int hf = 0; hf < 2; ++hf)
                    for (int g = 0; g < 3; ++g)
                        for (int j = 0; j < 16; ++j) {
                            const int row = b * 96 + hf * 48 + g * 16 + j, col = g * h + b * 32 + hf * 16 + j;
                            for (int k = 0; k < h; ++k)
                                blk[((size_t)d * 3 * h + row) * h + k] = wt[((size_t)d * h + k) * 3 * h + col];
                        }
        if (upload(blk.data(), (int64_t)blk.size(), m.whh_pair[l], SPLIT_BF16)) return 1;
    }
    m.fuse_l1 = m.l[0].hidden == 128 && m.l[0].in_dim < NEG_IN_LD;
    if (m.fuse_l1) {   // W_ih of layer 1 in the row order of whh_pair, K = 64, the (folded) input bias in column in_dim
        const int h = m.l[0].hidden, k = m.l[0].in_dim;
        const float* wih = host_blob + (m.l[0].wih - m.ws.blob);       // [2 * 3H (dir, gate, unit)][in_dim]
        const float* bih = host_blob + (m.l[0].bih - m.ws.blob);
        std::vector<float> win((size_t)2 * 3 * h * 64, 0.0f);
        for (int d = 0; d < 2; ++d)
            for (int b = 0; b < h / 32; ++b)
                for (int hf = 0; hf < 2; ++hf)
                    for (int g = 0; g < 3; ++g)
                        for (int j = 0; j < 16; ++j) {
                            const size_t row = (size_t)d * 3 * h + b * 96 + hf * 48 + g * 16 + j;
                            const size_t src = (size_t)d * 3 * h + g * h + b * 32 + hf * 16 + j;
                            for (int c = 0; c < k; ++c) win[row * 64 + c] = wih[src * k + c];
                            win[row * 64 + k] = bih[src];
                        }
        if (upload(win.data(), (int64_t)win.size(), m.win_pair, SPLIT_BF16)) return 1;
    }
    return 0;
}

namespace {
int dev_alloc(Engine& e, float** p, int64_t n_floats) {
    void* q = nullptr;
    CTO_CHECK(cudaMalloc(&q, sizeof(float) * std::max<int64_t>(n_floats, 4)));
    e.allocs.push_back(q);
    *p = static_cast<float*>(q);
    return 0;
}
}  // namespace

int engine_alloc(Engine& e, int64_t max_batch) {
    e.max_batch = max_batch;
    const int64_t b = max_batch;
    const AffModel& a = e.aff;
    const NegModel& g = e.neg;
    const int64_t xin = (int64_t)N_POS * N_CH;
    int rc = 0;
    rc |= dev_alloc(e, &e.x_aff, b * xin);
    rc |= dev_alloc(e, &e.x_neg, b * N_POS * NEG_IN_LD);
    rc |= dev_alloc(e, &e.a_t0, b * a.sz_x);
    rc |= dev_alloc(e, &e.a_xs, b * a.sz_x);
    rc |= dev_alloc(e, &e.a_y, b * a.sz_x);
    rc |= dev_alloc(e, &e.a_dq, b * a.sz_x);
    rc |= dev_alloc(e, &e.a_dkv, b * a.sz_kvin);
    rc |= dev_alloc(e, &e.a_q, b * a.sz_q);
    rc |= dev_alloc(e, &e.a_kv, b * a.sz_kv);
    rc |= dev_alloc(e, &e.a_att, b * a.sz_q);
    rc |= dev_alloc(e, &e.a_ff, b * a.sz_ff);
    {
        const int64_t sizes[6] = {b * a.sz_x, b * a.sz_kvin, b * a.sz_q, b * a.sz_x, b * a.sz_ff, b * a.sz_col};
        uint16_t** ptrs[6] = {e.p_dq, e.p_dkv, e.p_att, e.p_y, e.p_ff, e.p_col};
        for (int i = 0; i < 6 && !rc; ++i)
            for (int h = 0; h < 2 && !rc; ++h) {
                float* p = nullptr;
                rc |= dev_alloc(e, &p, (sizes[i] + 1) / 2);
                ptrs[i][h] = reinterpret_cast<uint16_t*>(p);
            }
    }
    e.bp_max = (b + 127) / 128 * 128;
    const int64_t xp = (int64_t)N_POS * 6 * std::max(g.l[0].hidden, g.l[1].hidden);
    rc |= dev_alloc(e, &e.n_xp, e.bp_max * xp);
    {   // zeroed once: rows of padded candidates are never written, they only have to stay finite
        const int64_t rows = N_POS * e.bp_max;
        const int64_t sizes[6] = {rows * NEG_IN_LD, rows * NEG_IN_LD, rows * 2 * g.l[0].hidden, rows * 2 * g.l[0].hidden,
                                  rows * 2 * g.l[1].hidden, rows * 2 * g.l[1].hidden};
        uint16_t** ptrs[6] = {&e.nx_hi, &e.nx_mid, &e.o1_hi, &e.o1_mid, &e.o2_hi, &e.o2_mid};
        for (int i = 0; i < 6 && !rc; ++i) {
            float* p = nullptr;
            rc |= dev_alloc(e, &p, (sizes[i] + 1) / 2);
            if (!rc) CTO_CHECK(cudaMemset(p, 0, sizeof(uint16_t) * sizes[i]));
            *ptrs[i] = reinterpret_cast<uint16_t*>(p);
        }
    }
    rc |= dev_alloc(e, &e.n_o1, b * N_POS * 2 * g.l[0].hidden);
    rc |= dev_alloc(e, &e.n_o2, b * N_POS * 2 * g.l[1].hidden);
    rc |= dev_alloc(e, &e.f1, b * FC_DIM);
    rc |= dev_alloc(e, &e.f2, b * 6 * FC_DIM);
    rc |= dev_alloc(e, &e.f1n, b * FC_DIM);
    rc |= dev_alloc(e, &e.f2n, b * 6 * FC_DIM);
    {
        float* d = nullptr;
        rc |= dev_alloc(e, &d, 8);
        e.fused_dbg = reinterpret_cast<int*>(d);
        if (!rc) CTO_CHECK(cudaMemset(e.fused_dbg, 0, sizeof(int) * 8));
    }
    if (rc) return 1;
    return 0;
}

void engine_free(Engine& e) {
    for (void* p : e.allocs) cudaFree(p);
    e.allocs.clear();
    if (e.copy_stream) cudaStreamDestroy(e.copy_stream);
    e.copy_stream = nullptr;
    if (e.aux_stream) cudaStreamDestroy(e.aux_stream);
    if (e.fork_ev) cudaEventDestroy(e.fork_ev);
    if (e.join_ev) cudaEventDestroy(e.join_ev);
    e.aux_stream = nullptr; e.fork_ev = e.join_ev = nullptr;
    if (e.pool) cudaMemPoolDestroy(e.pool);
    e.pool = nullptr;
    for (int s = 0; s < 3; ++s) aff_fused_release(e.aff.st[s]);
    release(e.aff.ws);
    release(e.neg.ws);
    release(e.neg.wih1_pad);
    release(e.neg.win_pair);
    release(e.neg.whh_pair[0]);
    release(e.neg.whh_pair[1]);
    if (e.tables) cudaFree(e.tables);
    for (auto& r : e.prof) { cudaEventDestroy(r.start); cudaEventDestroy(r.stop); }
    for (cudaEvent_t ev : e.ev_free) cudaEventDestroy(ev);
    e.prof.clear();
    e.ev_free.clear();
}

const char* prof_kind_name(int kind) {
    static const char* names[PK_COUNT] = {"aff_forward", "aff_stage1_fused", "aff_gemm_1x1", "aff_embed_conv", "aff_channel_ln", "aff_dwconv", "aff_attention", "aff_layers_fused", "aff_layers_fused_last",
                                          "aff_heads", "neg_proj1_gemm", "neg_gru1_recurrent", "neg_proj2_gemm",
                                          "neg_gru2_recurrent", "neg_fc1_gemm", "neg_heads"};
    return (kind >= 0 && kind < PK_COUNT) ? names[kind] : "?";
}

// AFF FLOP per candidate (multiply-add = 2 FLOP, 3-tap convolutions, attention products included):
// part 0 = everything, 1 = the transformer layers of the stages that run in the fused kernel (4 = those of the LAST stage, which
// is a different kernel instantiation and reported apart), 2 = first stage when it runs in the CUDA-core fused kernel, 3 = heads
static double aff_flops(const Engine& e, int part) {
    const AffModel& m = e.aff;
    double total = 0.0, layers_fused = 0.0, layers_last = 0.0, stage1 = 0.0;
    for (int s = 0; s < m.n_stages; ++s) {
        const CvtStage& st = m.st[s];
        const double c = st.c, inner = st.heads * DIM_HEAD, w = st.wout, wkv = st.wkv;
        const double embed = 2.0 * w * 3 * st.cin * c;
        const double layer = 2.0 * (w * c * inner + wkv * c * 2 * inner + 2.0 * st.heads * w * wkv * DIM_HEAD + w * inner * c + 2.0 * w * c * 4 * c);
        total += embed + st.depth * layer;
        if (st.fused_stream) (s == m.n_stages - 1 ? layers_last : layers_fused) += st.depth * layer;
        if (s == 0 && aff_stage1_fused_supported(st)) stage1 = embed + st.depth * layer;
    }
    const double heads = 2.0 * (m.feat * FC_DIM + m.n_heads * (FC_DIM * FC_DIM + 2 * FC_DIM));
    total += heads;
    return part == 0 ? total : (part == 1 ? layers_fused : (part == 2 ? stage1 : (part == 4 ? layers_last : heads)));
}

// multiply-add = 2 FLOP; matches SURVEY.md section 8(d) accounting
double prof_kind_flops_per_candidate(const Engine& e, int kind) {
    const GruLayerW* l = e.neg.l;
    const double t = N_POS;
    switch (kind) {
        case PK_AFF: return aff_flops(e, 0);
        case PK_AFF_LAYERS: return e.use_tc && e.use_fused ? aff_flops(e, 1) : 0.0;
        case PK_AFF_LAYERS_LAST: return e.use_tc && e.use_fused ? aff_flops(e, 4) : 0.0;
        case PK_AFF_STAGE1: return aff_flops(e, 2);
        case PK_AFF_HEADS: return aff_flops(e, 3);
        case PK_NEG_PROJ1: return (e.neg.fuse_l1 && e.use_tc) ? 0.0 : 2.0 * t * l[0].in_dim * 6 * l[0].hidden;   // fused: counted under PK_NEG_GRU1, this family is the input split
        case PK_NEG_GRU1: return 2.0 * t * l[0].hidden * 6 * l[0].hidden + (e.neg.fuse_l1 && e.use_tc ? 2.0 * t * l[0].in_dim * 6 * l[0].hidden : 0.0);
        case PK_NEG_PROJ2: return 2.0 * t * l[1].in_dim * 6 * l[1].hidden;
        case PK_NEG_GRU2: return 2.0 * t * l[1].hidden * 6 * l[1].hidden;
        case PK_NEG_FC1: return 2.0 * t * 2 * l[1].hidden * FC_DIM;
        case PK_NEG_HEADS: return 2.0 * e.neg.n_heads * (FC_DIM * FC_DIM + 2 * FC_DIM);
        default: return 0.0;
    }
}

static int take_event(Engine& e, cudaEvent_t* ev) {
    if (!e.ev_free.empty()) {
        *ev = e.ev_free.back();
        e.ev_free.pop_back();
        return 0;
    }
    CTO_CHECK(cudaEventCreate(ev));
    return 0;
}

static bool prof_wanted(const Engine& e, int kind) {
    if (!e.profile) return false;
    const bool fine = kind >= PK_AFF_STAGE1 && kind <= PK_AFF_HEADS;
    return e.profile >= 2 ? kind != PK_AFF : !fine;
}

int prof_begin(Engine& e, int kind, cudaStream_t s) {
    e.prof_open = prof_wanted(e, kind);
    if (!e.prof_open) return 0;
    Engine::ProfRec r{kind, nullptr, nullptr};
    if (take_event(e, &r.start) || take_event(e, &r.stop)) return 1;
    CTO_CHECK(cudaEventRecord(r.start, s));
    e.prof.push_back(r);
    return 0;
}

int prof_end(Engine& e, cudaStream_t s) {
    if (!e.prof_open) return 0;
    e.prof_open = false;
    CTO_CHECK(cudaEventRecord(e.prof.back().stop, s));
    return 0;
}

int prof_collect(Engine& e, double* ms, int64_t* count) {
    for (int k = 0; k < PK_COUNT; ++k) { ms[k] = 0.0; count[k] = 0; }
    for (auto& r : e.prof) {
        CTO_CHECK(cudaEventSynchronize(r.stop));
        float t = 0.0f;
        CTO_CHECK(cudaEventElapsedTime(&t, r.start, r.stop));
        ms[r.kind] += t;
        count[r.kind] += 1;
        e.ev_free.push_back(r.start);
        e.ev_free.push_back(r.stop);
    }
    e.prof.clear();
    return 0;
}

#define RUN(x) do { if (int _rc = (x)) return _rc; } while (0)

// dense contraction: tcgen05 bf16x3 on the tensor-core engine, fp32 CUDA cores on the exact engine
// (cto_engine_set_tensor_cores(e, 0)).  There is NO shape-driven fallback between the two: on the tensor-core engine a
// shape the tcgen05 kernel cannot take is an error, except for the one product that is CUDA-core BY DESIGN
// (`cuda_core_by_design`: the first stage's embed convolution, K = 3 * 34 = 102 is not a multiple of 8 and N <= 32).
static int gemm(const Engine& e, const AView& a, const float* w, const float* bias, const float* residual, int64_t ldr,
                float* c, int64_t ldc, int64_t m, int n, int k, int act, cudaStream_t s, bool cuda_core_by_design = false) {
    if (!e.use_tc || cuda_core_by_design) return launch_gemm_nt(a, w, bias, residual, ldr, c, ldc, m, n, k, act, s);
    const WeightSet* ws = e.aff.ws.owns(w) ? &e.aff.ws : (e.neg.ws.owns(w) ? &e.neg.ws : nullptr);
    CTO_REQUIRE(!a.conv && ws && (w - ws->blob) % SEG_ALIGN == 0 && gemm_tc_supported(a.ptr, a.lda, w, m, n, k, c, ldc, residual, ldr),
                "tensor-core engine: no tcgen05 kernel for the product m=%lld n=%d k=%d (conv rows: %d); this network "
                "configuration is not supported (use cto_engine_set_tensor_cores(e, 0) for the exact fp32 engine)",
                (long long)m, n, k, a.conv);
    const int64_t off = w - ws->blob;
    return launch_gemm_tc(a.ptr, a.lda, ws->bhi + off, ws->bmid + off, bias, residual, ldr, c, ldc, m, n, k, act, s);
}

static int run_heads(const Engine& e, const HeadW& h, const float* feat, int feat_dim, int n_heads, int64_t n, float* f1, float* f2,
                     float* logits, cudaStream_t s) {
    // fc1 -> SELU -> per-head fc2 -> SELU -> fc3 -> SELU  (clairs/model.py:239-253)
    RUN(gemm(e, plain_a(feat, feat_dim), h.fc1_w, h.fc1_b, nullptr, 0, f1, FC_DIM, n, FC_DIM, feat_dim, ACT_SELU, s));
    RUN(gemm(e, plain_a(f1, FC_DIM), h.fc2_w, h.fc2_b, nullptr, 0, f2, (int64_t)n_heads * FC_DIM, n,
                       n_heads * FC_DIM, FC_DIM, ACT_SELU, s));
    RUN(launch_head_fc3(f2, h.fc3_w, h.fc3_b, logits, n, n_heads, s));
    return 0;
}

// one launch timed under a profile kind (no-op unless the engine profiles that kind)
#define TIMED(kind, call)            \
    do {                             \
        RUN(prof_begin(e, kind, s)); \
        RUN(call);                   \
        RUN(prof_end(e, s));         \
    } while (0)

// All transformer layers of one stage on the residual stream e.a_xs [n, wout, C], in place (clairs/model.py:143-147).
static int aff_stage_layers(Engine& e, const CvtStage& st, int64_t n, cudaStream_t s) {
    const AffModel& m = e.aff;
    const int c = st.c, inner = st.heads * DIM_HEAD;
    const int64_t rows = n * st.wout, rows_kv = n * st.wkv;
    // every GEMM of a stage whose width is a multiple of 64 runs on the tensor cores, so the tensors between the
    // kernels that only feed a GEMM travel as bf16 hi / mid planes (no converter pass in the GEMM)
    if (e.use_tc && e.use_fused && st.fused_stream) {
        // every transformer layer of the stage in one tcgen05 kernel, x stays in tensor memory (aff_fused.cu)
        TIMED((&st == &m.st[m.n_stages - 1]) ? PK_AFF_LAYERS_LAST : PK_AFF_LAYERS, launch_aff_layers(st, e.a_xs, n, e.fused_dbg, s));
        return 0;
    }
    const bool planes = e.use_tc && c % 64 == 0;
    for (int d = 0; d < st.depth && planes; ++d) {
        const CvtLayer& L = st.layers[d];
        const WeightSet& ws = m.ws;
        auto pg = [&](uint16_t* const* ap, int k, const float* w, const float* bias, int nn, int64_t mm) {
            GemmTc g;
            g.flags = GEMM_A_PRESPLIT;
            g.a_hi = ap[0]; g.a_mid = ap[1]; g.lda = k;
            g.w_hi = ws.bhi + (w - ws.blob); g.w_mid = ws.bmid + (w - ws.blob); g.ldw = k;
            g.bias = bias; g.m = mm; g.n = nn; g.k = k;
            return g;
        };
        TIMED(PK_AFF_DWCONV, launch_ln_dwconv(e.a_xs, L.ln1_g, L.ln1_b, L.q_dw, L.kv_dw, nullptr, nullptr, n, st.wout, st.wkv, c,
                                              s, e.p_dq[0], e.p_dq[1], e.p_dkv[0], e.p_dkv[1]));
        GemmTc g = pg(e.p_dq, c, L.q_pw, L.q_bias, inner, rows);
        g.c = e.a_q; g.ldc = inner;
        TIMED(PK_AFF_GEMM, launch_gemm_tc_ex(g, s));
        g = pg(e.p_dkv, c, L.kv_pw, L.kv_bias, 2 * inner, rows_kv);
        g.c = e.a_kv; g.ldc = 2 * inner;
        TIMED(PK_AFF_GEMM, launch_gemm_tc_ex(g, s));
        TIMED(PK_AFF_ATTENTION, launch_attention(e.a_q, e.a_kv, nullptr, n, st.wout, st.wkv, st.heads, s, e.p_att[0], e.p_att[1]));
        g = pg(e.p_att, inner, L.out_w, L.out_b, c, rows);
        g.c = e.a_xs; g.ldc = c; g.residual = e.a_xs; g.ldr = c;
        TIMED(PK_AFF_GEMM, launch_gemm_tc_ex(g, s));
        TIMED(PK_AFF_LN, launch_channel_ln(e.a_xs, L.ln2_g, L.ln2_b, nullptr, rows, c, s, e.p_y[0], e.p_y[1]));
        g = pg(e.p_y, c, L.ff1_w, L.ff1_b, 4 * c, rows);
        g.flags |= GEMM_OUT_SPLIT; g.c_hi = e.p_ff[0]; g.c_mid = e.p_ff[1]; g.ldc = 4 * c; g.act = ACT_GELU;
        TIMED(PK_AFF_GEMM, launch_gemm_tc_ex(g, s));
        g = pg(e.p_ff, 4 * c, L.ff2_w, L.ff2_b, c, rows);
        g.c = e.a_xs; g.ldc = c; g.residual = e.a_xs; g.ldr = c;
        TIMED(PK_AFF_GEMM, launch_gemm_tc_ex(g, s));
    }
    for (int d = 0; d < st.depth && !planes; ++d) {
        const CvtLayer& L = st.layers[d];
        // x = Attention(LN(x)) + x   (clairs/model.py:145); LN and both depth-wise convs are one kernel
        TIMED(PK_AFF_DWCONV, launch_ln_dwconv(e.a_xs, L.ln1_g, L.ln1_b, L.q_dw, L.kv_dw, e.a_dq, e.a_dkv, n, st.wout,
                                              st.wkv, c, s));
        TIMED(PK_AFF_GEMM, gemm(e, plain_a(e.a_dq, c), L.q_pw, L.q_bias, nullptr, 0, e.a_q, inner, rows, inner, c,
                                ACT_NONE, s));
        TIMED(PK_AFF_GEMM, gemm(e, plain_a(e.a_dkv, c), L.kv_pw, L.kv_bias, nullptr, 0, e.a_kv, 2 * inner, rows_kv,
                                2 * inner, c, ACT_NONE, s));
        TIMED(PK_AFF_ATTENTION, launch_attention(e.a_q, e.a_kv, e.a_att, n, st.wout, st.wkv, st.heads, s));
        TIMED(PK_AFF_GEMM, gemm(e, plain_a(e.a_att, inner), L.out_w, L.out_b, e.a_xs, c, e.a_xs, c, rows, c, inner,
                                ACT_NONE, s));
        // x = FF(LN(x)) + x          (clairs/model.py:146)
        TIMED(PK_AFF_LN, launch_channel_ln(e.a_xs, L.ln2_g, L.ln2_b, e.a_y, rows, c, s));
        TIMED(PK_AFF_GEMM, gemm(e, plain_a(e.a_y, c), L.ff1_w, L.ff1_b, nullptr, 0, e.a_ff, 4 * c, rows, 4 * c, c, ACT_GELU, s));
        TIMED(PK_AFF_GEMM, gemm(e, plain_a(e.a_ff, 4 * c), L.ff2_w, L.ff2_b, e.a_xs, c, e.a_xs, c, rows, c, 4 * c,
                                ACT_NONE, s));
    }
    return 0;
}

// Kernel-level parity hook (cto_aff_stage_layers): the layers of stage `si` on a caller-provided residual stream.
// Kernel-level hook (cto_neg_recurrence): the layer-2 recurrence alone, on a caller-supplied transposed input projection.
int neg_recurrence_on(Engine& e, const float* xproj, int64_t n, uint16_t* out_hi, uint16_t* out_mid, int two_chains, cudaStream_t s) {
    NegModel& m = e.neg;
    CTO_REQUIRE(m.n_heads > 0 && m.whh_pair[1].bhi && e.use_tc, "neg_recurrence: needs a loaded NEG network on the tensor-core engine");
    CTO_REQUIRE(n > 0 && n <= e.max_batch, "neg_recurrence: n %lld outside (0, max_batch]", (long long)n);
    const int h2 = m.l[1].hidden;
    const int64_t bp = (n + 127) / 128 * 128, ldx = (int64_t)N_POS * bp;
    if (two_chains) {
        CTO_REQUIRE(h2 == 192, "neg_recurrence: the two-chain kernel is built for hidden size 192, not %d", h2);
        return launch_gru4(xproj, ldx, bp, m.whh_pair[1].bhi, m.whh_pair[1].bmid, m.l[1].bhn, out_hi, out_mid, N_POS, 1, n, h2, e.fused_dbg, s);
    }
    return launch_gru3(xproj, ldx, bp, m.whh_pair[1].bhi, m.whh_pair[1].bmid, m.l[1].bhn, out_hi, out_mid, N_POS, 1, n, h2, s);
}

int aff_stage_layers_on(Engine& e, int si, float* x, int64_t n, cudaStream_t s) {
    CTO_REQUIRE(si >= 0 && si < e.aff.n_stages && n <= e.max_batch, "aff_stage_layers: stage %d / batch %lld out of range", si, (long long)n);
    const CvtStage& st = e.aff.st[si];
    const size_t bytes = sizeof(float) * (size_t)n * st.wout * st.c;
    CTO_CHECK(cudaMemcpyAsync(e.a_xs, x, bytes, cudaMemcpyDeviceToDevice, s));
    RUN(aff_stage_layers(e, st, n, s));
    CTO_CHECK(cudaMemcpyAsync(x, e.a_xs, bytes, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int aff_forward(Engine& e, const float* x, int64_t n, float* logits, cudaStream_t s) {
    CTO_REQUIRE(n <= e.max_batch, "aff_forward: batch %lld > engine max_batch %lld", (long long)n, (long long)e.max_batch);
    const AffModel& m = e.aff;
    const float* cur = x;
    cudaEvent_t aff_stop = nullptr;
    if (e.profile == 1) {                      // AFF as one block (its parts are only timed at level 2)
        RUN(prof_begin(e, PK_AFF, s));
        aff_stop = e.prof.back().stop;
        e.prof_open = false;
    }
    for (int si = 0; si < m.n_stages; ++si) {
        const CvtStage& st = m.st[si];
        const int c = st.c;
        const int64_t rows = n * st.wout;
        if (si == 0 && e.use_tc && aff_stage1_fused_supported(st)) {
            TIMED(PK_AFF_STAGE1, launch_aff_stage1(st, cur, e.a_xs, n, s));
            cur = e.a_xs;
            continue;
        }
        // embed conv (3-tap, stride 2, pad 1) + channel LN  (clairs/model.py:195-196)
        if (e.use_tc && c % 64 == 0 && (3 * st.cin) % 8 == 0) {
            // tensor-core stages: explicit conv rows as split planes (one small kernel), then a pre-split GEMM
            RUN(prof_begin(e, PK_AFF_EMBED, s));
            RUN(launch_im2col3_split(cur, n, st.win, st.wout, st.cin, e.p_col[0], e.p_col[1], s));
            GemmTc g;
            g.flags = GEMM_A_PRESPLIT;
            g.a_hi = e.p_col[0]; g.a_mid = e.p_col[1]; g.lda = 3 * st.cin;
            g.w_hi = m.ws.bhi + (st.embed_w - m.ws.blob); g.w_mid = m.ws.bmid + (st.embed_w - m.ws.blob); g.ldw = 3 * st.cin;
            g.bias = st.embed_b; g.c = e.a_t0; g.ldc = c; g.m = rows; g.n = c; g.k = 3 * st.cin;
            RUN(launch_gemm_tc_ex(g, s));
            RUN(prof_end(e, s));
        } else
        TIMED(PK_AFF_EMBED, gemm(e, conv_a(cur, st.win, st.wout, st.cin), st.embed_w, st.embed_b, nullptr, 0, e.a_t0, c, rows,
                                 c, 3 * st.cin, ACT_NONE, s, /*cuda_core_by_design=*/si == 0));
        TIMED(PK_AFF_LN, launch_channel_ln(e.a_t0, st.ln_g, st.ln_b, e.a_xs, rows, c, s));
        RUN(aff_stage_layers(e, st, n, s));
        // the next stage reads a_xs while writing a_t0, so no copy is needed
        cur = e.a_xs;
    }
    TIMED(PK_AFF_HEADS, run_heads(e, m.head, cur, m.feat, m.n_heads, n, e.f1, e.f2, logits, s));
    if (aff_stop) CTO_CHECK(cudaEventRecord(aff_stop, s));
    return 0;
}

// Tensor-core NEG path.  Both input projections are computed TRANSPOSED, xproj^T[6H, 33*bp] = W_ih * X^T, so that
// the recurrence kernel reads them with fully coalesced loads (gru_tc3.cu); every GEMM operand on this path is
// already split into bf16 hi / mid planes by its producer, so the GEMM kernel runs without its converter warps.
// x: fp32 [n, 33, NEG_IN_LD], or nullptr when the caller has already filled the input planes e.nx_hi / e.nx_mid
// (neg_forward_from_counts: rescale and split in one kernel)
static int neg_forward_tc(Engine& e, const float* x, int64_t n, float* logits, cudaStream_t s) {
    const NegModel& m = e.neg;
    const int64_t bp = (n + 127) / 128 * 128, ldx = (int64_t)N_POS * bp;
    CTO_REQUIRE(bp <= e.bp_max, "neg_forward: batch %lld > workspace", (long long)n);
    const int h1 = m.l[0].hidden, h2 = m.l[1].hidden;
    // the input planes carry a constant 1.0 in column in_dim when layer 1 runs fused (its bias rides in W_ih)
    if (x) RUN(launch_split_time_major(x, n, N_POS, NEG_IN_LD, bp, e.nx_hi, e.nx_mid, s, m.fuse_l1 ? m.l[0].in_dim : -1));
    GemmTc g;
    g.flags = GEMM_A_PRESPLIT | GEMM_BIAS_PER_ROW | GEMM_TILES_N_MAJOR | (e.use_two_chains ? GEMM_WIDE_N : 0);
    g.c = e.n_xp; g.ldc = ldx; g.n = (int)ldx;
    if (m.fuse_l1) {
        RUN(prof_begin(e, PK_NEG_GRU1, s));
        RUN(launch_gru1_fused(e.nx_hi, e.nx_mid, NEG_IN_LD, bp, m.win_pair.bhi, m.win_pair.bmid, m.whh_pair[0].bhi,
                              m.whh_pair[0].bmid, m.l[0].bhn, e.o1_hi, e.o1_mid, 1, bp, n, h1, s));
        RUN(prof_end(e, s));
    } else {
        RUN(prof_begin(e, PK_NEG_PROJ1, s));
        g.a_hi = m.wih1_pad.bhi; g.a_mid = m.wih1_pad.bmid; g.lda = NEG_IN_LD;
        g.w_hi = e.nx_hi; g.w_mid = e.nx_mid; g.ldw = NEG_IN_LD;
        g.bias = m.l[0].bih; g.m = 6 * h1; g.k = NEG_IN_LD;
        RUN(launch_gemm_tc_ex(g, s));
        RUN(prof_end(e, s));
        RUN(prof_begin(e, PK_NEG_GRU1, s));
        RUN(launch_gru3(e.n_xp, ldx, bp, m.whh_pair[0].bhi, m.whh_pair[0].bmid, m.l[0].bhn, e.o1_hi, e.o1_mid, 1, bp, n, h1, s));
        RUN(prof_end(e, s));
    }
    RUN(prof_begin(e, PK_NEG_PROJ2, s));
    const int64_t off2 = m.l[1].wih - m.ws.blob;
    g.a_hi = m.ws.bhi + off2; g.a_mid = m.ws.bmid + off2; g.lda = 2 * h1;
    g.w_hi = e.o1_hi; g.w_mid = e.o1_mid; g.ldw = 2 * h1;
    g.bias = m.l[1].bih; g.m = 6 * h2; g.k = 2 * h1;
    if (e.use_two_chains && gemm_pair_supported(g)) RUN(launch_gemm_pair(g, s));
    else RUN(launch_gemm_tc_ex(g, s));
    RUN(prof_end(e, s));
    RUN(prof_begin(e, PK_NEG_GRU2, s));
    if (e.use_two_chains && h2 == 192)
        RUN(launch_gru4(e.n_xp, ldx, bp, m.whh_pair[1].bhi, m.whh_pair[1].bmid, m.l[1].bhn, e.o2_hi, e.o2_mid, N_POS, 1, n, h2, e.fused_dbg, s));
    else
        RUN(launch_gru3(e.n_xp, ldx, bp, m.whh_pair[1].bhi, m.whh_pair[1].bmid, m.l[1].bhn, e.o2_hi, e.o2_mid, N_POS, 1, n, h2, s));
    RUN(prof_end(e, s));
    const int feat = N_POS * 2 * h2;
    const HeadW& hd = m.head;
    RUN(prof_begin(e, PK_NEG_FC1, s));
    const int64_t off1 = hd.fc1_w - m.ws.blob;
    GemmTc f;
    f.flags = GEMM_A_PRESPLIT;
    f.a_hi = e.o2_hi; f.a_mid = e.o2_mid; f.lda = feat;
    f.w_hi = m.ws.bhi + off1; f.w_mid = m.ws.bmid + off1; f.ldw = feat;
    f.bias = hd.fc1_b; f.c = e.f1n; f.ldc = FC_DIM; f.m = n; f.n = FC_DIM; f.k = feat; f.act = ACT_SELU;
    RUN(launch_gemm_tc_ex(f, s));
    RUN(prof_end(e, s));
    RUN(prof_begin(e, PK_NEG_HEADS, s));
    RUN(gemm(e, plain_a(e.f1n, FC_DIM), hd.fc2_w, hd.fc2_b, nullptr, 0, e.f2n, (int64_t)m.n_heads * FC_DIM, n,
                       m.n_heads * FC_DIM, FC_DIM, ACT_SELU, s));
    RUN(launch_head_fc3(e.f2n, hd.fc3_w, hd.fc3_b, logits, n, m.n_heads, s));
    return prof_end(e, s);
}

int neg_forward_from_counts(Engine& e, const int16_t* x, const int32_t* depth, int64_t n, float* logits, cudaStream_t s) {
    CTO_REQUIRE(n <= e.max_batch, "neg_forward: batch %lld > engine max_batch %lld", (long long)n, (long long)e.max_batch);
    static_assert(NEG_IN_LD == NEG_PLANE_LD, "plane row stride");
    if (e.use_tc) {
        RUN(prof_begin(e, PK_NEG_PROJ1, s));               // "proj1" family = producing the layer-1 input planes
        RUN(launch_rescale_split_time_major(x, depth, n, (n + 127) / 128 * 128, e.nx_hi, e.nx_mid, s,
                                            e.neg.fuse_l1 ? e.neg.l[0].in_dim : -1));
        RUN(prof_end(e, s));
        return neg_forward_tc(e, nullptr, n, logits, s);
    }
    RUN(launch_rescale(x, depth, n, e.x_neg, NEG_IN_LD, s));
    return neg_forward(e, e.x_neg, n, logits, s);
}

int neg_forward(Engine& e, const float* x, int64_t n, float* logits, cudaStream_t s) {
    CTO_REQUIRE(n <= e.max_batch, "neg_forward: batch %lld > engine max_batch %lld", (long long)n, (long long)e.max_batch);
    const NegModel& m = e.neg;
    if (e.use_tc) return neg_forward_tc(e, x, n, logits, s);
    // fp32 CUDA-core path (cto_engine_set_tensor_cores(e, 0)): row-major projections + thread-per-unit recurrence
    const float* cur = x;
    float* outs[2] = {e.n_o1, e.n_o2};
    for (int l = 0; l < 2; ++l) {
        const GruLayerW& g = m.l[l];
        const int h = g.hidden;
        RUN(prof_begin(e, l ? PK_NEG_PROJ2 : PK_NEG_PROJ1, s));
        RUN(gemm(e, plain_a(cur, l == 0 ? NEG_IN_LD : g.in_dim), g.wih, g.bih, nullptr, 0, e.n_xp, 6 * h, n * N_POS, 6 * h,
                 g.in_dim, ACT_NONE, s));
        RUN(prof_end(e, s));
        RUN(prof_begin(e, l ? PK_NEG_GRU2 : PK_NEG_GRU1, s));
        RUN(launch_gru_recurrent(e.n_xp, g.whh_t, g.bhn, outs[l], n, h, s));
        RUN(prof_end(e, s));
        cur = outs[l];
    }
    const int feat = N_POS * 2 * m.l[1].hidden;
    const HeadW& hd = m.head;
    RUN(prof_begin(e, PK_NEG_FC1, s));
    RUN(gemm(e, plain_a(cur, feat), hd.fc1_w, hd.fc1_b, nullptr, 0, e.f1n, FC_DIM, n, FC_DIM, feat, ACT_SELU, s));
    RUN(prof_end(e, s));
    RUN(prof_begin(e, PK_NEG_HEADS, s));
    RUN(gemm(e, plain_a(e.f1n, FC_DIM), hd.fc2_w, hd.fc2_b, nullptr, 0, e.f2n, (int64_t)m.n_heads * FC_DIM, n,
                       m.n_heads * FC_DIM, FC_DIM, ACT_SELU, s));
    RUN(launch_head_fc3(e.f2n, hd.fc3_w, hd.fc3_b, logits, n, m.n_heads, s));
    return prof_end(e, s);
}

}  // namespace cto
