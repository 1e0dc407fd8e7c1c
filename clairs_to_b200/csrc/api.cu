// C ABI (include/clairs_to_b200.h) over the CUDA kernels.
#include "../../include/clairs_to_b200.h"
#include "engine.cuh"
#include <string.h>
#include <algorithm>
#include <new>
#include <vector>
#include <climits>
#include <mutex>

namespace cto {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

static int64_t g_launches = 0;
void count_launch(int n) { __atomic_add_fetch(&g_launches, (int64_t)n, __ATOMIC_RELAXED); }
int64_t launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int device_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    int v = __atomic_load_n(&cached[dev], __ATOMIC_RELAXED);
    if (!v) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        __atomic_store_n(&cached[dev], v, __ATOMIC_RELAXED);
    }
    return v;
}

// Scratch memory of the stand-alone entry points (row index, candidate scan, device tokenizer): a private stream-ordered pool
// per device that KEEPS its memory.  cudaMallocAsync on the default pool hands everything back to the driver at the next
// stream synchronisation (release threshold 0), and these entry points synchronise several times per call: the re-allocation
// cost 5-50 ms per call, at random (profiles/debug_device_tokenizer.py).  The application's default pool is left alone.
cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t s) {
    static cudaMemPool_t pools[64] = {nullptr};
    static std::mutex mu;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!pools[dev]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            e = cudaMemPoolCreate(&pools[dev], &props);
            if (e != cudaSuccess) return e;
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool = pools[dev];
    }
    return cudaMallocFromPoolAsync(p, bytes > 16 ? bytes : 16, pool, s);
}

int launch_encode_pileup(const uint8_t* planes, const int32_t* grp_off, const uint8_t* ref_code, const int32_t* ind_off,
                         const uint32_t* ind_entry, const int32_t* win_pos, int64_t n_candidates, int64_t n_groups,
                         int16_t* tensor, int32_t* depth, cudaStream_t stream);

}  // namespace cto

namespace cto { extern int g_gemm_debug; extern long long* g_gemm_timing; extern long long* g_fused_timing; }
using namespace cto;

struct cto_engine {
    Engine e;
};

namespace {
struct DevStream {
    uint8_t *planes = nullptr, *ref_code = nullptr;
    int32_t *grp_off = nullptr, *ind_off = nullptr, *win_pos = nullptr;
    uint32_t* ind_entry = nullptr;
};

void free_stream(DevStream& d, cudaStream_t s) {
    void* ptrs[] = {d.planes, d.ref_code, d.grp_off, d.ind_off, d.win_pos, d.ind_entry};
    for (void* p : ptrs)
        if (p) cudaFreeAsync(p, s);
}

// Span bookkeeping of the pipelined host call: [lo, hi) of an array that is already on its way to the device.  A new span
// that starts inside it only copies the part beyond `hi` (windows of consecutive chunks overlap; re-copying bytes that
// the previous chunk's kernels are still reading was an unordered write/read pair).
struct Copied {
    int64_t lo = 0, hi = 0;
    // returns the sub-range of [a, b) that still has to be copied and records it
    void claim(int64_t a, int64_t b, int64_t& ca, int64_t& cb) {
        if (hi > lo && a >= lo && a <= hi) {
            ca = hi < b ? hi : b;
            cb = b;
            if (b > hi) hi = b;
        } else {
            ca = a;
            cb = b;
            lo = a;
            hi = b;
        }
    }
};
}  // namespace


extern "C" {

int cto_abi_version(void) { return CTO_ABI_VERSION; }
const char* cto_last_error(void) { return get_error(); }

int cto_device_check(int* sm_count) {
    int dev = 0, major = 0, minor = 0;
    CTO_CHECK(cudaGetDevice(&dev));
    // attribute queries, not cudaGetDeviceProperties: that call takes milliseconds and this check sits in front of every
    // stand-alone entry point (18 ms per call were measured in front of a 0.2 ms kernel)
    CTO_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    CTO_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    CTO_REQUIRE(major == 10, "clairs_to_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    if (sm_count) CTO_CHECK(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    return 0;
}

int cto_encode_pileup(const uint8_t* planes, const int32_t* grp_off, const uint8_t* ref_code, const int32_t* ind_off,
                      const uint32_t* ind_entry, const int32_t* win_pos, int64_t n_candidates, int64_t n_groups,
                      int16_t* tensor, int32_t* depth, void* stream) {
    CTO_REQUIRE(n_candidates >= 0, "encode_pileup: negative candidate count");
    CTO_REQUIRE(n_candidates == 0 || (grp_off && ref_code && ind_off && win_pos && tensor), "encode_pileup: NULL array");
    return launch_encode_pileup(planes, grp_off, ref_code, ind_off, ind_entry, win_pos, n_candidates, n_groups, tensor, depth,
                                (cudaStream_t)stream);
}

int cto_engine_create(const float* aff_blob, int64_t aff_len, const int32_t* aff_cfg, int aff_cfg_len,
                      const float* neg_blob, int64_t neg_len, const int32_t* neg_cfg, int neg_cfg_len, int64_t max_batch,
                      cto_engine** out) {
    CTO_REQUIRE(out && aff_blob && neg_blob && aff_cfg && neg_cfg, "engine_create: NULL argument");
    CTO_REQUIRE(max_batch > 0 && max_batch <= (1 << 20), "engine_create: max_batch %lld out of range", (long long)max_batch);
    if (cto_device_check(nullptr)) return 3;
    cto_engine* h = new (std::nothrow) cto_engine();
    CTO_REQUIRE(h, "engine_create: out of host memory");
    int rc = aff_load(h->e.aff, aff_blob, aff_len, aff_cfg, aff_cfg_len);
    if (!rc) rc = neg_load(h->e.neg, neg_blob, neg_len, neg_cfg, neg_cfg_len);
    if (!rc && h->e.aff.n_heads != h->e.neg.n_heads) {
        set_error("engine_create: AFF has %d heads, NEG has %d", h->e.aff.n_heads, h->e.neg.n_heads);
        rc = 2;
    }
    if (!rc) rc = engine_alloc(h->e, max_batch);
    if (!rc) {
        // private stream-ordered pool for the end-to-end host call: keeps its memory between calls
        int dev = 0;
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        if (cudaGetDevice(&dev) != cudaSuccess) rc = 1;
        props.location.id = dev;
        if (!rc && cudaMemPoolCreate(&h->e.pool, &props) != cudaSuccess) {
            set_error("engine_create: cudaMemPoolCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = 1;
        }
        if (!rc) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(h->e.pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if (rc) {
        engine_free(h->e);
        delete h;
        return rc;
    }
    *out = h;
    return 0;
}

void cto_engine_destroy(cto_engine* h) {
    if (!h) return;
    cudaDeviceSynchronize();
    engine_free(h->e);
    delete h;
}

int cto_engine_heads(const cto_engine* h) { return h ? h->e.aff.n_heads : 0; }

int cto_engine_set_likelihood(cto_engine* h, const double* tables, int n_heads) {
    CTO_REQUIRE(h && tables, "set_likelihood: NULL argument");
    CTO_REQUIRE(n_heads == h->e.aff.n_heads, "set_likelihood: tables for %d heads, engine has %d", n_heads,
                h->e.aff.n_heads);
    if (!h->e.tables) {
        CTO_CHECK(cudaMalloc(&h->e.tables, sizeof(double) * (122 * 6 + 3)));
        CTO_CHECK(cudaMemset(h->e.tables, 0, sizeof(double) * (122 * 6 + 3)));      // thresholds default to 0: every QUAL passes
    }
    CTO_CHECK(cudaMemcpy(h->e.tables, tables, sizeof(double) * 122 * n_heads, cudaMemcpyHostToDevice));
    h->e.table_heads = n_heads;
    return 0;
}

int cto_engine_set_qual_thresholds(cto_engine* h, double qual_pass, double qual_phaseable, double qual_unphaseable) {
    CTO_REQUIRE(h && h->e.tables, "set_qual_thresholds: set the likelihood tables first");
    const double thr[3] = {qual_pass, qual_phaseable, qual_unphaseable};
    CTO_CHECK(cudaMemcpy(h->e.tables + 122 * 6, thr, sizeof(thr), cudaMemcpyHostToDevice));
    return 0;
}

int cto_rescale(const int16_t* x, const int32_t* depth, int64_t n, float* out, void* stream) {
    return launch_rescale(x, depth, n, out, N_CH, (cudaStream_t)stream);
}

int cto_forward_aff(cto_engine* h, const float* x, int64_t n, float* logits, void* stream) {
    CTO_REQUIRE(h && (n == 0 || (x && logits)), "forward_aff: NULL argument");
    const int nh = h->e.aff.n_heads;
    for (int64_t o = 0; o < n; o += h->e.max_batch) {
        const int64_t nb = std::min(h->e.max_batch, n - o);
        if (int rc = aff_forward(h->e, x + o * N_POS * N_CH, nb, logits + o * nh * 2, (cudaStream_t)stream)) return rc;
    }
    return 0;
}

int cto_forward_neg(cto_engine* h, const float* x, int64_t n, float* logits, void* stream) {
    CTO_REQUIRE(h && (n == 0 || (x && logits)), "forward_neg: NULL argument");
    const int nh = h->e.neg.n_heads;
    for (int64_t o = 0; o < n; o += h->e.max_batch) {
        const int64_t nb = std::min(h->e.max_batch, n - o);
        if (int rc = launch_pad_rows(x + o * N_POS * N_CH, nb * N_POS, N_CH, h->e.x_neg, NEG_IN_LD, (cudaStream_t)stream))
            return rc;
        if (int rc = neg_forward(h->e, h->e.x_neg, nb, logits + o * nh * 2, (cudaStream_t)stream)) return rc;
    }
    return 0;
}

int cto_softmax_posterior(cto_engine* h, const float* la, const float* ln, int64_t n, float* probs, double* post,
                          int32_t* call, double* qual, int32_t* filter, void* stream) {
    CTO_REQUIRE(h && (n == 0 || (la && ln)), "softmax_posterior: NULL argument");
    const bool want = post || call || qual || filter;
    const double* tables = want ? h->e.tables : nullptr;
    CTO_REQUIRE(!want || tables, "softmax_posterior: posterior requested but no likelihood tables set");
    return launch_softmax_posterior(la, ln, n, h->e.aff.n_heads, tables, probs, post, call, (cudaStream_t)stream, qual, filter);
}

int cto_posterior_from_probs(const double* tables_host, int n_heads, const double* p_aff_dev, const double* p_neg_dev,
                             int64_t n, double* post_dev, int32_t* call_dev, void* stream) {
    CTO_REQUIRE(tables_host && (n_heads == 4 || n_heads == 6), "posterior_from_probs: bad tables / head count %d", n_heads);
    CTO_REQUIRE(n == 0 || (p_aff_dev && p_neg_dev && post_dev && call_dev), "posterior_from_probs: NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    double* t = nullptr;
    CTO_CHECK(cudaMallocAsync((void**)&t, sizeof(double) * 122 * n_heads, s));
    CTO_CHECK(cudaMemcpyAsync(t, tables_host, sizeof(double) * 122 * n_heads, cudaMemcpyHostToDevice, s));
    const int rc = launch_posterior_from_probs(p_aff_dev, p_neg_dev, n, n_heads, t, post_dev, call_dev, s);
    cudaFreeAsync(t, s);
    return rc;
}

int64_t cto_launch_count(void) { return launches(); }

void cto_debug_set(int flags) {
#ifdef CTO_DEBUG_KNOBS
    cto::g_gemm_debug = flags;
#else
    cto::g_gemm_debug = flags & 1;      // release library: phase counters only; the wrong-result timing variants are not compiled in
#endif
}
void cto_debug_timing(long long* dev_buf) { cto::g_gemm_timing = dev_buf; }
void cto_debug_timing_fused(long long* dev_buf) { cto::g_fused_timing = dev_buf; }
int cto_engine_set_overlap(cto_engine* h, int enable) {
    CTO_REQUIRE(h, "engine_set_overlap: NULL engine");
    h->e.overlap_networks = enable != 0;
    return 0;
}

int cto_engine_set_tensor_cores(cto_engine* h, int mode) {
    CTO_REQUIRE(h, "engine_set_tensor_cores: NULL engine");
    CTO_REQUIRE(mode >= 0 && mode <= 2, "engine_set_tensor_cores: mode %d (0 exact fp32, 1 tensor cores, 2 tensor cores with the round-1 per-op kernels: no fused AFF layers, one-chain GRU)", mode);
    h->e.use_tc = mode != 0;
    h->e.use_fused = mode == 1;
    h->e.use_two_chains = mode == 1;
    return 0;
}

int cto_aff_stage_layers(cto_engine* h, int stage, float* x_dev, int64_t n, void* stream) {
    CTO_REQUIRE(h && x_dev, "aff_stage_layers: NULL argument");
    return aff_stage_layers_on(h->e, stage, x_dev, n, (cudaStream_t)stream);
}

int cto_neg_recurrence(cto_engine* h, const float* xproj_dev, int64_t n, uint16_t* out_hi_dev, uint16_t* out_mid_dev, int two_chains,
                       void* stream) {
    CTO_REQUIRE(h && xproj_dev && out_hi_dev && out_mid_dev, "neg_recurrence: NULL argument");
    return neg_recurrence_on(h->e, xproj_dev, n, out_hi_dev, out_mid_dev, two_chains, (cudaStream_t)stream);
}

int cto_engine_workspace(cto_engine* h, int which, void* dst_dev, int64_t* bytes) {
    const void* src = nullptr;
    const void** dev_ptr = &src;
    CTO_REQUIRE(h && bytes, "engine_workspace: NULL argument");
    Engine& e = h->e;
    CTO_REQUIRE(e.neg.n_heads > 0 && e.n_xp, "engine_workspace: no NEG workspace");
    const int64_t rows = (int64_t)N_POS * e.bp_max, h1 = e.neg.l[0].hidden, h2 = e.neg.l[1].hidden;
    switch (which) {
        case 0: *dev_ptr = e.n_xp; *bytes = 4 * rows * 6 * std::max(h1, h2); break;
        case 1: *dev_ptr = e.o1_hi; *bytes = 2 * rows * 2 * h1; break;
        case 2: *dev_ptr = e.o1_mid; *bytes = 2 * rows * 2 * h1; break;
        case 3: *dev_ptr = e.o2_hi; *bytes = 2 * rows * 2 * h2; break;
        case 4: *dev_ptr = e.o2_mid; *bytes = 2 * rows * 2 * h2; break;
        default: CTO_REQUIRE(false, "engine_workspace: unknown tensor %d", which);
    }
    if (dst_dev) {
        CTO_CHECK(cudaDeviceSynchronize());
        CTO_CHECK(cudaMemcpy(dst_dev, src, (size_t)*bytes, cudaMemcpyDeviceToDevice));
    }
    return 0;
}

int cto_engine_fused_status(cto_engine* h, int32_t* out8) {
    CTO_REQUIRE(h && out8, "engine_fused_status: NULL argument");
    CTO_CHECK(cudaDeviceSynchronize());
    CTO_CHECK(cudaMemcpy(out8, h->e.fused_dbg, sizeof(int32_t) * 8, cudaMemcpyDeviceToHost));
    return 0;
}

int cto_gemm_nt(const float* a, int64_t lda, const float* w, const float* bias, const float* residual, int64_t ldr,
                float* c, int64_t ldc, int64_t m, int n, int k, int act, int use_tensor_cores, void* stream) {
    if (!use_tensor_cores) return launch_gemm_nt(plain_a(a, lda), w, bias, residual, ldr, c, ldc, m, n, k, act, (cudaStream_t)stream);
    // tensor-core modes (bit 0 set): the operands are split here on the fly; the engine does it once at load time
    // (weights) or in the producing kernel (activations).  bit 1: A pre-split into bf16 planes, bit 2: C written as
    // bf16 planes and recombined, bit 3: bias indexed by the output row + n-major tile order, bit 4 (with bit 1): 128 x 256 tiles
    // when the problem has enough column tiles (the transposed GRU input projections), bit 5 (with bits 1 and 3, K <= 256): the
    // CTA-pair kernel with A resident in shared memory (gemm_pair.cu) when the shape fits.
    const bool presplit = use_tensor_cores & 2, out_split = use_tensor_cores & 4, by_row = use_tensor_cores & 8;
    CTO_REQUIRE(!presplit || lda == k, "gemm_nt: pre-split A needs a dense A (lda == k)");
    CTO_REQUIRE(!out_split || (ldc == n && !residual), "gemm_nt: split C needs a dense C (ldc == n) and no residual");
    cudaStream_t s = (cudaStream_t)stream;
    uint16_t *whi = nullptr, *wmid = nullptr, *ahi = nullptr, *amid = nullptr, *chi = nullptr, *cmid = nullptr;
    auto alloc = [&](uint16_t** p, size_t n_elem) { return cudaMallocAsync((void**)p, sizeof(uint16_t) * n_elem, s); };
    CTO_CHECK(alloc(&whi, (size_t)n * k));
    CTO_CHECK(alloc(&wmid, (size_t)n * k));
    int rc = launch_split_bf16(w, whi, wmid, (int64_t)n * k, s);
    GemmTc g;
    g.a = a; g.lda = lda; g.w_hi = whi; g.w_mid = wmid; g.ldw = k; g.bias = bias; g.residual = residual; g.ldr = ldr;
    g.c = c; g.ldc = ldc; g.m = m; g.n = n; g.k = k; g.act = act;
    if (presplit && !rc) {
        CTO_CHECK(alloc(&ahi, (size_t)m * k));
        CTO_CHECK(alloc(&amid, (size_t)m * k));
        rc = launch_split_bf16(a, ahi, amid, m * k, s);
        g.flags |= GEMM_A_PRESPLIT; g.a_hi = ahi; g.a_mid = amid;
    }
    if (out_split && !rc) {
        CTO_CHECK(alloc(&chi, (size_t)m * n));
        CTO_CHECK(alloc(&cmid, (size_t)m * n));
        g.flags |= GEMM_OUT_SPLIT; g.c_hi = chi; g.c_mid = cmid;
    }
    if (by_row) g.flags |= GEMM_BIAS_PER_ROW | GEMM_TILES_N_MAJOR;
    if (use_tensor_cores & 16) g.flags |= GEMM_WIDE_N;
    if (!rc) rc = ((use_tensor_cores & 32) && gemm_pair_supported(g)) ? launch_gemm_pair(g, s) : launch_gemm_tc_ex(g, s);
    if (out_split && !rc) rc = launch_join_bf16(chi, cmid, c, m * n, s);
    for (uint16_t* p : {whi, wmid, ahi, amid, chi, cmid})
        if (p) cudaFreeAsync(p, s);
    return rc;
}

int cto_engine_profile(cto_engine* h, int enable) {
    CTO_REQUIRE(h, "engine_profile: NULL engine");
    h->e.profile = enable < 0 ? 0 : enable;
    return 0;
}

int cto_engine_profile_kinds(void) { return PK_COUNT; }
const char* cto_engine_profile_name(int kind) { return prof_kind_name(kind); }

int cto_engine_profile_read(cto_engine* h, double* ms, int64_t* launches_out, double* flops_per_candidate) {
    CTO_REQUIRE(h && ms && launches_out, "engine_profile_read: NULL argument");
    if (int rc = prof_collect(h->e, ms, launches_out)) return rc;
    if (flops_per_candidate)
        for (int k = 0; k < PK_COUNT; ++k) flops_per_candidate[k] = prof_kind_flops_per_candidate(h->e, k);
    return 0;
}

int cto_strand_counts(const int16_t* x, int64_t n, int32_t* fwd, int32_t* rev, void* stream) {
    CTO_REQUIRE(n == 0 || (x && fwd && rev), "strand_counts: NULL argument");
    return launch_strand_counts(x, n, fwd, rev, (cudaStream_t)stream);
}

int cto_predict(cto_engine* h, const int16_t* x_aff, const int32_t* depth_aff, const int16_t* x_neg,
                const int32_t* depth_neg, int64_t n, float* logits_aff, float* logits_neg, float* probs, double* post,
                int32_t* call, int32_t* fwd, int32_t* rev, double* qual, int32_t* filter, void* stream) {
    CTO_REQUIRE(h && (n == 0 || (x_aff && x_neg && depth_aff && depth_neg && logits_aff && logits_neg)),
                "predict: NULL argument");
    Engine& e = h->e;
    cudaStream_t s = (cudaStream_t)stream;
    const int nh = e.aff.n_heads;
    const int64_t xin = (int64_t)N_POS * N_CH;
    // The two networks share nothing but their inputs until the posterior.  With profiling off AFF runs on a second stream
    // beside NEG (fork / join events around every engine chunk); with profiling on they run back to back on the caller's
    // stream so that the per-kernel events time one kernel at a time.
    const bool overlap = e.overlap_networks && !e.profile && n > 0;
    if (overlap && !e.aux_stream) {
        CTO_CHECK(cudaStreamCreateWithFlags(&e.aux_stream, cudaStreamNonBlocking));
        CTO_CHECK(cudaEventCreateWithFlags(&e.fork_ev, cudaEventDisableTiming));
        CTO_CHECK(cudaEventCreateWithFlags(&e.join_ev, cudaEventDisableTiming));
    }
    for (int64_t o = 0; o < n; o += e.max_batch) {
        const int64_t nb = std::min(e.max_batch, n - o);
        cudaStream_t sa = s;
        if (overlap) {
            sa = e.aux_stream;
            CTO_CHECK(cudaEventRecord(e.fork_ev, s));
            CTO_CHECK(cudaStreamWaitEvent(sa, e.fork_ev, 0));
        }
        if (int rc = neg_forward_from_counts(e, x_neg + o * xin, depth_neg + o, nb, logits_neg + o * nh * 2, s)) return rc;
        if (int rc = launch_rescale(x_aff + o * xin, depth_aff + o, nb, e.x_aff, N_CH, sa)) return rc;
        if (int rc = aff_forward(e, e.x_aff, nb, logits_aff + o * nh * 2, sa)) return rc;
        if (overlap) {
            CTO_CHECK(cudaEventRecord(e.join_ev, sa));
            CTO_CHECK(cudaStreamWaitEvent(s, e.join_ev, 0));
        }
    }
    if (fwd && rev)
        if (int rc = launch_strand_counts(x_aff, n, fwd, rev, s)) return rc;
    if (probs || post || call || qual || filter) {
        const bool want = post || call || qual || filter;
        const double* tables = want ? e.tables : nullptr;
        CTO_REQUIRE(!want || tables, "predict: posterior requested but no likelihood tables set");
        if (int rc = launch_softmax_posterior(logits_aff, logits_neg, n, nh, tables, probs, post, call, s, qual, filter)) return rc;
    }
    return 0;
}

int cto_run_sites_host(cto_engine* h, const cto_host_stream* aff, const cto_host_stream* neg, int64_t n,
                       float* probs_host, double* post_host, int32_t* call_host, double* qual_host, int32_t* filter_host,
                       int16_t* tensor_aff_host, int16_t* tensor_neg_host, void* stream) {
    CTO_REQUIRE(h && aff, "run_sites_host: NULL argument");
    if (n <= 0) return 0;
    Engine& e = h->e;
    cudaStream_t s = (cudaStream_t)stream;
    const int nh = e.aff.n_heads;
    const int64_t xin = (int64_t)N_POS * N_CH;
    if (!e.copy_stream) CTO_CHECK(cudaStreamCreateWithFlags(&e.copy_stream, cudaStreamNonBlocking));
    cudaStream_t cs = e.copy_stream;

    const cto_host_stream* hs[2] = {aff, neg};
    const int n_streams = neg ? 2 : 1;
    DevStream d[2];
    int16_t* tens[2] = {nullptr, nullptr};
    int32_t* dep[2] = {nullptr, nullptr};
    int32_t* call = nullptr;
    float *la = nullptr, *ln = nullptr, *probs = nullptr;
    double *post = nullptr, *qual = nullptr;
    int32_t* flt = nullptr;
    std::vector<cudaEvent_t> events;
    int rc = 0;
    // this call allocates ~8 KB per candidate and frees it at the end: a PRIVATE stream-ordered pool that keeps its
    // memory between calls (the default pool of the host application is left alone)
    auto alloc = [&](void** p, int64_t bytes) {
        if (cudaMallocFromPoolAsync(p, bytes > 16 ? bytes : 16, e.pool, s) != cudaSuccess) rc = 1;
    };
    // device arrays are only ALLOCATED here and filled span by span below (plane bytes, per-row arrays and the window
    // table alike), so that the host->device copy of chunk c+1 overlaps the kernels of chunk c
    for (int k = 0; k < n_streams && !rc; ++k) {
        const cto_host_stream* x = hs[k];
        alloc((void**)&d[k].planes, ((x->n_groups * 8 + 15) & ~int64_t(15)) + 16);
        alloc((void**)&d[k].ind_entry, sizeof(uint32_t) * x->n_ind);
        alloc((void**)&d[k].grp_off, sizeof(int32_t) * (x->n_rows + 1));
        alloc((void**)&d[k].ind_off, sizeof(int32_t) * (x->n_rows + 1));
        alloc((void**)&d[k].ref_code, x->n_rows);
        alloc((void**)&d[k].win_pos, sizeof(int32_t) * n * N_POS);
        alloc((void**)&tens[k], sizeof(int16_t) * n * xin);
        alloc((void**)&dep[k], sizeof(int32_t) * n);
        if (rc) break;
    }
    alloc((void**)&la, sizeof(float) * n * nh * 2);
    alloc((void**)&ln, sizeof(float) * n * nh * 2);
    alloc((void**)&probs, sizeof(float) * n * nh * 4);
    alloc((void**)&post, sizeof(double) * n * nh);
    alloc((void**)&call, sizeof(int32_t) * n);
    alloc((void**)&qual, sizeof(double) * n);
    alloc((void**)&flt, sizeof(int32_t) * n);
    if (rc) set_error("run_sites_host: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    const bool want_post = e.tables && (post_host || call_host || qual_host || filter_host);
    const bool want_qual = e.tables && (qual_host || filter_host);

    auto new_event = [&](cudaEvent_t* ev) {
        if (cudaEventCreateWithFlags(ev, cudaEventDisableTiming) != cudaSuccess) { rc = 1; return; }
        events.push_back(*ev);
    };
    if (!rc) {
        cudaEvent_t ready;
        new_event(&ready);
        if (!rc) {
            cudaEventRecord(ready, s);                       // allocations are usable from the copy stream after this
            cudaStreamWaitEvent(cs, ready, 0);
        }
    }
    // The copy of chunk c+1 overlaps the kernels of chunk c, so only the FIRST chunk's copy is exposed: the chunk
    // sizes ramp up (1/16, 1/4 of an engine chunk, then full chunks; multiples of 128 candidates).
    const int64_t step = e.max_batch;
    const bool ramp = n > step;
    int chunk_idx = 0;
    Copied done_rows[2], done_bytes[2], done_ind[2];
    for (int64_t c0 = 0, nc = 0; c0 < n && !rc; c0 += nc, ++chunk_idx) {
        int64_t want = step;
        if (ramp && chunk_idx < 2) want = std::max<int64_t>(128, (step / (chunk_idx == 0 ? 16 : 4)) / 128 * 128);
        nc = std::min(want, n - c0);
        for (int k = 0; k < n_streams; ++k) {
            const cto_host_stream* x = hs[k];
            // rows touched by this chunk -> one contiguous span of plane bytes (rows are position sorted)
            int32_t r0 = INT32_MAX, r1 = -1;
            const int32_t* wp = x->win_pos + c0 * N_POS;
            for (int64_t i = 0; i < nc * N_POS; ++i) {
                const int32_t r = wp[i];
                if (r < 0) continue;
                r0 = r < r0 ? r : r0;
                r1 = r > r1 ? r : r1;
            }
            {   // the window table of the chunk goes even when every slot is an absent row
                const cudaError_t we = cudaMemcpyAsync(d[k].win_pos + c0 * N_POS, wp, sizeof(int32_t) * nc * N_POS, cudaMemcpyHostToDevice, cs);
                if (we != cudaSuccess) { set_error("run_sites_host: H2D copy failed: %s", cudaGetErrorString(we)); rc = 1; break; }
            }
            if (r1 < 0) continue;
            if (r1 >= x->n_rows) { set_error("run_sites_host: win_pos row %d outside the %lld pileup rows", r1, (long long)x->n_rows); rc = 2; break; }
            cudaError_t ce = cudaSuccess;
            int64_t ca, cb;
            // per-row arrays of the rows this chunk touches (offsets are absolute, so slices are enough); element r1 + 1 of
            // the offset arrays is re-sent with the next span (same value)
            done_rows[k].claim(r0, (int64_t)r1 + 1, ca, cb);
            if (cb > ca) {
                ce = cudaMemcpyAsync(d[k].grp_off + ca, x->grp_off + ca, sizeof(int32_t) * (cb - ca + 1), cudaMemcpyHostToDevice, cs);
                if (ce == cudaSuccess) ce = cudaMemcpyAsync(d[k].ind_off + ca, x->ind_off + ca, sizeof(int32_t) * (cb - ca + 1), cudaMemcpyHostToDevice, cs);
                if (ce == cudaSuccess) ce = cudaMemcpyAsync(d[k].ref_code + ca, x->ref_code + ca, cb - ca, cudaMemcpyHostToDevice, cs);
            }
            done_bytes[k].claim((int64_t)x->grp_off[r0] * 8, (int64_t)x->grp_off[r1 + 1] * 8, ca, cb);
            if (ce == cudaSuccess && cb > ca) ce = cudaMemcpyAsync(d[k].planes + ca, x->planes + ca, cb - ca, cudaMemcpyHostToDevice, cs);
            done_ind[k].claim(x->ind_off[r0], x->ind_off[r1 + 1], ca, cb);
            if (ce == cudaSuccess && cb > ca)
                ce = cudaMemcpyAsync(d[k].ind_entry + ca, x->ind_entry + ca, sizeof(uint32_t) * (cb - ca), cudaMemcpyHostToDevice, cs);
            if (ce != cudaSuccess) { set_error("run_sites_host: H2D copy failed: %s", cudaGetErrorString(ce)); rc = 1; break; }
        }
        if (rc) break;
        cudaEvent_t copied;
        new_event(&copied);
        if (rc) break;
        cudaEventRecord(copied, cs);
        cudaStreamWaitEvent(s, copied, 0);
        for (int k = 0; k < n_streams && !rc; ++k)
            rc = launch_encode_pileup(d[k].planes, d[k].grp_off, d[k].ref_code, d[k].ind_off, d[k].ind_entry, d[k].win_pos + c0 * N_POS,
                                      nc, hs[k]->n_groups, tens[k] + c0 * xin, dep[k] + c0, s);
        const int kn = neg ? 1 : 0;
        if (!rc)
            rc = cto_predict(h, tens[0] + c0 * xin, dep[0] + c0, tens[kn] + c0 * xin, dep[kn] + c0, nc, la + c0 * nh * 2,
                             ln + c0 * nh * 2, probs + c0 * nh * 4, want_post ? post + c0 * nh : nullptr,
                             want_post ? call + c0 : nullptr, nullptr, nullptr, want_qual ? qual + c0 : nullptr,
                             want_qual ? flt + c0 : nullptr, s);
    }
    if (!rc) {
        cudaError_t ce = cudaSuccess;
        if (probs_host) ce = cudaMemcpyAsync(probs_host, probs, sizeof(float) * n * nh * 4, cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess && want_post && post_host)
            ce = cudaMemcpyAsync(post_host, post, sizeof(double) * n * nh, cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess && want_post && call_host)
            ce = cudaMemcpyAsync(call_host, call, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess && want_qual && qual_host)
            ce = cudaMemcpyAsync(qual_host, qual, sizeof(double) * n, cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess && want_qual && filter_host)
            ce = cudaMemcpyAsync(filter_host, flt, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess && tensor_aff_host)
            ce = cudaMemcpyAsync(tensor_aff_host, tens[0], sizeof(int16_t) * n * xin, cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess && tensor_neg_host)
            ce = cudaMemcpyAsync(tensor_neg_host, tens[neg ? 1 : 0], sizeof(int16_t) * n * xin, cudaMemcpyDeviceToHost, s);
        if (ce != cudaSuccess) {
            set_error("run_sites_host: D2H copy failed: %s", cudaGetErrorString(ce));
            rc = 1;
        }
    }
    cudaStreamSynchronize(cs);                                // every span copy has landed (or failed) before frees
    for (int k = 0; k < 2; ++k) {
        free_stream(d[k], s);
        if (tens[k]) cudaFreeAsync(tens[k], s);
        if (dep[k]) cudaFreeAsync(dep[k], s);
    }
    void* ptrs[] = {la, ln, probs, post, call, qual, flt};
    for (void* p : ptrs)
        if (p) cudaFreeAsync(p, s);
    cudaError_t se = cudaStreamSynchronize(s);
    for (cudaEvent_t ev : events) cudaEventDestroy(ev);
    if (!rc && se != cudaSuccess) {
        set_error("run_sites_host: %s", cudaGetErrorString(se));
        rc = 1;
    }
    return rc;
}

}  // extern "C"
