// Inline PTX helpers shared by the tensor-core GRU kernels (gru_tc.cu, gru_tc2.cu).
#pragma once
#include "nn_kernels.cuh"
#include <cuda.h>

namespace cto {
namespace tc {

// helpers shared with gemm_tc.cu (same translation-unit-local definitions)
// one lane of a fully converged warp; everything around it stays warp-uniform so that descriptors live in
// uniform registers (issuing tcgen05 / TMA from inside an `if (lane == 0)` region makes ptxas wrap every
// instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop, ~70 cycles per MMA)
__device__ __forceinline__ bool g_elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "@px mov.s32 %0, 1;\n\t"
        "}" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t g_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void g_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(g_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(g_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void g_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(g_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void g_tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(g_smem_u32(dst)), "l"(map), "r"(g_smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void g_tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(g_smem_u32(dst)), "l"(map), "r"(g_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void g_prefetch_map(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void g_tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(g_smem_u32(dst)), "l"(map), "r"(g_smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void g_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(g_smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t g_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void g_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void g_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(g_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void g_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float g_round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ uint64_t g_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void g_tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

__device__ __forceinline__ void g_tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
// gate non-linearities through MUFU.EX2 / MUFU.RCP: absolute error ~1e-7, far inside the 1e-3 contract
__device__ __forceinline__ float g_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// raw MUFU ops: __fdividef wraps its reciprocal in range checks (FSETP + predicated scaling, ~5 instructions per call);
// the gate math keeps every operand of the reciprocal inside [1, 2^87], so the bare instruction is exact enough
__device__ __forceinline__ float g_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float g_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// GRU cell update for one element (torch.nn.GRU, gate order r|z|n): pre-activations ar = W_hr h (+ W_ir x + b_r), az likewise,
// an_h = W_hn h + b_hn, an_x = W_in x + b_in.  Five MUFU ops: z = 1 / (1 + ez) and n = 1 - 2 / (en + 1) share one
// reciprocal; the clamps keep the product of the two denominators far from fp32 overflow and change sigmoid / tanh by
// less than 1 ulp.  r needs no clamp: ex2 -> inf gives rcp -> 0, the correct limit.
__device__ __forceinline__ float g_gru_cell(float ar, float az, float an_h, float an_x, float h) {
    constexpr float L2E = 1.4426950408889634f;
    const float r = g_rcp(1.0f + g_ex2(-L2E * ar));
    const float xz = fminf(fmaxf(az, -30.0f), 30.0f);
    const float y = fminf(fmaxf(fmaf(r, an_h, an_x), -15.0f), 15.0f);
    const float dz = 1.0f + g_ex2(-L2E * xz), dn = 1.0f + g_ex2(2.0f * L2E * y);
    const float inv = g_rcp(dz * dn);
    const float z = dn * inv;
    const float n = fmaf(-2.0f * dz, inv, 1.0f);
    return fmaf(z, h - n, n);                       // (1 - z) * n + z * h
}
__device__ __forceinline__ float g_tanh(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }


int make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows);
int make_map_bf16(CUtensorMap* map, const uint16_t* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows);
int make_map_plain(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols);
int make_map_rows_nd(CUtensorMap* map, const float* ptr, int rank, int64_t cols, int64_t ld, const int64_t* dim_size,
                     const int64_t* row_stride, const int* box);

}  // namespace tc
}  // namespace cto
