// Shared helpers for the clairs-to_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

namespace cto {

constexpr int N_POS = 33;     // shared/param.py:60  (2 * flankingBaseNum + 1)
constexpr int N_CH = 34;      // shared/param.py:56  (len(pileup_channel))
constexpr int CENTER = 16;    // shared/param.py:59
constexpr int MIN_MQ = 20;    // literal in src/create_tensor_pileup_calling.py:147-148
constexpr int QUAL_ABSENT = 254;
constexpr int MIN_RESCALE_COV = 50;  // shared/param.py:26

// thread-local error string behind cto_last_error()
void set_error(const char* fmt, ...);
const char* get_error();

#define CTO_CHECK(expr)                                                                       \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            cto::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,               \
                           cudaGetErrorString(_e));                                           \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)

#define CTO_REQUIRE(cond, ...)                                                                \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            cto::set_error(__VA_ARGS__);                                                      \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// SM count of the CURRENT device (cached per device; several engines on several GPUs may live in one process)
int device_sm_count();
// opt-in dynamic shared memory above 48 KB.  The attribute belongs to the (function, device) pair, so it is set on
// every call instead of behind a process-wide flag (a second engine on another GPU failed to launch)
template <typename F>
static inline cudaError_t set_max_dynamic_smem(F* kernel, int bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

// stream-ordered scratch allocation from a private per-device pool that keeps its memory (free with cudaFreeAsync)
cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t s);

// number of kernels this library has launched in this process (bench.py reports it as gpu_launches)
void count_launch(int n = 1);
int64_t launches();

}  // namespace cto
