// common.cuh -- shared helpers of libaurdf (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges show up in nsys / ncu timelines, cost nothing without a tool

#include "aurdf.h"

namespace aurdf {

// thread-local last-error string behind aurdf_last_error_string()
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define AURDF_CUDA_CHECK(expr)                                      \
    do {                                                            \
        cudaError_t e__ = (expr);                                   \
        if (e__ != cudaSuccess) return ::aurdf::cuda_fail(e__, #expr); \
    } while (0)

#define AURDF_REQUIRE(cond, msg)                \
    do {                                        \
        if (!(cond)) {                          \
            ::aurdf::set_error("%s", msg);      \
            return AURDF_EINVAL;                \
        }                                       \
    } while (0)

constexpr int kNumSMs = 148;  // B200

// NVTX range over a C-ABI entry point (tracing)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// current device and its SM count (cached per device; idempotent, so the unsynchronised fill is benign)
int current_device_sms(int *device);
// run `set` once per (slot, device): kernel attributes that never change (idempotent as well)
bool once_per_device(int slot, int device);

// ---- storage-dtype-agnostic coordinate load (exact widening of float32) -------------------
__device__ __forceinline__ double ld_coord(const void *base, int dtype, size_t idx) {
    return dtype == AURDF_F32 ? (double)__ldg((const float *)base + idx) : __ldg((const double *)base + idx);
}

// ---- warp reductions ----------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine; SASS: UBLKCP / SYNCS) ---------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// general 4x4 inverse, Gauss-Jordan with partial pivoting (np.linalg.inv stand-in)
__device__ inline bool inv4(const double *A, double *Ai) {
    double M[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            M[i][j] = A[4 * i + j];
            M[i][j + 4] = (i == j) ? 1.0 : 0.0;
        }
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        int piv = c;
#pragma unroll
        for (int r = c + 1; r < 4; ++r)
            if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
        if (M[piv][c] == 0.0) ok = false;
        if (piv != c) {
            for (int j = 0; j < 8; ++j) {
                const double t = M[c][j];
                M[c][j] = M[piv][j];
                M[piv][j] = t;
            }
        }
        const double d = M[c][c];
#pragma unroll
        for (int j = 0; j < 8; ++j) M[c][j] /= d;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            const double f = M[r][c];
#pragma unroll
            for (int j = 0; j < 8; ++j) M[r][j] -= f * M[c][j];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) Ai[4 * i + j] = M[i][j + 4];
    return ok;
}


}  // namespace aurdf
