// se3.cu -- batched SE(3) apply kernels (sm_100a).
//
//   aurdf_se3_apply      calculate_pc(), AutoURDF PointCloud/mlp_reg.py:155-170:
//                        pc_k = X_k @ T_k[:3,:3].T + T_k[:3,3]
//   aurdf_se3_apply_bwd  its adjoint, so train() (mlp_reg.py:93) can keep autograd
//   aurdf_se3_to_local   inv(T_k) @ [X;1], mlp_reg.py:211-213 / cluster_icp.py:96-98
//
// All three are streaming, HBM-bound ops: 24 B/point (f32) or 48 B/point (f64) of compulsory
// traffic, one pose (64/128 B) per group.  One CTA walks one group so the pose stays in
// registers; groups are ragged (CSR offsets).
#include "common.cuh"

namespace aurdf {

constexpr int kSe3Threads = 128;

template <typename T>
__global__ void __launch_bounds__(kSe3Threads)
se3_apply_kernel(const T *__restrict__ xyz, const int *__restrict__ off, const T *__restrict__ Tm,
                 T *__restrict__ out) {
    const int k = blockIdx.x;
    const int p0 = off[k], n = off[k + 1] - p0;
    const T *M = Tm + 16 * (size_t)k;
    const T r00 = M[0], r01 = M[1], r02 = M[2], t0 = M[3];
    const T r10 = M[4], r11 = M[5], r12 = M[6], t1 = M[7];
    const T r20 = M[8], r21 = M[9], r22 = M[10], t2 = M[11];
    for (int i = blockIdx.y * kSe3Threads + threadIdx.x; i < n; i += gridDim.y * kSe3Threads) {
        const size_t e = 3 * (size_t)(p0 + i);
        const T x = xyz[e], y = xyz[e + 1], z = xyz[e + 2];
        out[e] = (x * r00 + y * r01 + z * r02) + t0;
        out[e + 1] = (x * r10 + y * r11 + z * r12) + t1;
        out[e + 2] = (x * r20 + y * r21 + z * r22) + t2;
    }
}

// grad_X = g R ; grad_T[:3,:3] = sum_i g_i^T x_i ; grad_T[:3,3] = sum_i g_i ; row 3 = 0
template <typename T>
__global__ void __launch_bounds__(kSe3Threads)
se3_apply_bwd_kernel(const T *__restrict__ g, const T *__restrict__ xyz, const int *__restrict__ off,
                     const T *__restrict__ Tm, T *__restrict__ gx, T *__restrict__ gT) {
    __shared__ double s_red[kSe3Threads / 32][12];
    const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p0 = off[k], n = off[k + 1] - p0;
    const T *M = Tm + 16 * (size_t)k;
    const T r00 = M[0], r01 = M[1], r02 = M[2];
    const T r10 = M[4], r11 = M[5], r12 = M[6];
    const T r20 = M[8], r21 = M[9], r22 = M[10];
    double acc[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = 0.0;
    for (int i = tid; i < n; i += kSe3Threads) {
        const size_t e = 3 * (size_t)(p0 + i);
        const T g0 = g[e], g1 = g[e + 1], g2 = g[e + 2];
        const T x = xyz[e], y = xyz[e + 1], z = xyz[e + 2];
        if (gx) {
            gx[e] = g0 * r00 + g1 * r10 + g2 * r20;
            gx[e + 1] = g0 * r01 + g1 * r11 + g2 * r21;
            gx[e + 2] = g0 * r02 + g1 * r12 + g2 * r22;
        }
        acc[0] += (double)g0 * x; acc[1] += (double)g0 * y; acc[2] += (double)g0 * z; acc[3] += g0;
        acc[4] += (double)g1 * x; acc[5] += (double)g1 * y; acc[6] += (double)g1 * z; acc[7] += g1;
        acc[8] += (double)g2 * x; acc[9] += (double)g2 * y; acc[10] += (double)g2 * z; acc[11] += g2;
    }
    if (!gT) return;
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = warp_sum(acc[j]);
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 12; ++j) s_red[warp][j] = acc[j];
    }
    __syncthreads();
    if (tid < 16) {
        double v = 0.0;
        if (tid < 12) {
#pragma unroll
            for (int w = 0; w < kSe3Threads / 32; ++w) v += s_red[w][tid];
        }
        gT[16 * (size_t)k + tid] = (T)v;
    }
}

__global__ void __launch_bounds__(kSe3Threads)
se3_to_local_kernel(const double *__restrict__ xyz, const int *__restrict__ off, const double *__restrict__ Tm,
                    double *__restrict__ out) {
    __shared__ double s_inv[16];
    const int k = blockIdx.x;
    if (threadIdx.x == 0) {
        double A[16], Ai[16];
        for (int j = 0; j < 16; ++j) A[j] = Tm[16 * (size_t)k + j];
        const bool ok = inv4(A, Ai);
        for (int j = 0; j < 16; ++j) s_inv[j] = ok ? Ai[j] : nan("");
    }
    __syncthreads();
    const int p0 = off[k], n = off[k + 1] - p0;
    for (int i = blockIdx.y * kSe3Threads + threadIdx.x; i < n; i += gridDim.y * kSe3Threads) {
        const size_t e = 3 * (size_t)(p0 + i);
        const double x = xyz[e], y = xyz[e + 1], z = xyz[e + 2];
        out[e] = s_inv[0] * x + s_inv[1] * y + s_inv[2] * z + s_inv[3];
        out[e + 1] = s_inv[4] * x + s_inv[5] * y + s_inv[6] * z + s_inv[7];
        out[e + 2] = s_inv[8] * x + s_inv[9] * y + s_inv[10] * z + s_inv[11];
    }
}

// CTAs per group along y: enough to cover the average group, capped so the grid stays modest
static int ctas_per_group(int n_groups, int64_t n_points) {
    if (n_groups <= 0) return 1;
    int64_t avg = (n_points + n_groups - 1) / n_groups;
    int64_t y = (avg + 4 * kSe3Threads - 1) / (4 * kSe3Threads);
    if (y < 1) y = 1;
    if (y > 64) y = 64;
    return (int)y;
}

}  // namespace aurdf

using namespace aurdf;

extern "C" int aurdf_se3_apply(const void *xyz, const int32_t *off, const void *T, int32_t n_groups,
                               int64_t n_points, int dtype, void *out_xyz, aurdf_stream_t stream_) {
    aurdf::NvtxRange nvtx_range("aurdf_se3_apply");
    cudaStream_t stream = (cudaStream_t)stream_;
    AURDF_REQUIRE(n_groups >= 0 && n_points >= 0, "aurdf_se3_apply: negative size");
    if (n_groups == 0 || n_points == 0) return AURDF_OK;
    AURDF_REQUIRE(xyz && off && T && out_xyz, "aurdf_se3_apply: NULL pointer");
    AURDF_REQUIRE(dtype == AURDF_F32 || dtype == AURDF_F64, "aurdf_se3_apply: bad dtype");
    dim3 grid(n_groups, ctas_per_group(n_groups, n_points));
    if (dtype == AURDF_F32)
        se3_apply_kernel<float><<<grid, kSe3Threads, 0, stream>>>((const float *)xyz, off, (const float *)T, (float *)out_xyz);
    else
        se3_apply_kernel<double><<<grid, kSe3Threads, 0, stream>>>((const double *)xyz, off, (const double *)T, (double *)out_xyz);
    AURDF_CUDA_CHECK(cudaGetLastError());
    return AURDF_OK;
}

extern "C" int aurdf_se3_apply_bwd(const void *grad_out, const void *xyz, const int32_t *off, const void *T,
                                   int32_t n_groups, int64_t n_points, int dtype, void *grad_xyz, void *grad_T,
                                   aurdf_stream_t stream_) {
    aurdf::NvtxRange nvtx_range("aurdf_se3_apply_bwd");
    cudaStream_t stream = (cudaStream_t)stream_;
    AURDF_REQUIRE(n_groups >= 0 && n_points >= 0, "aurdf_se3_apply_bwd: negative size");
    if (n_groups == 0) return AURDF_OK;
    AURDF_REQUIRE(grad_out && xyz && off && T, "aurdf_se3_apply_bwd: NULL pointer");
    AURDF_REQUIRE(dtype == AURDF_F32 || dtype == AURDF_F64, "aurdf_se3_apply_bwd: bad dtype");
    if (dtype == AURDF_F32)
        se3_apply_bwd_kernel<float><<<n_groups, kSe3Threads, 0, stream>>>((const float *)grad_out, (const float *)xyz, off,
                                                                         (const float *)T, (float *)grad_xyz, (float *)grad_T);
    else
        se3_apply_bwd_kernel<double><<<n_groups, kSe3Threads, 0, stream>>>((const double *)grad_out, (const double *)xyz, off,
                                                                          (const double *)T, (double *)grad_xyz, (double *)grad_T);
    AURDF_CUDA_CHECK(cudaGetLastError());
    return AURDF_OK;
}

extern "C" int aurdf_se3_to_local(const double *xyz, const int32_t *off, const double *T, int32_t n_groups,
                                  int64_t n_points, double *out_xyz, aurdf_stream_t stream_) {
    aurdf::NvtxRange nvtx_range("aurdf_se3_to_local");
    cudaStream_t stream = (cudaStream_t)stream_;
    AURDF_REQUIRE(n_groups >= 0 && n_points >= 0, "aurdf_se3_to_local: negative size");
    if (n_groups == 0 || n_points == 0) return AURDF_OK;
    AURDF_REQUIRE(xyz && off && T && out_xyz, "aurdf_se3_to_local: NULL pointer");
    dim3 grid(n_groups, ctas_per_group(n_groups, n_points));
    se3_to_local_kernel<<<grid, kSe3Threads, 0, stream>>>(xyz, off, T, out_xyz);
    AURDF_CUDA_CHECK(cudaGetLastError());
    return AURDF_OK;
}
