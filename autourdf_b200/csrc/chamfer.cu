// chamfer.cu -- float32 brute-force 1-NN under L1 or squared-L2 + the chamfer backward (sm_100a).
//
// Replaces the hot op of the reference's default registration path: pytorch3d 0.7.7
// chamfer_distance(pred, y, norm=1) -> ops.knn_points(K=1) in both directions (AutoURDF
// PointCloud/mlp_reg.py:96, fwd + bwd ~600 times per frame; Sim/evaluation.py:81).
//   distance   ((0 + |dx|) + |dy|) + |dz|   (norm 1)   or   ((0 + dx*dx) + dy*dy) + dz*dz   (norm 2),
//              float32, pytorch3d's accumulation order, no FMA contraction (file built -fmad=false)
//   argmin     first minimum wins (strict '<' in pytorch3d's scan) = lowest index
//   backward   grad_p1 += g * sign(p1 - p2_nn)      (norm 1)   |  g * 2 (p1 - p2_nn)   (norm 2)
//              grad_p2[nn] -= the same (atomicAdd scatter, as pytorch3d does)
// Work decomposition: clouds of ~5000 points are far too small for a one-CTA-per-query-block grid,
// so the target range is split as well (grid.z) and partial minima are merged with one 64-bit
// atomicMin per query on the packed key (float bits of the distance << 32 | index): distances are
// non-negative, so unsigned order on the bits is numeric order and ties resolve to the lowest index.
#include "common.cuh"

namespace aurdf {

constexpr int kChThreads = 256;
constexpr int kChQPT = 2;        // queries per thread
constexpr int kChChunk = 1024;   // targets staged per shared-memory chunk (float4, 16 KB)

template <int NORM>
__device__ __forceinline__ float pdist(float qx, float qy, float qz, const float4 &t) {
    const float dx = qx - t.x, dy = qy - t.y, dz = qz - t.z;
    if (NORM == 1) return (fabsf(dx) + fabsf(dy)) + fabsf(dz);
    return (dx * dx + dy * dy) + dz * dz;
}

template <int NORM>
__global__ void __launch_bounds__(kChThreads)
nn_f32_kernel(const float *__restrict__ query, const int *__restrict__ qoff, const float *__restrict__ target,
              const int *__restrict__ toff, unsigned long long *__restrict__ keys) {
    __shared__ float4 st[kChChunk];
    const int g = blockIdx.x, tid = threadIdx.x;
    const int q0 = qoff[g], nq = qoff[g + 1] - q0;
    const int t0 = toff[g], nt = toff[g + 1] - t0;
    // this CTA's slice of the targets (grid.z slices, chunk-aligned)
    const int chunks = (nt + kChChunk - 1) / kChChunk;
    const int per = (chunks + gridDim.z - 1) / gridDim.z;
    const int c_lo = blockIdx.z * per, c_hi = min(chunks, c_lo + per);
    if (c_lo >= c_hi) return;
    for (int qb = blockIdx.y * kChThreads * kChQPT; qb < nq; qb += gridDim.y * kChThreads * kChQPT) {
        float qx[kChQPT], qy[kChQPT], qz[kChQPT], bd[kChQPT];
        int bj[kChQPT];
#pragma unroll
        for (int k = 0; k < kChQPT; ++k) {
            const int i = qb + tid + k * kChThreads;
            const size_t e = 3 * (size_t)(q0 + min(i, nq - 1));
            qx[k] = query[e]; qy[k] = query[e + 1]; qz[k] = query[e + 2];
            bd[k] = INFINITY; bj[k] = 0x7fffffff;
        }
        for (int c = c_lo; c < c_hi; ++c) {
            const int base = c * kChChunk, n = min(kChChunk, nt - base);
            __syncthreads();
            for (int j = tid; j < n; j += kChThreads) {
                const size_t e = 3 * (size_t)(t0 + base + j);
                st[j] = make_float4(target[e], target[e + 1], target[e + 2], 0.f);
            }
            __syncthreads();
#pragma unroll 4
            for (int j = 0; j < n; ++j) {
                const float4 t = st[j];
#pragma unroll
                for (int k = 0; k < kChQPT; ++k) {
                    const float d = pdist<NORM>(qx[k], qy[k], qz[k], t);
                    if (d < bd[k]) { bd[k] = d; bj[k] = base + j; }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kChQPT; ++k) {
            const int i = qb + tid + k * kChThreads;
            if (i < nq && bj[k] != 0x7fffffff) {
                const unsigned long long key = ((unsigned long long)__float_as_uint(bd[k]) << 32) | (unsigned)bj[k];
                atomicMin(keys + q0 + i, key);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
nn_unpack_kernel(const unsigned long long *__restrict__ keys, long long n, int *__restrict__ idx, float *__restrict__ dist) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    const bool none = k == 0xffffffffffffffffULL;   // empty target group
    idx[i] = none ? -1 : (int)(k & 0xffffffffu);
    if (dist) dist[i] = none ? INFINITY : __uint_as_float((unsigned)(k >> 32));
}

// grad_p1[i] = g_i * d dist / d p1 ; grad_p2[nn(i)] -= the same.  g_i = gdist[i] if given, else gscale[group]
template <int NORM>
__global__ void __launch_bounds__(256)
nn_bwd_kernel(const float *__restrict__ p1, const int *__restrict__ off1, const float *__restrict__ p2,
              const int *__restrict__ off2, const int *__restrict__ idx, const float *__restrict__ gdist, int n_groups,
              long long n1, float *__restrict__ grad1, float *__restrict__ grad2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    // group of point i (few groups: linear scan of the offsets)
    int g = 0;
    while (g + 1 < n_groups && i >= off1[g + 1]) ++g;
    const int j = idx[i];
    if (j < 0) return;
    const float w = gdist[i];
    const size_t e1 = 3 * (size_t)i, e2 = 3 * (size_t)(off2[g] + j);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float diff = p1[e1 + d] - p2[e2 + d];
        float gr;
        if (NORM == 1) gr = w * (float)((diff > 0.f) - (diff < 0.f));
        else gr = w * 2.f * diff;
        if (grad1) atomicAdd(grad1 + e1 + d, gr);
        if (grad2) atomicAdd(grad2 + e2 + d, -gr);
    }
}

}  // namespace aurdf

using namespace aurdf;

extern "C" size_t aurdf_nn_f32_workspace_bytes(int64_t n_queries) {
    return n_queries < 0 ? 0 : (size_t)n_queries * sizeof(unsigned long long);
}

extern "C" int aurdf_nn_f32(const float *query_xyz, const int32_t *query_off, const float *target_xyz,
                            const int32_t *target_off, int32_t n_groups, int64_t n_queries, int64_t n_targets, int norm,
                            int32_t *out_idx, float *out_dist, void *workspace, size_t workspace_bytes,
                            aurdf_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AURDF_REQUIRE(n_groups >= 0 && n_queries >= 0 && n_targets >= 0, "aurdf_nn_f32: negative size");
    AURDF_REQUIRE(norm == 1 || norm == 2, "aurdf_nn_f32: norm must be 1 or 2");
    if (n_groups == 0 || n_queries == 0) return AURDF_OK;
    AURDF_REQUIRE(query_xyz && query_off && target_off && out_idx && workspace, "aurdf_nn_f32: NULL pointer");
    if (workspace_bytes < (size_t)n_queries * sizeof(unsigned long long)) {
        set_error("aurdf_nn_f32: workspace_bytes %zu < %zu", workspace_bytes, (size_t)n_queries * 8);
        return AURDF_EWORKSPACE;
    }
    unsigned long long *keys = (unsigned long long *)workspace;
    AURDF_CUDA_CHECK(cudaMemsetAsync(keys, 0xff, (size_t)n_queries * sizeof(unsigned long long), stream));
    if (n_targets > 0) {
        const int64_t avg_q = (n_queries + n_groups - 1) / n_groups, avg_t = (n_targets + n_groups - 1) / n_groups;
        int64_t gy = (avg_q + kChThreads * kChQPT - 1) / (kChThreads * kChQPT);
        if (gy < 1) gy = 1;
        if (gy > 65535) gy = 65535;
        // split the targets until the grid covers ~2 waves of the 148 SMs (never finer than one chunk)
        int64_t chunks = (avg_t + kChChunk - 1) / kChChunk;
        int64_t gz = (2 * kNumSMs + n_groups * gy - 1) / (n_groups * gy);
        if (gz > chunks) gz = chunks;
        if (gz < 1) gz = 1;
        if (gz > 64) gz = 64;
        dim3 grid(n_groups, (unsigned)gy, (unsigned)gz);
        if (norm == 1) nn_f32_kernel<1><<<grid, kChThreads, 0, stream>>>(query_xyz, query_off, target_xyz, target_off, keys);
        else nn_f32_kernel<2><<<grid, kChThreads, 0, stream>>>(query_xyz, query_off, target_xyz, target_off, keys);
    }
    nn_unpack_kernel<<<(unsigned)((n_queries + 255) / 256), 256, 0, stream>>>(keys, n_queries, out_idx, out_dist);
    AURDF_CUDA_CHECK(cudaGetLastError());
    return AURDF_OK;
}

extern "C" int aurdf_nn_f32_bwd(const float *p1_xyz, const int32_t *p1_off, const float *p2_xyz, const int32_t *p2_off,
                                const int32_t *idx, const float *grad_dist, int32_t n_groups, int64_t n_p1, int norm,
                                float *grad_p1, float *grad_p2, aurdf_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AURDF_REQUIRE(n_groups >= 0 && n_p1 >= 0, "aurdf_nn_f32_bwd: negative size");
    AURDF_REQUIRE(norm == 1 || norm == 2, "aurdf_nn_f32_bwd: norm must be 1 or 2");
    if (n_groups == 0 || n_p1 == 0) return AURDF_OK;
    AURDF_REQUIRE(p1_xyz && p1_off && p2_xyz && p2_off && idx && grad_dist, "aurdf_nn_f32_bwd: NULL pointer");
    const unsigned grid = (unsigned)((n_p1 + 255) / 256);
    if (norm == 1) nn_bwd_kernel<1><<<grid, 256, 0, stream>>>(p1_xyz, p1_off, p2_xyz, p2_off, idx, grad_dist, n_groups, n_p1, grad_p1, grad_p2);
    else nn_bwd_kernel<2><<<grid, 256, 0, stream>>>(p1_xyz, p1_off, p2_xyz, p2_off, idx, grad_dist, n_groups, n_p1, grad_p1, grad_p2);
    AURDF_CUDA_CHECK(cudaGetLastError());
    return AURDF_OK;
}
