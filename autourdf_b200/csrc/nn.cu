// nn.cu -- batched brute-force nearest neighbour under squared L2 (sm_100a).
//
// The correspondence search of open3d GetRegistrationResultAndCorrespondences (behind the
// registration_icp call at AutoURDF PointCloud/cluster_icp.py:157) exposed on its own:
// for every query point of group g, argmin_j over the group's targets of
// ((dx*dx)+dy*dy)+dz*dz in float64, each operation rounded (nanoflann L2_Simple_Adaptor
// order), lowest index on exact ties.  Targets are staged through shared memory as
// float64 SoA chunks; all lanes of a warp read the same target (broadcast, conflict free).
#include <math.h>

#include "common.cuh"

namespace aurdf {

constexpr int kNnThreads = 128;
constexpr int kNnChunk = 1024;

__global__ void __launch_bounds__(kNnThreads)
nn_l2_kernel(const void *__restrict__ query, const int *__restrict__ qoff, const void *__restrict__ target,
             const int *__restrict__ toff, int dtype, int *__restrict__ out_idx, double *__restrict__ out_d2) {
    __shared__ double sx[kNnChunk], sy[kNnChunk], sz[kNnChunk];
    const int g = blockIdx.x, tid = threadIdx.x;
    const int q0 = qoff[g], nq = qoff[g + 1] - q0;
    const int t0 = toff[g], nt = toff[g + 1] - t0;
    for (int qb = blockIdx.y * kNnThreads; qb < nq; qb += gridDim.y * kNnThreads) {
        const int i = qb + tid;
        const bool active = i < nq;
        double x = 0, y = 0, z = 0;
        if (active) {
            const size_t e = 3 * (size_t)(q0 + i);
            x = ld_coord(query, dtype, e); y = ld_coord(query, dtype, e + 1); z = ld_coord(query, dtype, e + 2);
        }
        double bd = INFINITY;
        int bj = -1;
        for (int c0 = 0; c0 < nt; c0 += kNnChunk) {
            const int n = min(kNnChunk, nt - c0);
            __syncthreads();
            for (int j = tid; j < n; j += kNnThreads) {
                const size_t e = 3 * (size_t)(t0 + c0 + j);
                sx[j] = ld_coord(target, dtype, e); sy[j] = ld_coord(target, dtype, e + 1); sz[j] = ld_coord(target, dtype, e + 2);
            }
            __syncthreads();
            if (active) {
#pragma unroll 4
                for (int j = 0; j < n; ++j) {
                    const double dx = __dsub_rn(x, sx[j]), dy = __dsub_rn(y, sy[j]), dz = __dsub_rn(z, sz[j]);
                    const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    if (d < bd) { bd = d; bj = c0 + j; }
                }
            }
        }
        if (active) {
            out_idx[q0 + i] = bj;
            if (out_d2) out_d2[q0 + i] = bd;
        }
    }
}

}  // namespace aurdf

using namespace aurdf;

extern "C" int aurdf_nn_l2(const void *query_xyz, const int32_t *query_off, const void *target_xyz,
                           const int32_t *target_off, int pts_dtype, int32_t n_groups, int64_t n_queries,
                           int32_t *out_idx, double *out_d2, aurdf_stream_t stream_) {
    aurdf::NvtxRange nvtx_range("aurdf_nn_l2");
    cudaStream_t stream = (cudaStream_t)stream_;
    AURDF_REQUIRE(n_groups >= 0 && n_queries >= 0, "aurdf_nn_l2: negative size");
    if (n_groups == 0 || n_queries == 0) return AURDF_OK;
    AURDF_REQUIRE(query_xyz && query_off && target_off && out_idx, "aurdf_nn_l2: NULL pointer");
    AURDF_REQUIRE(pts_dtype == AURDF_F32 || pts_dtype == AURDF_F64, "aurdf_nn_l2: bad dtype");
    int64_t avg = (n_queries + n_groups - 1) / n_groups;
    int64_t gy = (avg + kNnThreads - 1) / kNnThreads;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    dim3 grid(n_groups, (unsigned)gy);
    nn_l2_kernel<<<grid, kNnThreads, 0, stream>>>(query_xyz, query_off, target_xyz, target_off, pts_dtype, out_idx, out_d2);
    AURDF_CUDA_CHECK(cudaGetLastError());
    return AURDF_OK;
}
