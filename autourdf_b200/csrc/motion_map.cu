// motion_map.cu -- pairwise cluster motion-distance map.
//
// Replaces CoordMap.coord_dist_map (AutoURDF PointCloud/coord_map.py:230-307), which walks
// T x K x K Python loops with one roma / torch call per element.  Input is what the registration
// loop wrote to matrix/{t:04}.npy: T frames of K cluster poses (row-major 4x4 doubles).
//
//   diff mode (:253-286), per frame step i (T-1 of them):
//     trans_diff_j = t_j(i+1) - t_j(i)
//     rotvec_j     = roma.rotmat_to_rotvec(R_j(i)^T R_j(i+1))
//     d_xyz[j,k]   = |trans_diff_j - trans_diff_k| / (2 bounding_box)
//     d_rpy[j,k]   = roma.utils.rotvec_geodesic_distance(rotvec_j, rotvec_k) / pi
//     map[j,k,i]   = |d_xyz[j,:] - d_xyz[k,:]|_2 + |d_rpy[j,:] - d_rpy[k,:]|_2
//   pose mode (:287-302), per frame i:
//     map[j,k,i]   = |t_j - t_k| / (2 bounding_box) + roma.rotmat_geodesic_distance(R_j, R_k) / pi
//   sum_map[j,k] = sum_i |map[j,k,i]|                                             (:304-305)
//
// roma is not vendored in the reference; its algorithms (scipy's Rotation algorithms) are restated
// here exactly as in oracle/coord_map_oracle.py.  Everything is float64.
//
// Kernels: pair_kernel (one CTA per step: per-cluster motion in shared memory, then the K x K
// first-level distances), rows_kernel (diff mode: K x K row-difference norms; d is symmetric, so the
// column read d[m][k] is coalesced), sum_kernel.  HBM/L2 streaming, 16 (T-1) K^2 bytes of scratch.
#include <math.h>

#include "common.cuh"

namespace aurdf {

constexpr int kMapThreads = 256;
constexpr int kMapMaxK = 384;   // 16 doubles of shared memory per cluster

// roma.rotmat_to_unitquat (scipy from_matrix -> as_quat), xyzw, normalised
__device__ void rotmat_to_unitquat(const double *R, double *q) {
    const double d0 = R[0], d1 = R[4], d2 = R[8], tr = (d0 + d1) + d2;
    int c = 0;                      // argmax over (d0, d1, d2, tr), first maximum wins
    double best = d0;
    if (d1 > best) { best = d1; c = 1; }
    if (d2 > best) { best = d2; c = 2; }
    if (tr > best) { c = 3; }
    if (c != 3) {
        const int i = c, j = (i + 1) % 3, k = (j + 1) % 3;
        q[i] = 1.0 - tr + 2.0 * R[4 * i];
        q[j] = R[3 * j + i] + R[3 * i + j];
        q[k] = R[3 * k + i] + R[3 * i + k];
        q[3] = R[3 * k + j] - R[3 * j + k];
    } else {
        q[0] = R[7] - R[5];
        q[1] = R[2] - R[6];
        q[2] = R[3] - R[1];
        q[3] = 1.0 + tr;
    }
    const double n = sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

// roma.unitquat_to_rotvec (shortest arc) followed by roma.rotvec_to_unitquat: what
// rotvec_geodesic_distance sees of a rotvec produced by rotmat_to_rotvec
__device__ void quat_via_rotvec(double *q) {
    if (q[3] < 0.0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    const double half = atan2(sqrt((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]), q[3]);
    const double ang = 2.0 * half;
    const double s = fabs(ang) <= 1e-3 ? 2.0 + ang * ang / 12.0 + 7.0 * (ang * ang) * (ang * ang) / 2880.0 : ang / sin(half);
    const double vx = s * q[0], vy = s * q[1], vz = s * q[2];
    const double a = sqrt((vx * vx + vy * vy) + vz * vz);
    const double s2 = a <= 1e-3 ? 0.5 - a * a / 48.0 + (a * a) * (a * a) / 3840.0 : sin(a / 2.0) / a;
    q[0] = s2 * vx; q[1] = s2 * vy; q[2] = s2 * vz; q[3] = cos(a / 2.0);
}

__global__ void __launch_bounds__(kMapThreads)
pair_kernel(const double *__restrict__ M, int T, int K, double lam_bbox, double lam_rot, int diff,
            double *__restrict__ d_xyz, double *__restrict__ d_rpy, double *__restrict__ out_map, int steps) {
    extern __shared__ double sm[];
    double *s_t = sm;            // K x 3: translation (difference)
    double *s_q = sm + 3 * K;    // K x 4: unit quaternion (diff mode)
    double *s_R = sm + 7 * K;    // K x 9: rotation (pose mode)
    const int i = blockIdx.x, tid = threadIdx.x;
    for (int j = tid; j < K; j += kMapThreads) {
        const double *A = M + ((size_t)i * K + j) * 16;
        if (diff) {
            const double *B = A + (size_t)K * 16;   // same cluster, next frame
            s_t[3 * j] = B[3] - A[3]; s_t[3 * j + 1] = B[7] - A[7]; s_t[3 * j + 2] = B[11] - A[11];
            double rel[9], q[4];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c)   // (R_i^T R_{i+1})[r][c]
                    rel[3 * r + c] = (A[r] * B[c] + A[4 + r] * B[4 + c]) + A[8 + r] * B[8 + c];
            rotmat_to_unitquat(rel, q);
            quat_via_rotvec(q);
            s_q[4 * j] = q[0]; s_q[4 * j + 1] = q[1]; s_q[4 * j + 2] = q[2]; s_q[4 * j + 3] = q[3];
        } else {
            s_t[3 * j] = A[3]; s_t[3 * j + 1] = A[7]; s_t[3 * j + 2] = A[11];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) s_R[9 * j + 3 * r + c] = A[4 * r + c];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < K * K; idx += kMapThreads) {
        const int j = idx / K, k = idx - j * K;
        const double dx = s_t[3 * j] - s_t[3 * k], dy = s_t[3 * j + 1] - s_t[3 * k + 1], dz = s_t[3 * j + 2] - s_t[3 * k + 2];
        const double dt = lam_bbox * sqrt((dx * dx + dy * dy) + dz * dz);
        if (diff) {
            double nm = 0.0, np_ = 0.0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double a = s_q[4 * k + c] - s_q[4 * j + c], b = s_q[4 * k + c] + s_q[4 * j + c];
                nm += a * a;
                np_ += b * b;
            }
            const double dr = lam_rot * (4.0 * asin(0.5 * fmin(sqrt(nm), sqrt(np_))));
            d_xyz[(size_t)i * K * K + idx] = dt;
            d_rpy[(size_t)i * K * K + idx] = dr;
        } else {
            double f = 0.0;
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                const double a = s_R[9 * k + c] - s_R[9 * j + c];
                f += a * a;
            }
            const double dr = lam_rot * (2.0 * asin(fmin(sqrt(f) / (2.0 * 1.4142135623730951), 1.0)));
            out_map[(size_t)idx * steps + i] = dt + dr;
        }
    }
}

// diff mode, second level: map[j,k,i] = |d_xyz[j,:] - d_xyz[k,:]| + |d_rpy[j,:] - d_rpy[k,:]|
__global__ void __launch_bounds__(kMapThreads)
rows_kernel(const double *__restrict__ d_xyz, const double *__restrict__ d_rpy, int K, int steps,
            double *__restrict__ out_map) {
    const int i = blockIdx.x;
    const double *X = d_xyz + (size_t)i * K * K, *Rr = d_rpy + (size_t)i * K * K;
    for (int idx = threadIdx.x; idx < K * K; idx += kMapThreads) {
        const int j = idx / K, k = idx - j * K;
        double sx = 0.0, sr = 0.0;
        for (int m = 0; m < K; ++m) {
            // d is symmetric bit for bit, so row k is read as column k (coalesced across the warp)
            const double a = X[(size_t)j * K + m] - X[(size_t)m * K + k];
            const double b = Rr[(size_t)j * K + m] - Rr[(size_t)m * K + k];
            sx += a * a;
            sr += b * b;
        }
        out_map[(size_t)idx * steps + i] = sqrt(sx) + sqrt(sr);
    }
}

__global__ void sum_kernel(const double *__restrict__ map, int KK, int steps, double *__restrict__ out_sum) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= KK) return;
    double s = 0.0;
    for (int i = 0; i < steps; ++i) s += fabs(map[(size_t)idx * steps + i]);
    out_sum[idx] = s;
}

}  // namespace aurdf

using namespace aurdf;

extern "C" size_t aurdf_coord_dist_map_workspace_bytes(int32_t n_frames, int32_t n_coords, int32_t diff) {
    if (n_frames < 0 || n_coords < 0) return 0;
    if (!diff || n_frames < 2) return 256;
    return align_up((size_t)2 * (size_t)(n_frames - 1) * n_coords * n_coords * sizeof(double), 256) + 256;
}

extern "C" int aurdf_coord_dist_map(const double *matrices, int32_t n_frames, int32_t n_coords, double bounding_box,
                                    int32_t diff, double *out_map, double *out_sum, void *workspace,
                                    size_t workspace_bytes, aurdf_stream_t stream_) {
    aurdf::NvtxRange nvtx_range("aurdf_coord_dist_map");
    cudaStream_t stream = (cudaStream_t)stream_;
    AURDF_REQUIRE(n_frames >= 0 && n_coords >= 0, "aurdf_coord_dist_map: negative size");
    AURDF_REQUIRE(n_coords <= kMapMaxK, "aurdf_coord_dist_map: more than 384 clusters");
    AURDF_REQUIRE(bounding_box > 0.0, "aurdf_coord_dist_map: bounding_box must be > 0");
    const int steps = diff ? (n_frames > 0 ? n_frames - 1 : 0) : n_frames;
    const int KK = n_coords * n_coords;
    if (KK == 0) return AURDF_OK;
    AURDF_REQUIRE(out_sum != nullptr, "aurdf_coord_dist_map: NULL out_sum");
    if (steps == 0) {
        AURDF_CUDA_CHECK(cudaMemsetAsync(out_sum, 0, (size_t)KK * sizeof(double), stream));
        return AURDF_OK;
    }
    AURDF_REQUIRE(matrices && out_map, "aurdf_coord_dist_map: NULL pointer");
    double *d_xyz = nullptr, *d_rpy = nullptr;
    if (diff) {
        if (!workspace || workspace_bytes < aurdf_coord_dist_map_workspace_bytes(n_frames, n_coords, diff)) {
            set_error("aurdf_coord_dist_map: workspace_bytes %zu too small", workspace_bytes);
            return AURDF_EWORKSPACE;
        }
        AURDF_REQUIRE(((uintptr_t)workspace & 255) == 0, "aurdf_coord_dist_map: workspace must be 256-byte aligned");
        d_xyz = (double *)workspace;
        d_rpy = d_xyz + (size_t)steps * KK;
    }
    const double lam_rot = 1.0 / 3.141592653589793, lam_bbox = 1.0 / (bounding_box * 2.0);
    const size_t smem = (size_t)16 * n_coords * sizeof(double);
    pair_kernel<<<steps, kMapThreads, smem, stream>>>(matrices, n_frames, n_coords, lam_bbox, lam_rot, diff ? 1 : 0,
                                                      d_xyz, d_rpy, out_map, steps);
    if (diff) rows_kernel<<<steps, kMapThreads, 0, stream>>>(d_xyz, d_rpy, n_coords, steps, out_map);
    sum_kernel<<<(KK + 255) / 256, 256, 0, stream>>>(out_map, KK, steps, out_sum);
    AURDF_CUDA_CHECK(cudaGetLastError());
    return AURDF_OK;
}
