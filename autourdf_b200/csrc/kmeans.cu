// kmeans.cu -- cluster re-sampling (sm_100a): seeded Lloyd k-means + move into local frames.
//
// Replaces resample_cluster() (AutoURDF PointCloud/mlp_reg.py:172-237, normal=False), the step
// right after the ICP sweep of every frame:
//     sklearn.cluster.k_means(pc_np, init=matrices[:, :3, 3], n_clusters=K, n_init=1)      (:204)
//     cluster_k = (inv(matrices[k]) @ [pc_np[labels == k]; 1])[:3].T                        (:207-213)
// One CTA per frame runs the whole Lloyd iteration of scikit-learn's _kmeans_single_lloyd in
// float64: data centred on its mean, E-step argmin of |c|^2 - 2 x.c (first minimum wins), M-step
// means, empty clusters re-seeded with the farthest points, stop when the labels repeat or when
// the summed squared centre shift <= tol * mean feature variance, final E-step unless the labels
// had converged.  Then a stable (order-preserving) partition of the points by label and the
// inverse SE(3) of each cluster.  A frame is 24 B/point of compulsory reads and 28 B/point of
// writes; the loop itself runs out of L2/shared memory, K*4 FP64 FMAs per point and iteration.
#include <math.h>

#include "common.cuh"

namespace aurdf {

constexpr int kKmThreads = 1024;
constexpr int kKmWarps = kKmThreads / 32;
constexpr int kKmMaxK = 128;

__device__ __forceinline__ double block_sum(double v, double *s_buf) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll 8
    for (int w = 0; w < kKmWarps; ++w) t += s_buf[w];
    return t;
}

__global__ void __launch_bounds__(kKmThreads)
kmeans_resample_kernel(const double *__restrict__ cloud, const int *__restrict__ cloud_off,
                       const double *__restrict__ matrices, int K, int max_iter, double tol,
                       int *__restrict__ labels, double *__restrict__ centers_out, double *__restrict__ local_xyz,
                       int *__restrict__ local_off, int *__restrict__ n_iter_out, double *__restrict__ inertia_out) {
    __shared__ double s_c[kKmMaxK][3];      // current centres (centred coordinates)
    __shared__ double s_cn[kKmMaxK];        // |c|^2
    __shared__ double s_sum[kKmMaxK][3];
    __shared__ int s_cnt[kKmMaxK];
    __shared__ double s_shift[kKmMaxK];
    __shared__ double s_buf[kKmWarps];
    __shared__ double s_inv[kKmMaxK][12];
    __shared__ int s_off[kKmMaxK + 1];
    __shared__ int s_run[kKmMaxK];
    __shared__ unsigned short s_wtab[kKmWarps][kKmMaxK];
    __shared__ double s_bv[kKmWarps];
    __shared__ int s_bi[kKmWarps];
    __shared__ int s_taken[kKmMaxK];
    __shared__ int s_flag;

    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p0 = cloud_off[f], N = cloud_off[f + 1] - p0;
    const double *X = cloud + 3 * (size_t)p0;
    int *lab = labels + p0;
    const double *M = matrices + 16 * (size_t)f * K;

    // ---- centre the data (sklearn: X -= X.mean(axis=0)) and scale the tolerance
    double sx = 0, sy = 0, sz = 0;
    for (int i = tid; i < N; i += kKmThreads) { sx += X[3 * i]; sy += X[3 * i + 1]; sz += X[3 * i + 2]; }
    const double invN = N > 0 ? 1.0 / (double)N : 0.0;
    const double mx = block_sum(sx, s_buf) * invN, my = block_sum(sy, s_buf) * invN, mz = block_sum(sz, s_buf) * invN;
    sx = sy = sz = 0;
    for (int i = tid; i < N; i += kKmThreads) { sx += X[3 * i] - mx; sy += X[3 * i + 1] - my; sz += X[3 * i + 2] - mz; }
    const double cx0 = block_sum(sx, s_buf) * invN, cy0 = block_sum(sy, s_buf) * invN, cz0 = block_sum(sz, s_buf) * invN;
    sx = sy = sz = 0;
    for (int i = tid; i < N; i += kKmThreads) {
        const double a = (X[3 * i] - mx) - cx0, b = (X[3 * i + 1] - my) - cy0, c = (X[3 * i + 2] - mz) - cz0;
        sx += a * a; sy += b * b; sz += c * c;
    }
    const double tol_abs = ((block_sum(sx, s_buf) + block_sum(sy, s_buf) + block_sum(sz, s_buf)) * invN / 3.0) * tol;

    for (int k = tid; k < K; k += kKmThreads) {
        s_c[k][0] = M[16 * k + 3] - mx; s_c[k][1] = M[16 * k + 7] - my; s_c[k][2] = M[16 * k + 11] - mz;
    }
    for (int i = tid; i < N; i += kKmThreads) lab[i] = -1;
    __syncthreads();

    // E-step of one point: argmin_k |c_k|^2 - 2 x.c_k, first minimum
    auto assign = [&](double x, double y, double z) {
        int best = 0;
        double bd = INFINITY;
        for (int k = 0; k < K; ++k) {
            const double d = s_cn[k] - 2.0 * ((x * s_c[k][0] + y * s_c[k][1]) + z * s_c[k][2]);
            if (d < bd) { bd = d; best = k; }
        }
        return best;
    };

    bool strict = false;
    int n_iter = 0;
    for (int it = 0; it < max_iter; ++it) {
        for (int k = tid; k < K; k += kKmThreads) {
            s_cn[k] = (s_c[k][0] * s_c[k][0] + s_c[k][1] * s_c[k][1]) + s_c[k][2] * s_c[k][2];
            s_sum[k][0] = s_sum[k][1] = s_sum[k][2] = 0.0;
            s_cnt[k] = 0;
        }
        __syncthreads();
        int changed = 0;
        for (int i = tid; i < N; i += kKmThreads) {
            const double x = X[3 * i] - mx, y = X[3 * i + 1] - my, z = X[3 * i + 2] - mz;
            const int b = assign(x, y, z);
            changed |= (b != lab[i]);
            lab[i] = b;
            atomicAdd(&s_sum[b][0], x); atomicAdd(&s_sum[b][1], y); atomicAdd(&s_sum[b][2], z);
            atomicAdd(&s_cnt[b], 1);
        }
        const int any_changed = __syncthreads_or(changed);
        // ---- empty clusters: re-seed with the points farthest from their (old) centres
        int n_empty = 0;
        for (int k = 0; k < K; ++k) n_empty += (s_cnt[k] == 0);
        if (n_empty > 0 && N > 0) {
            int done = 0;
            for (int e = 0; e < K; ++e) {
                if (s_cnt[e] != 0) continue;   // uniform: s_cnt[e] of an empty cluster only changes in its own turn
                double bv = -1.0;
                int bi = 0x7fffffff;
                for (int i = tid; i < N; i += kKmThreads) {
                    bool taken = false;
                    for (int t = 0; t < done; ++t) taken |= (s_taken[t] == i);
                    if (taken) continue;
                    const int l = lab[i];
                    const double a = (X[3 * i] - mx) - s_c[l][0], b = (X[3 * i + 1] - my) - s_c[l][1], c = (X[3 * i + 2] - mz) - s_c[l][2];
                    const double d = (a * a + b * b) + c * c;
                    if (d > bv || (d == bv && i < bi)) { bv = d; bi = i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                if (lane == 0) { s_bv[warp] = bv; s_bi[warp] = bi; }
                __syncthreads();
                if (tid == 0) {
                    for (int w = 1; w < kKmWarps; ++w)
                        if (s_bv[w] > bv || (s_bv[w] == bv && s_bi[w] < bi)) { bv = s_bv[w]; bi = s_bi[w]; }
                    const int old = lab[bi];
                    const double x = X[3 * bi] - mx, y = X[3 * bi + 1] - my, z = X[3 * bi + 2] - mz;
                    s_sum[old][0] -= x; s_sum[old][1] -= y; s_sum[old][2] -= z;
                    s_cnt[old] -= 1;
                    s_sum[e][0] = x; s_sum[e][1] = y; s_sum[e][2] = z;
                    s_cnt[e] = 1;
                    s_taken[done] = bi;
                }
                ++done;
                __syncthreads();
            }
        }
        // ---- M-step and centre shift
        for (int k = tid; k < K; k += kKmThreads) {
            const double w = (double)s_cnt[k];
            const double nx = s_sum[k][0] / w, ny = s_sum[k][1] / w, nz = s_sum[k][2] / w;
            const double dx = nx - s_c[k][0], dy = ny - s_c[k][1], dz = nz - s_c[k][2];
            s_shift[k] = (dx * dx + dy * dy) + dz * dz;
            s_c[k][0] = nx; s_c[k][1] = ny; s_c[k][2] = nz;
        }
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int k = 0; k < K; ++k) tot += s_shift[k];
            s_flag = !any_changed ? 1 : (tot <= tol_abs ? 2 : 0);
        }
        __syncthreads();
        n_iter = it + 1;
        if (s_flag == 1) { strict = true; break; }
        if (s_flag == 2) break;
    }
    if (!strict) {   // final E-step so that the labels match the returned centres
        for (int k = tid; k < K; k += kKmThreads)
            s_cn[k] = (s_c[k][0] * s_c[k][0] + s_c[k][1] * s_c[k][1]) + s_c[k][2] * s_c[k][2];
        __syncthreads();
        for (int i = tid; i < N; i += kKmThreads) lab[i] = assign(X[3 * i] - mx, X[3 * i + 1] - my, X[3 * i + 2] - mz);
    }
    __syncthreads();

    // ---- inertia, centres, per-cluster counts of the FINAL labels
    for (int k = tid; k < K; k += kKmThreads) { s_cnt[k] = 0; s_run[k] = 0; }
    __syncthreads();
    double in = 0.0;
    for (int i = tid; i < N; i += kKmThreads) {
        const int l = lab[i];
        const double a = (X[3 * i] - mx) - s_c[l][0], b = (X[3 * i + 1] - my) - s_c[l][1], c = (X[3 * i + 2] - mz) - s_c[l][2];
        in += (a * a + b * b) + c * c;
        atomicAdd(&s_cnt[l], 1);
    }
    in = block_sum(in, s_buf);
    if (tid == 0) {
        if (inertia_out) inertia_out[f] = in;
        if (n_iter_out) n_iter_out[f] = n_iter;
        int acc = 0;
        for (int k = 0; k < K; ++k) { s_off[k] = acc; acc += s_cnt[k]; }
        s_off[K] = acc;
    }
    for (int k = tid; k < K; k += kKmThreads) {
        centers_out[3 * ((size_t)f * K + k)] = s_c[k][0] + mx;
        centers_out[3 * ((size_t)f * K + k) + 1] = s_c[k][1] + my;
        centers_out[3 * ((size_t)f * K + k) + 2] = s_c[k][2] + mz;
        double A[16], Ai[16];
        for (int j = 0; j < 16; ++j) A[j] = M[16 * k + j];
        const bool ok = inv4(A, Ai);
        for (int j = 0; j < 12; ++j) s_inv[k][j] = ok ? Ai[j] : nan("");
    }
    __syncthreads();
    for (int k = tid; k <= K; k += kKmThreads) local_off[(size_t)f * (K + 1) + k] = s_off[k];

    // ---- stable partition by label (pc_np[labels == k] keeps the cloud order) + local frames
    for (int base = 0; base < N; base += kKmThreads) {
        for (int e = tid; e < kKmWarps * K; e += kKmThreads) s_wtab[e / K][e % K] = 0;
        __syncthreads();
        const int i = base + tid;
        const int l = i < N ? lab[i] : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, l);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (l >= 0 && rank == 0) s_wtab[warp][l] = (unsigned short)__popc(peers);
        __syncthreads();
        for (int k = tid; k < K; k += kKmThreads) {
            int acc = s_run[k];
            for (int w = 0; w < kKmWarps; ++w) {
                const int t = s_wtab[w][k];
                s_wtab[w][k] = (unsigned short)acc;   // N per frame is far below 65536 * ... see host check
                acc += t;
            }
            s_run[k] = acc;
        }
        __syncthreads();
        if (l >= 0) {
            const int pos = s_off[l] + s_wtab[warp][l] + rank;
            const double x = X[3 * i], y = X[3 * i + 1], z = X[3 * i + 2];
            const double *T = s_inv[l];
            double *o = local_xyz + 3 * ((size_t)p0 + pos);
            o[0] = T[0] * x + T[1] * y + T[2] * z + T[3];
            o[1] = T[4] * x + T[5] * y + T[6] * z + T[7];
            o[2] = T[8] * x + T[9] * y + T[10] * z + T[11];
        }
        __syncthreads();
    }
}

}  // namespace aurdf

using namespace aurdf;

extern "C" int aurdf_resample_clusters(const double *cloud_xyz, const int32_t *cloud_off, const double *matrices,
                                       int32_t n_frames, int32_t n_clusters, int32_t max_points_per_frame,
                                       int32_t max_iter, double tol, int32_t *out_labels, double *out_centers,
                                       double *out_local_xyz, int32_t *out_local_off, int32_t *out_n_iter,
                                       double *out_inertia, aurdf_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AURDF_REQUIRE(n_frames >= 0 && max_iter >= 0, "aurdf_resample_clusters: negative size");
    AURDF_REQUIRE(n_clusters >= 1 && n_clusters <= kKmMaxK, "aurdf_resample_clusters: n_clusters must be in 1..128");
    AURDF_REQUIRE(max_points_per_frame >= 0 && max_points_per_frame < 65536,
                  "aurdf_resample_clusters: at most 65535 points per frame");
    if (n_frames == 0) return AURDF_OK;
    AURDF_REQUIRE(cloud_xyz && cloud_off && matrices && out_labels && out_centers && out_local_xyz && out_local_off,
                  "aurdf_resample_clusters: NULL pointer");
    kmeans_resample_kernel<<<n_frames, kKmThreads, 0, stream>>>(cloud_xyz, cloud_off, matrices, n_clusters, max_iter, tol,
                                                               out_labels, out_centers, out_local_xyz, out_local_off,
                                                               out_n_iter, out_inertia);
    AURDF_CUDA_CHECK(cudaGetLastError());
    return AURDF_OK;
}
