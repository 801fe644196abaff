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


// ------------------------------------------------------------------------------------------
// Fused chamfer distance (both directions + reductions in ONE launch) and its backward.
//
// At the reference's size (P ~ 5000, mlp_reg.py:96, ~600 calls per frame) the operator is launch-bound when
// it is assembled from memset + nn + unpack per direction plus torch reductions.  chamfer_fwd_kernel covers
// both directions with one grid: a work item is (direction, batch element, block of 512 queries, slice of the
// targets).  Every CTA writes the packed (distance bits << 32 | index) minima of its slice to scratch; the
// last CTA to finish a query block (self-resetting arrival counter) merges the slices, emits the indices for
// the backward, and reduces the distances to one partial sum in a fixed order; the last query block of all
// adds the partial sums up -- again in a fixed order, so the loss is deterministic -- and applies pytorch3d's
// point / batch reductions.
// ------------------------------------------------------------------------------------------
constexpr int kCfQ = kChThreads * kChQPT;   // queries per work item

struct ChamferPlan {
    int N, P[2];          // batch, points of x and y
    int nqb[2], nz[2], slice[2];
    int items[2];         // work items per direction
    int groups[2];        // query blocks per direction (N * nqb)
    size_t off_cnt, off_done, off_psum, off_keys[2], total;
};

static ChamferPlan make_chamfer_plan(int N, int P1, int P2) {
    ChamferPlan c;
    c.N = N; c.P[0] = P1; c.P[1] = P2;
    for (int d = 0; d < 2; ++d) {
        const int pq = c.P[d], pt = c.P[1 - d];
        c.nqb[d] = (pq + kCfQ - 1) / kCfQ;
        c.groups[d] = N * c.nqb[d];
        // about two CTAs per SM and direction; a slice is a multiple of 64 targets
        int nz = c.groups[d] > 0 ? (2 * kNumSMs + c.groups[d] - 1) / c.groups[d] : 1;
        int slice = nz > 0 ? (pt + nz - 1) / nz : pt;
        slice = (slice + 63) / 64 * 64;
        if (slice < 64) slice = 64;
        if (slice > kChChunk) slice = kChChunk;
        c.slice[d] = slice;
        c.nz[d] = pt > 0 ? (pt + slice - 1) / slice : 0;
        c.items[d] = c.groups[d] * c.nz[d];
    }
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o = align_up(o + b, 256); return r; };
    c.off_cnt = take((size_t)(c.groups[0] + c.groups[1] + 1) * sizeof(int));
    c.off_done = take(sizeof(int));
    c.off_psum = take((size_t)(c.groups[0] + c.groups[1] + 1) * sizeof(float));
    for (int d = 0; d < 2; ++d) c.off_keys[d] = take((size_t)c.items[d] * kCfQ * sizeof(unsigned long long));
    c.total = o;
    return c;
}

struct ChamferArgs {
    const float *pts[2];
    int *idx[2];
    float *loss;
    int *cnt, *done;
    float *psum;
    unsigned long long *keys[2];
    int N, P[2], nqb[2], nz[2], slice[2], items0, groups[2];
    float w[2];           // reduction weight of a direction: 1 / (points if point mean) / (batch if batch mean)
};

template <int NORM>
__global__ void __launch_bounds__(kChThreads)
chamfer_fwd_kernel(const ChamferArgs a) {
    __shared__ float4 st[kChChunk];
    __shared__ float s_w[kChThreads / 32];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int item = blockIdx.x;
    const int d = item >= a.items0 ? 1 : 0;
    if (d) item -= a.items0;
    const int z = item % a.nz[d];
    const int grp = item / a.nz[d];               // n * nqb + qb
    const int n = grp / a.nqb[d], qb = grp - n * a.nqb[d];
    const int pq = a.P[d], pt = a.P[1 - d];
    const float *q = a.pts[d] + 3 * (size_t)n * pq;
    const float *t = a.pts[1 - d] + 3 * (size_t)n * pt;
    const int j_lo = z * a.slice[d], nj = min(a.slice[d], pt - j_lo);

    float qx[kChQPT], qy[kChQPT], qz[kChQPT], bd[kChQPT];
    int bj[kChQPT];
#pragma unroll
    for (int k = 0; k < kChQPT; ++k) {
        const int i = min(qb * kCfQ + tid + k * kChThreads, pq - 1);
        qx[k] = q[3 * (size_t)i]; qy[k] = q[3 * (size_t)i + 1]; qz[k] = q[3 * (size_t)i + 2];
        bd[k] = INFINITY; bj[k] = 0x7fffffff;
    }
    // stage the slice (padded to a multiple of four with points no query can prefer)
    const int njp = (nj + 3) & ~3;
    for (int j = tid; j < njp; j += kChThreads) {
        if (j < nj) {
            const size_t e = 3 * (size_t)(j_lo + j);
            st[j] = make_float4(t[e], t[e + 1], t[e + 2], 0.f);
        } else {
            st[j] = make_float4(1e18f, 1e18f, 1e18f, 0.f);
        }
    }
    __syncthreads();
    // four targets per trip: one 3-input min tree decides whether the (rare) update branch runs; inside it the
    // first minimum wins, as in pytorch3d's sequential strict '<' scan
    for (int j = 0; j < njp; j += 4) {
        const float4 t0 = st[j], t1 = st[j + 1], t2 = st[j + 2], t3 = st[j + 3];
#pragma unroll
        for (int k = 0; k < kChQPT; ++k) {
            const float d0 = pdist<NORM>(qx[k], qy[k], qz[k], t0), d1 = pdist<NORM>(qx[k], qy[k], qz[k], t1);
            const float d2 = pdist<NORM>(qx[k], qy[k], qz[k], t2), d3 = pdist<NORM>(qx[k], qy[k], qz[k], t3);
            const float m = fminf(fminf(d0, d1), fminf(d2, d3));
            if (m < bd[k]) {
                if (d0 < bd[k]) { bd[k] = d0; bj[k] = j; }
                if (d1 < bd[k]) { bd[k] = d1; bj[k] = j + 1; }
                if (d2 < bd[k]) { bd[k] = d2; bj[k] = j + 2; }
                if (d3 < bd[k]) { bd[k] = d3; bj[k] = j + 3; }
            }
        }
    }
    unsigned long long *keys = a.keys[d] + (size_t)item * kCfQ;
#pragma unroll
    for (int k = 0; k < kChQPT; ++k)
        keys[tid + k * kChThreads] = ((unsigned long long)__float_as_uint(bd[k]) << 32) | (unsigned)(j_lo + bj[k]);
    // arrival: the last CTA of this query block merges the slices
    __threadfence();
    __syncthreads();
    int *cnt = a.cnt + (d ? a.groups[0] : 0) + grp;
    if (tid == 0) s_last = atomicAdd(cnt, 1) == a.nz[d] - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const unsigned long long *gk = a.keys[d] + (size_t)grp * a.nz[d] * kCfQ;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kChQPT; ++k) {
        const int lq = tid + k * kChThreads, i = qb * kCfQ + lq;
        if (i < pq) {
            // eight slices per trip: the loads are independent, so their L2 latencies overlap (one at a time, the
            // merge of ~30 slices was a third of the kernel at P = 5000)
            unsigned long long best = 0xffffffffffffffffULL;
            for (int z0 = 0; z0 < a.nz[d]; z0 += 8) {
                unsigned long long v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = z0 + u < a.nz[d] ? __ldcg(gk + (size_t)(z0 + u) * kCfQ + lq) : 0xffffffffffffffffULL;
#pragma unroll
                for (int u = 0; u < 8; ++u) best = v[u] < best ? v[u] : best;
            }
            a.idx[d][(size_t)n * pq + i] = (int)(best & 0xffffffffu);
            sum += __uint_as_float((unsigned)(best >> 32));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_w[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        float tot = lane < kChThreads / 32 ? s_w[lane] : 0.f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);   // fixed tree: deterministic
        int last_of_all = 0;
        if (lane == 0) {
            a.psum[(d ? a.groups[0] : 0) + grp] = tot;
            *cnt = 0;                                  // ready for the next call
            __threadfence();
            last_of_all = atomicAdd(a.done, 1) == a.groups[0] + a.groups[1] - 1;
        }
        last_of_all = __shfl_sync(0xffffffffu, last_of_all, 0);
        if (last_of_all) {   // last query block of all: the loss, every lane a strided share, then a fixed tree
            __threadfence();
            float l0 = 0.f, l1 = 0.f;
            for (int g = lane; g < a.groups[0]; g += 32) l0 += __ldcg(a.psum + g);
            for (int g = lane; g < a.groups[1]; g += 32) l1 += __ldcg(a.psum + a.groups[0] + g);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                l0 += __shfl_xor_sync(0xffffffffu, l0, o);
                l1 += __shfl_xor_sync(0xffffffffu, l1, o);
            }
            if (lane == 0) {
                *a.loss = l0 * a.w[0] + l1 * a.w[1];
                *a.done = 0;
            }
        }
    }
}

// one thread per point of x, then of y: its own term and the scatter onto its nearest neighbour
template <int NORM>
__global__ void __launch_bounds__(256)
chamfer_bwd_kernel(const float *__restrict__ x, const float *__restrict__ y, const int *__restrict__ idx_x,
                   const int *__restrict__ idx_y, const float *__restrict__ gloss, int N, int P1, int P2, float w1, float w2,
                   float *__restrict__ gx, float *__restrict__ gy) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n1 = (long long)N * P1, n2 = (long long)N * P2;
    if (i >= n1 + n2) return;
    const bool first = i < n1;
    const long long k = first ? i : i - n1;
    const int pq = first ? P1 : P2, pt = first ? P2 : P1;
    const int n = (int)(k / pq);
    const int j = first ? idx_x[k] : idx_y[k];
    const float *a = (first ? x : y) + 3 * k;
    const float *b = (first ? y : x) + 3 * ((size_t)n * pt + j);
    float *ga = first ? gx : gy, *gb = first ? gy : gx;
    const float w = __ldg(gloss) * (first ? w1 : w2);
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {
        const float diff = a[dd] - b[dd];
        float gr;
        if (NORM == 1) gr = w * (float)((diff > 0.f) - (diff < 0.f));
        else gr = w * 2.f * diff;
        if (ga) atomicAdd(ga + 3 * k + dd, gr);
        if (gb) atomicAdd(gb + 3 * ((size_t)n * pt + j) + dd, -gr);
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
    aurdf::NvtxRange nvtx_range("aurdf_nn_f32");
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
    aurdf::NvtxRange nvtx_range("aurdf_nn_f32_bwd");
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

extern "C" size_t aurdf_chamfer_workspace_bytes(int32_t n_batch, int32_t p1, int32_t p2) {
    if (n_batch < 0 || p1 < 0 || p2 < 0) return 0;
    return make_chamfer_plan(n_batch, p1, p2).total;
}

extern "C" int aurdf_chamfer_workspace_init(void *workspace, size_t workspace_bytes, aurdf_stream_t stream_) {
    AURDF_REQUIRE(workspace, "aurdf_chamfer_workspace_init: NULL workspace");
    AURDF_CUDA_CHECK(cudaMemsetAsync(workspace, 0, workspace_bytes, (cudaStream_t)stream_));
    return AURDF_OK;
}

static void chamfer_weights(int N, int P1, int P2, int point_mean, int batch_mean, float *w) {
    w[0] = 1.f / ((point_mean ? (float)(P1 > 0 ? P1 : 1) : 1.f) * (batch_mean ? (float)(N > 0 ? N : 1) : 1.f));
    w[1] = 1.f / ((point_mean ? (float)(P2 > 0 ? P2 : 1) : 1.f) * (batch_mean ? (float)(N > 0 ? N : 1) : 1.f));
}

extern "C" int aurdf_chamfer_fwd(const float *x, const float *y, int32_t N, int32_t P1, int32_t P2, int norm, int point_mean,
                                 int batch_mean, int32_t *idx_x, int32_t *idx_y, float *loss, void *workspace,
                                 size_t workspace_bytes, aurdf_stream_t stream_) {
    aurdf::NvtxRange nvtx_range("aurdf_chamfer_fwd");
    cudaStream_t stream = (cudaStream_t)stream_;
    AURDF_REQUIRE(N > 0 && P1 > 0 && P2 > 0, "aurdf_chamfer_fwd: empty batch or cloud");
    AURDF_REQUIRE(norm == 1 || norm == 2, "aurdf_chamfer_fwd: norm must be 1 or 2");
    AURDF_REQUIRE(x && y && idx_x && idx_y && loss && workspace, "aurdf_chamfer_fwd: NULL pointer");
    const ChamferPlan c = make_chamfer_plan(N, P1, P2);
    if (workspace_bytes < c.total) {
        set_error("aurdf_chamfer_fwd: workspace_bytes %zu < %zu", workspace_bytes, c.total);
        return AURDF_EWORKSPACE;
    }
    char *ws = (char *)workspace;
    ChamferArgs a;
    a.pts[0] = x; a.pts[1] = y; a.idx[0] = idx_x; a.idx[1] = idx_y; a.loss = loss;
    a.cnt = (int *)(ws + c.off_cnt); a.done = (int *)(ws + c.off_done); a.psum = (float *)(ws + c.off_psum);
    a.keys[0] = (unsigned long long *)(ws + c.off_keys[0]); a.keys[1] = (unsigned long long *)(ws + c.off_keys[1]);
    a.N = N;
    for (int d = 0; d < 2; ++d) {
        a.P[d] = c.P[d]; a.nqb[d] = c.nqb[d]; a.nz[d] = c.nz[d]; a.slice[d] = c.slice[d]; a.groups[d] = c.groups[d];
    }
    a.items0 = c.items[0];
    chamfer_weights(N, P1, P2, point_mean, batch_mean, a.w);
    const unsigned grid = (unsigned)(c.items[0] + c.items[1]);
    if (norm == 1) chamfer_fwd_kernel<1><<<grid, kChThreads, 0, stream>>>(a);
    else chamfer_fwd_kernel<2><<<grid, kChThreads, 0, stream>>>(a);
    AURDF_CUDA_CHECK(cudaGetLastError());
    return AURDF_OK;
}

extern "C" int aurdf_chamfer_bwd(const float *x, const float *y, const int32_t *idx_x, const int32_t *idx_y,
                                 const float *grad_loss, int32_t N, int32_t P1, int32_t P2, int norm, int point_mean,
                                 int batch_mean, float *grad_x, float *grad_y, aurdf_stream_t stream_) {
    aurdf::NvtxRange nvtx_range("aurdf_chamfer_bwd");
    cudaStream_t stream = (cudaStream_t)stream_;
    AURDF_REQUIRE(N > 0 && P1 > 0 && P2 > 0, "aurdf_chamfer_bwd: empty batch or cloud");
    AURDF_REQUIRE(norm == 1 || norm == 2, "aurdf_chamfer_bwd: norm must be 1 or 2");
    AURDF_REQUIRE(x && y && idx_x && idx_y && grad_loss, "aurdf_chamfer_bwd: NULL pointer");
    if (!grad_x && !grad_y) return AURDF_OK;
    float w[2];
    chamfer_weights(N, P1, P2, point_mean, batch_mean, w);
    const long long n = (long long)N * P1 + (long long)N * P2;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (norm == 1) chamfer_bwd_kernel<1><<<grid, 256, 0, stream>>>(x, y, idx_x, idx_y, grad_loss, N, P1, P2, w[0], w[1], grad_x, grad_y);
    else chamfer_bwd_kernel<2><<<grid, 256, 0, stream>>>(x, y, idx_x, idx_y, grad_loss, N, P1, P2, w[0], w[1], grad_x, grad_y);
    AURDF_CUDA_CHECK(cudaGetLastError());
    return AURDF_OK;
}
