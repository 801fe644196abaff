// icp_sweep.cu -- batched cluster-ICP sweep for sm_100a (B200).
//
// Replaces masked_icp() (AutoURDF PointCloud/cluster_icp.py:118-191) and the open3d
// registration_icp() call inside it (:157-159), for B (frame, cluster) tiles per call.
//
// Kernels (all on the caller's stream, no host synchronisation):
//   1. box_count_kernel   one CTA per tile: AABB of the predicted cluster (:133-140),
//                         count of frame points strictly inside it (:142-146)
//   2. tile_scan_kernel   exclusive scan of the (even-padded) counts -> compacted offsets,
//                         capacity check
//   3. mask_fill_kernel   order-preserving compaction of the masked target points into
//                         float64 SoA (x|y|z) + original index (:148)
//   4. icp_tiles_kernel   one CTA per tile, persistent to convergence: target chunk staged
//                         in shared memory by the TMA engine (cp.async.bulk + mbarrier),
//                         brute-force squared-L2 argmin in float64 with the reference's
//                         operation order, warp-shuffle reductions, 3x3 Jacobi SVD pose fit,
//                         SE(3) compose/apply, open3d's convergence rule.
//
// Distances and point transforms must round exactly like the float64 CPU reference (no FMA
// contraction), otherwise near-tie correspondences flip: every such operation is written with the
// __d*_rn intrinsics, which the compiler never contracts (the file is built with the default -fmad).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include <cooperative_groups.h>

#include "icp_common.cuh"

namespace cg = cooperative_groups;

namespace aurdf {

constexpr int kIcpThreads = 256;
constexpr int kIcpWarps = kIcpThreads / 32;
constexpr int kQChunk = 1024;     // target points staged per shared-memory chunk (24 KB f64 SoA + 16 KB float4)
constexpr int kClusterCtas = 8;   // CTAs per tile in the cluster variant (large tiles)
constexpr int kPSmemMax = 2048;   // most source points a tile may keep in shared memory
constexpr int kGridPcapMax = 2048; // most source points a CTA of the grid kernel keeps in shared memory (48 B each)

struct WsLayout {
    size_t box, cnt, toff, qx, qy, qz, qi, pspill, status, gs, gends, gpar, glist, total;
};

static WsLayout make_layout(int64_t B, int64_t total_src, int64_t cap) {
    WsLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t r = o;
        o = align_up(o + bytes, 256);
        return r;
    };
    L.box = take((size_t)B * 6 * sizeof(double));
    L.cnt = take((size_t)B * sizeof(int));
    L.toff = take((size_t)(B + 1) * sizeof(long long));
    L.status = take(8 * sizeof(int));
    L.qx = take((size_t)cap * sizeof(double));
    L.qy = take((size_t)cap * sizeof(double));
    L.qz = take((size_t)cap * sizeof(double));
    L.qi = take((size_t)cap * sizeof(int));
    L.pspill = take((size_t)total_src * 3 * sizeof(double));
    L.gs = take((size_t)cap * sizeof(float4));
    L.gends = take((size_t)(cap + 2 * B + 2) * sizeof(int));
    L.gpar = take((size_t)B * 8 * sizeof(float));
    L.glist = take((size_t)B * sizeof(int));
    L.total = o;
    return L;
}

// ------------------------------------------------------------------------------------------
// 1. AABB + count
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool inside_box(double x, double y, double z, const double *bx) {
    return x > bx[0] && x < bx[3] && y > bx[1] && y < bx[4] && z > bx[2] && z < bx[5];
}

__global__ void __launch_bounds__(kIcpThreads)
box_count_kernel(const void *__restrict__ box_xyz, int box_dtype, const int *__restrict__ box_off,
                 const void *__restrict__ tgt_xyz, int pts_dtype, const int *__restrict__ tgt_off,
                 const int *__restrict__ tile_frame, double box_scale, double *__restrict__ box_out,
                 int *__restrict__ cnt_out) {
    __shared__ double s_mm[kIcpWarps][6];
    __shared__ double s_box[6];
    __shared__ int s_cnt[kIcpWarps];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool no_mask = box_xyz == nullptr;  // plain registration_icp: every frame point is a target
    const int b0 = no_mask ? 0 : box_off[b], nb = no_mask ? 0 : box_off[b + 1] - b0;

    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < nb; i += kIcpThreads) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double v = ld_coord(box_xyz, box_dtype, 3 * (size_t)(b0 + i) + d);
            lo[d] = fmin(lo[d], v);
            hi[d] = fmax(hi[d], v);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            s_mm[warp][d] = lo[d];
            s_mm[warp][3 + d] = hi[d];
        }
    }
    __syncthreads();
    if (tid == 0) {
        for (int d = 0; d < 3; ++d) {
            double l = s_mm[0][d], h = s_mm[0][3 + d];
            for (int w = 1; w < kIcpWarps; ++w) {
                l = fmin(l, s_mm[w][d]);
                h = fmax(h, s_mm[w][3 + d]);
            }
            double blo, bhi;
            if (no_mask) {
                blo = -INFINITY;
                bhi = INFINITY;
            } else if (nb <= 0) {
                blo = INFINITY;
                bhi = -INFINITY;
            } else if (box_dtype == AURDF_F32) {
                // numpy keeps float32 through np.mean / python-scalar multiply (cluster_icp.py:138-140)
                float lf = (float)l, hf = (float)h;
                float c = __fmul_rn(__fadd_rn(lf, hf), 0.5f);
                float size = __fsub_rn(hf, lf);
                float hs = __fmul_rn((float)(0.5 * box_scale), size);
                blo = (double)__fsub_rn(c, hs);
                bhi = (double)__fadd_rn(c, hs);
            } else {
                double c = __dmul_rn(__dadd_rn(l, h), 0.5);
                double size = __dsub_rn(h, l);
                double hs = __dmul_rn(0.5 * box_scale, size);
                blo = __dsub_rn(c, hs);
                bhi = __dadd_rn(c, hs);
            }
            s_box[d] = blo;
            s_box[3 + d] = bhi;
            box_out[6 * (size_t)b + d] = blo;
            box_out[6 * (size_t)b + 3 + d] = bhi;
        }
    }
    __syncthreads();
    const int f = tile_frame[b];
    const int t0 = tgt_off[f], M = tgt_off[f + 1] - t0;
    int c = 0;
    for (int i = tid; i < M; i += kIcpThreads) {
        const size_t e = 3 * (size_t)(t0 + i);
        double x = ld_coord(tgt_xyz, pts_dtype, e), y = ld_coord(tgt_xyz, pts_dtype, e + 1),
               z = ld_coord(tgt_xyz, pts_dtype, e + 2);
        c += inside_box(x, y, z, s_box) ? 1 : 0;
    }
    c = warp_sum_int(c);
    if (lane == 0) s_cnt[warp] = c;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < kIcpWarps; ++w) t += s_cnt[w];
        cnt_out[b] = t;
    }
}

// ------------------------------------------------------------------------------------------
// 2. scan of even-padded counts (single CTA; B is at most a few 10^5)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
tile_scan_kernel(const int *__restrict__ cnt, int B, long long capacity, long long *__restrict__ toff,
                 int *__restrict__ status_int, int *__restrict__ status_user, const IcpParams P) {
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    __shared__ int s_wc[32];
    __shared__ int s_gbase;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        int i = base + tid;
        long long v = (i < B) ? (long long)((cnt[i] + 1) & ~1) : 0;
        long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            long long w = s_warp[lane];
            long long winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                long long n = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += n;
            }
            s_warp[lane] = winc - w;  // exclusive
        }
        __syncthreads();
        long long excl = s_carry + s_warp[warp] + inc - v;
        if (i < B) toff[i] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
    // the tiles the grid kernel owns, in tile order: its persistent CTAs take them from this list, so a sweep of
    // small tiles costs them one atomic each instead of a walk over the tile table
    if (tid == 0) s_gbase = 0;
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        const int i = base + tid;
        const bool g = i < B && tile_uses_grid(P, P.src_off[i + 1] - P.src_off[i], cnt[i]);
        const unsigned bal = __ballot_sync(0xffffffffu, g);
        if (lane == 0) s_wc[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            const int c = s_wc[w];
            before += w < warp ? c : 0;
            total += c;
        }
        if (g) P.glist[s_gbase + before + __popc(bal & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (tid == 0) s_gbase += total;
        __syncthreads();
    }
    if (tid == 0) {
        status_int[5] = s_gbase;   // number of listed tiles
        long long total = s_carry;
        toff[B] = total;
        int over = total > capacity ? 1 : 0;
        status_int[0] = over;
        status_int[1] = (int)(total & 0xffffffffLL);
        status_int[2] = (int)(total >> 32);
        status_int[3] = 0;   // tile queue of the small-tile kernel
        status_int[4] = 0;   // tile queue of the grid kernel
        if (status_user) {
            status_user[0] = over;
            status_user[1] = status_int[1];
            status_user[2] = status_int[2];
            status_user[3] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// 3. order-preserving compaction of the masked target points
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kIcpThreads)
mask_fill_kernel(const void *__restrict__ tgt_xyz, int pts_dtype, const int *__restrict__ tgt_off,
                 const int *__restrict__ tile_frame, const double *__restrict__ box,
                 const long long *__restrict__ toff, const int *__restrict__ status_int,
                 double *__restrict__ qx, double *__restrict__ qy, double *__restrict__ qz,
                 int *__restrict__ qi) {
    if (status_int[0]) return;
    __shared__ double s_box[6];
    __shared__ int s_wcnt[kIcpWarps];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 6) s_box[tid] = box[6 * (size_t)b + tid];
    __syncthreads();
    const int f = tile_frame[b];
    const int t0 = tgt_off[f], M = tgt_off[f + 1] - t0;
    const long long base = toff[b];
    int running = 0;
    for (int start = 0; start < M; start += kIcpThreads) {
        const int i = start + tid;
        double x = 0, y = 0, z = 0;
        bool in = false;
        if (i < M) {
            const size_t e = 3 * (size_t)(t0 + i);
            x = ld_coord(tgt_xyz, pts_dtype, e);
            y = ld_coord(tgt_xyz, pts_dtype, e + 1);
            z = ld_coord(tgt_xyz, pts_dtype, e + 2);
            in = inside_box(x, y, z, s_box);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kIcpWarps; ++w) {
            int c = s_wcnt[w];
            before += (w < warp) ? c : 0;
            total += c;
        }
        if (in) {
            const long long pos = base + running + before + __popc(bal & ((1u << lane) - 1u));
            qx[pos] = x;
            qy[pos] = y;
            qz[pos] = z;
            qi[pos] = i;
        }
        running += total;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// 4. fused per-tile ICP
// ------------------------------------------------------------------------------------------
// One CTA per tile, resident until that tile's ICP has converged.  Per iteration:
//   pass      every warp: P <- U P (fused), brute-force NN of its source points against the
//             target chunk in shared memory, 16 moment sums + inlier count per warp
//   barrier A
//   warp 0    lane 0: Kabsch / Jacobi-SVD pose update from the block totals -> U
//   warp 1    lane 0: fitness, rmse and open3d's convergence test          -> stop flag
//   barrier B
//   warp 3    16 lanes: T <- U T (off the critical path, overlaps the next pass)
// CS > 1: a thread-block cluster of CS CTAs shares one (large) tile.  Every CTA owns a contiguous
// slice of the source points and scans all targets; the per-CTA moment sums meet in the leader's
// shared memory through DSMEM, the leader fits the pose and tests convergence, and the other CTAs
// read the update back -- two cluster barriers per iteration.
template <int CS>
__global__ void __launch_bounds__(kIcpThreads, 2)
icp_tiles_kernel(const IcpParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sqx = reinterpret_cast<double *>(smem_raw);
    double *sqy = sqx + kQChunk;
    double *sqz = sqy + kQChunk;
    float4 *sqf = reinterpret_cast<float4 *>(sqz + kQChunk);  // float32 copy (x,y,z about the tile origin)
    double *spx_s = reinterpret_cast<double *>(sqf + kQChunk);
    double *spy_s = spx_s + p.p_cap;
    double *spz_s = spy_s + p.p_cap;
    int *scj_s = reinterpret_cast<int *>(spz_s + p.p_cap);

    __shared__ double s_part[kIcpWarps][16];  // per-warp moment sums
    __shared__ int s_cnt[kIcpWarps];          // per-warp inlier counts
    __shared__ double s_tot[2][16];           // block totals, one copy per consumer warp
    __shared__ double s_U[16];                // current update (row-major 4x4)
    __shared__ double s_T[16];                // accumulated pose
    __shared__ double s_prev[2];              // fitness, rmse of the previous pass
    __shared__ double s_warm[18];             // singular vectors of the previous fit
    __shared__ int s_stop;
    __shared__ double s_cl[CS][16];           // leader only: per-CTA moment sums of the cluster
    __shared__ int s_clc[CS];                 // leader only: per-CTA inlier counts
    __shared__ float s_amax[kIcpWarps];       // max |target coordinate - origin| per warp
    __shared__ __align__(8) uint64_t s_bar;

    if (p.status_int[0]) return;  // compacted-target capacity exceeded: leave outputs untouched

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x / CS;                       // tile
    int rank = 0;                                        // CTA within the tile's cluster
    if constexpr (CS > 1) rank = (int)cg::this_cluster().block_rank();
    const int ns_tile = p.src_off[b + 1] - p.src_off[b];
    // a tile whose box holds a handful of targets fits its poses in the reference's own arithmetic
    // (icp_common.cuh, namespace strict); its sums run in ascending source index, so one CTA owns all of it
    const bool strict_tile = p.cnt[b] <= p.strict_nt;
    const int per = strict_tile ? ns_tile : (ns_tile + CS - 1) / CS;
    const int lo = min(ns_tile, rank * per);
    const int s0 = p.src_off[b] + lo;                    // this CTA's slice of the source points
    const int ns = min(ns_tile, lo + per) - lo;
    const int nt = p.cnt[b];
    const long long q0 = p.toff[b];
    if (tile_uses_small(p, ns_tile, nt) || tile_uses_grid(p, ns_tile, nt)) return;   // another kernel owns this tile (whole cluster leaves)
    if constexpr (CS > 1) cg::this_cluster().sync();   // every CTA of the cluster runs before any touches a peer's shared memory
    const bool resident = nt <= kQChunk;
    const int nchunks = (nt + kQChunk - 1) / kQChunk;

    // source-point state: shared memory when it fits, else the global spill area
    double *px, *py, *pz;
    int *cj;
    if (ns <= p.p_cap) {
        px = spx_s; py = spy_s; pz = spz_s; cj = scj_s;
    } else {
        px = p.pspill + 3 * (size_t)s0; py = px + ns; pz = py + ns; cj = p.out_corr + s0;
    }

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        s_stop = 0;
    }
    if (tid < 16) {
        s_T[tid] = p.init_T[16 * (size_t)b + tid];
        s_U[tid] = (tid % 5 == 0) ? 1.0 : 0.0;
    }
    __syncthreads();
    uint32_t bar_phase = 0;

    // stage one target chunk with the TMA engine: 3 bulk copies (x, y, z) on one mbarrier
    auto load_chunk = [&](int c) {
        const int n = min(kQChunk, nt - c * kQChunk);
        const uint32_t bytes = (uint32_t)(((n + 1) & ~1) * sizeof(double));
        if (tid == 0) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&s_bar, 3 * bytes);
            const long long o = q0 + (long long)c * kQChunk;
            bulk_g2s(sqx, p.qx + o, bytes, &s_bar);
            bulk_g2s(sqy, p.qy + o, bytes, &s_bar);
            bulk_g2s(sqz, p.qz + o, bytes, &s_bar);
        }
        mbar_wait(&s_bar, bar_phase);
        bar_phase ^= 1;
    };
    if (resident && nt > 0) load_chunk(0);

    // P <- T0 * S
    {
        const bool aff0 = s_T[12] == 0.0 && s_T[13] == 0.0 && s_T[14] == 0.0 && s_T[15] == 1.0;
        for (int i = tid; i < ns; i += kIcpThreads) {
            const size_t e = 3 * (size_t)(s0 + i);
            double x = ld_coord(p.src, p.pts_dtype, e), y = ld_coord(p.src, p.pts_dtype, e + 1),
                   z = ld_coord(p.src, p.pts_dtype, e + 2);
            transform_point(s_T, aff0, x, y, z);
            px[i] = x; py[i] = y; pz[i] = z;
        }
    }
    __syncthreads();

    // split factor: S lanes share one source point when the tile is narrower than the CTA
    int S = 1;
    while (S < 32 && ns * (S * 2) <= kIcpThreads) S *= 2;
    const int pts_per_round = kIcpThreads / S;
    const int rounds = (ns + pts_per_round - 1) / pts_per_round;
    const int sub = tid & (S - 1);
    // moments are accumulated about the tile's first target point (kills cancellation)
    const double ox = nt > 0 ? __ldg(p.qx + q0) : 0.0, oy = nt > 0 ? __ldg(p.qy + q0) : 0.0,
                 oz = nt > 0 ? __ldg(p.qz + q0) : 0.0;

    // float32 pre-filter: targets about the origin, rounded to float32 (built once for a resident
    // tile, per staged chunk for a streamed one), and the largest coordinate magnitude A_q, which
    // scales the rounding-error bound of the filter
    const bool use_f32 = nt > 0;
    float aq = 0.f;
    if (use_f32) {
        float amax = 0.f;
        for (int j = tid; j < nt; j += kIcpThreads) {
            float fx, fy, fz;
            if (resident) {
                fx = (float)(sqx[j] - ox); fy = (float)(sqy[j] - oy); fz = (float)(sqz[j] - oz);
                sqf[j] = make_float4(fx, fy, fz, 0.f);
            } else {
                fx = (float)(__ldg(p.qx + q0 + j) - ox); fy = (float)(__ldg(p.qy + q0 + j) - oy);
                fz = (float)(__ldg(p.qz + q0 + j) - oz);
            }
            amax = fmaxf(amax, fmaxf(fabsf(fx), fmaxf(fabsf(fy), fabsf(fz))));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        if (lane == 0) s_amax[warp] = amax;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < kIcpWarps; ++w) aq = fmaxf(aq, s_amax[w]);
    }

    long long dbg_n = 0;
    auto stamp = [&]() {
        if (p.dbg_clock && blockIdx.x == 0 && tid == 0 && dbg_n < 4096) p.dbg_clock[dbg_n++] = clock64();
    };

    // One correspondence pass.  apply: first move the points by the current update s_U.
    // Leaves per-warp partial sums in s_part / s_cnt (valid after the next barrier).
    auto pass = [&](bool apply) {
        double acc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = 0.0;
        int cnt = 0;
        for (int r = 0; r < rounds; ++r) {
            const int i = tid / S + r * pts_per_round;
            const bool active = i < ns;
            double x = 0, y = 0, z = 0;
            if (active) {
                x = px[i]; y = py[i]; z = pz[i];
                if (apply) transform_point(s_U, true, x, y, z);
            }
            if (apply) {
                if (S > 1) __syncwarp();  // the S lanes of a point have all read the old value
                if (active && sub == 0) { px[i] = x; py[i] = y; pz[i] = z; }
            }
            if (apply) stamp();  // [+1] P update done
            double bd = INFINITY;
            int bj = -1;
            // ---- float32 pre-filter: best and second-best float32 distance of this point.  If the
            // gap exceeds twice the worst-case float32 error the float32 winner is provably the
            // float64 argmin and only its exact distance is evaluated; otherwise the point takes
            // the exact scan below (probability ~1e-3 per point).
            bool need_exact = active;
            if (use_f32) {
                float m1 = INFINITY, m2 = INFINITY;
                int j1 = -1;
                const float fx = (float)(x - ox), fy = (float)(y - oy), fz = (float)(z - oz);
                for (int c = 0; c < nchunks; ++c) {
                    const int n = min(kQChunk, nt - c * kQChunk);
                    const int jbase = c * kQChunk;
                    if (!resident) {   // stage the chunk and its float32 copy
                        __syncthreads();
                        load_chunk(c);
                        for (int j = tid; j < n; j += kIcpThreads)
                            sqf[j] = make_float4((float)(sqx[j] - ox), (float)(sqy[j] - oy), (float)(sqz[j] - oz), 0.f);
                        __syncthreads();
                    }
                    if (active) {
#pragma unroll 4
                        for (int j = sub; j < n; j += S) {
                            const float4 q = sqf[j];
                            const float dx = fx - q.x, dy = fy - q.y, dz = fz - q.z;
                            const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                            m2 = fminf(m2, fmaxf(m1, d));
                            j1 = d < m1 ? jbase + j : j1;
                            m1 = fminf(m1, d);
                        }
                    }
                }
                for (int o = S >> 1; o > 0; o >>= 1) {
                    const float om1 = __shfl_xor_sync(0xffffffffu, m1, o), om2 = __shfl_xor_sync(0xffffffffu, m2, o);
                    const int oj1 = __shfl_xor_sync(0xffffffffu, j1, o);
                    m2 = fminf(fminf(m2, om2), fmaxf(m1, om1));
                    if (om1 < m1 || (om1 == m1 && oj1 >= 0 && (j1 < 0 || oj1 < j1))) { m1 = om1; j1 = oj1; }
                }
                if (active && j1 >= 0) {
                    // |d32 - d_true| <= 3.01u(R + 2 sqrt3 dl sqrtR + 3 dl^2) + 2 sqrt3 dl sqrtR + 3 dl^2 with
                    // u = 2^-24 and dl <= 4 u max(A_q, |a|) the per-coordinate error of a float32 difference;
                    // tau over-covers that by more than 4x
                    const float u = 5.9604645e-8f;
                    const float amag = fmaxf(aq, fmaxf(fabsf(fx), fmaxf(fabsf(fy), fabsf(fz))));
                    const float dl = 4.f * u * amag;
                    const float tau = 16.f * (dl * sqrtf(m2) * 1.001f + dl * dl + u * m2);
                    if (nt == 1 || (m2 - m1 > 2.f * tau && m2 < INFINITY)) {
                        need_exact = false;
                        if (sub == 0) {
                            double qx_, qy_, qz_;
                            if (resident) { qx_ = sqx[j1]; qy_ = sqy[j1]; qz_ = sqz[j1]; }
                            else { qx_ = __ldg(p.qx + q0 + j1); qy_ = __ldg(p.qy + q0 + j1); qz_ = __ldg(p.qz + q0 + j1); }
                            const double dx = __dsub_rn(x, qx_), dy = __dsub_rn(y, qy_), dz = __dsub_rn(z, qz_);
                            bd = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                            bj = j1;
                        }
                    }
                }
            }
            // exact rescan: per warp for a resident tile; block-wide vote for a streamed one, whose
            // chunk loop holds block barriers
            const bool warp_scans = resident ? (bool)__any_sync(0xffffffffu, need_exact) : (bool)__syncthreads_or(need_exact);
            for (int c = 0; warp_scans && c < nchunks; ++c) {
                if (!resident) {
                    __syncthreads();  // everyone done with the previous chunk
                    load_chunk(c);
                }
                const int n = min(kQChunk, nt - c * kQChunk);
                const int jbase = c * kQChunk;
                if (need_exact) {
                    // four independent distance chains per trip, merged by a min-tree (lower index
                    // wins ties at every node, as a sequential strict '<' scan would): the only
                    // loop-carried dependency is one compare per four targets
                    auto dist2 = [&](int j) {
                        const double dx = __dsub_rn(x, sqx[j]), dy = __dsub_rn(y, sqy[j]), dz = __dsub_rn(z, sqz[j]);
                        return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    };
                    int j = sub;
                    for (; j + 3 * S < n; j += 4 * S) {
                        const double d0 = dist2(j), d1 = dist2(j + S), d2 = dist2(j + 2 * S), d3 = dist2(j + 3 * S);
                        const bool p01 = d1 < d0, p23 = d3 < d2;
                        const double m01 = p01 ? d1 : d0, m23 = p23 ? d3 : d2;
                        const int j01 = p01 ? j + S : j, j23 = p23 ? j + 3 * S : j + 2 * S;
                        const bool pm = m23 < m01;
                        const double m = pm ? m23 : m01;
                        const int jm = pm ? j23 : j01;
                        if (m < bd) { bd = m; bj = jbase + jm; }
                    }
                    for (; j < n; j += S) {
                        const double d = dist2(j);
                        if (d < bd) { bd = d; bj = jbase + j; }
                    }
                }
            }
            if (apply) stamp();  // [+2] NN scan done
            // (d, j) lexicographic min across the S lanes of this point
            for (int o = S >> 1; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
                if (oj >= 0 && (od < bd || (od == bd && oj < bj) || bj < 0)) { bd = od; bj = oj; }
            }
            const bool inl = active && sub == 0 && bj >= 0 && bd < p.r2;
            if (active && sub == 0) cj[i] = inl ? bj : -1;
            cnt += __popc(__ballot_sync(0xffffffffu, inl));
            if (inl) {
                double qx_, qy_, qz_;
                if (resident) { qx_ = sqx[bj]; qy_ = sqy[bj]; qz_ = sqz[bj]; }
                else { qx_ = __ldg(p.qx + q0 + bj); qy_ = __ldg(p.qy + q0 + bj); qz_ = __ldg(p.qz + q0 + bj); }
                const double ax = x - ox, ay = y - oy, az = z - oz;
                const double bx = qx_ - ox, by = qy_ - oy, bz = qz_ - oz;
                acc[0] += bd;
                acc[1] += ax; acc[2] += ay; acc[3] += az;
                acc[4] += bx; acc[5] += by; acc[6] += bz;
                acc[7] += bx * ax; acc[8] += bx * ay; acc[9] += bx * az;
                acc[10] += by * ax; acc[11] += by * ay; acc[12] += by * az;
                acc[13] += bz * ax; acc[14] += bz * ay; acc[15] += bz * az;
            }
        }
        if (apply) stamp();  // [+3] merge + moment accumulation done
        warp_sum16(acc, lane);
        if ((lane & 1) == 0) s_part[warp][lane >> 1] = acc[0];
        if (lane == 0) s_cnt[warp] = cnt;
    };

    // block totals for consumer warp w (0: pose fit, 1: convergence test); returns the count
    auto totals = [&](int w) -> int {
        if (lane < 16) {
            double t = s_part[0][lane];
#pragma unroll
            for (int k = 1; k < kIcpWarps; ++k) t += s_part[k][lane];
            s_tot[w][lane] = t;
        }
        int c = 0;
#pragma unroll
        for (int k = 0; k < kIcpWarps; ++k) c += s_cnt[k];
        __syncwarp();
        return c;
    };
    // cluster variant: totals over all CTAs of the tile, from the slots the CTAs filled in the leader
    auto cluster_totals = [&](int w) -> int {
        if (lane < 16) {
            double t = s_cl[0][lane];
#pragma unroll
            for (int r = 1; r < CS; ++r) t += s_cl[r][lane];
            s_tot[w][lane] = t;
        }
        int c = 0;
#pragma unroll
        for (int r = 0; r < CS; ++r) c += s_clc[r];
        __syncwarp();
        return c;
    };
    // every CTA: publish its totals to the leader, then the two cluster barriers around the fit
    auto cluster_publish = [&]() {
        if constexpr (CS > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            if (warp == 0) {
                const int c = totals(0);
                double *dst = cluster.map_shared_rank(&s_cl[0][0], 0);
                int *dstc = cluster.map_shared_rank(&s_clc[0], 0);
                if (lane < 16) dst[rank * 16 + lane] = s_tot[0][lane];
                if (lane == 16) dstc[rank] = c;
            }
            cluster.sync();
        }
    };
    auto cluster_fetch = [&]() {
        if constexpr (CS > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            cluster.sync();
            if (rank != 0) {
                const double *srcU = cluster.map_shared_rank(&s_U[0], 0);
                const int *srcS = cluster.map_shared_rank(&s_stop, 0);
                if (tid < 12) s_U[tid] = srcU[tid];
                if (tid == 12) s_stop = *srcS;
            }
        }
    };

    // Kabsch / umeyama update from the totals of consumer slot 0 (lane 0 of warp 0)
    bool have_warm = false;
    auto fit_pose = [&](int c) {
        const double *t = s_tot[0];
        double Um[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        if (c > 0) {
            const double inv = rcp_nr2((double)c);
            const double ma[3] = {t[1] * inv, t[2] * inv, t[3] * inv};
            const double mb[3] = {t[4] * inv, t[5] * inv, t[6] * inv};
            double sigma[3][3], R[3][3];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) sigma[r][cc] = t[7 + 3 * r + cc] * inv - mb[r] * ma[cc];
            if (!kabsch_rotation_newton(sigma, R)) {   // reflection / rank-deficient / large step
                double sg2[3][3], R2[3][3];   // copies: the out-of-line call takes addresses, sigma / R stay in registers
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) sg2[r][cc] = sigma[r][cc];
                kabsch_rotation(sg2, R2, s_warm, have_warm);
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) R[r][cc] = R2[r][cc];
                have_warm = true;
            }
            const double mua[3] = {ma[0] + ox, ma[1] + oy, ma[2] + oz};
            const double mub[3] = {mb[0] + ox, mb[1] + oy, mb[2] + oz};
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                Um[4 * r + 0] = R[r][0]; Um[4 * r + 1] = R[r][1]; Um[4 * r + 2] = R[r][2];
                Um[4 * r + 3] = mub[r] - (R[r][0] * mua[0] + R[r][1] * mua[1] + R[r][2] * mua[2]);
            }
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) s_U[k] = Um[k];
    };

    // strict pose fit (warp 0, all lanes): two-pass sums in ascending source index, one accumulator per lane,
    // then the reference's Jacobi SVD on lane 0.  Targets of such a tile are resident (n_t <= strict_nt).
    auto fit_strict = [&](int c) {
        double Um[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        if (c > 0) {   // warp-uniform
            const double *src1 = lane == 0 ? px : (lane == 1 ? py : (lane == 2 ? pz : (lane == 3 ? sqx : (lane == 4 ? sqy : sqz))));
            double acc = 0.0;
            if (lane < 6) {
                for (int i = 0; i < ns; ++i) {
                    const int j = cj[i];
                    if (j >= 0) acc = strict::add(acc, src1[lane < 3 ? i : j]);
                }
            }
            const double one_over_n = strict::dvd(1.0, (double)c);
            const double mean = strict::mul(acc, one_over_n);   // lanes 0-2: source mean, 3-5: target mean
            const int rr = lane < 9 ? lane / 3 : 0, cc = lane < 9 ? lane - 3 * rr : 0;
            const double ms_c = __shfl_sync(0xffffffffu, mean, cc), md_r = __shfl_sync(0xffffffffu, mean, 3 + rr);
            const double *pa = cc == 0 ? px : (cc == 1 ? py : pz);
            const double *pb = rr == 0 ? sqx : (rr == 1 ? sqy : sqz);
            double sg = 0.0;
            if (lane < 9) {
                for (int i = 0; i < ns; ++i) {
                    const int j = cj[i];
                    if (j >= 0) sg = strict::add(sg, strict::mul(strict::sub(pb[j], md_r), strict::sub(pa[i], ms_c)));
                }
                sg = strict::mul(sg, one_over_n);
            }
            double sigma[3][3], ms[3], md[3];
#pragma unroll
            for (int k = 0; k < 9; ++k) sigma[k / 3][k % 3] = __shfl_sync(0xffffffffu, sg, k);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                ms[k] = __shfl_sync(0xffffffffu, mean, k);
                md[k] = __shfl_sync(0xffffffffu, mean, 3 + k);
            }
            if (lane == 0) strict::pose_from_sigma(sigma, ms, md, Um);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 12; ++k) s_U[k] = Um[k];
        }
    };

    // T <- U * T, one lane per entry, entries summed left to right with each operation rounded
    auto compose_pose = [&]() {
        double v = 0.0;
        if (lane < 16) {
            const int r = lane >> 2, cc = lane & 3;
            v = __dmul_rn(s_U[4 * r], s_T[cc]);
            v = __dadd_rn(v, __dmul_rn(s_U[4 * r + 1], s_T[4 + cc]));
            v = __dadd_rn(v, __dmul_rn(s_U[4 * r + 2], s_T[8 + cc]));
            v = __dadd_rn(v, __dmul_rn(s_U[4 * r + 3], s_T[12 + cc]));
        }
        __syncwarp();
        if (lane < 16) s_T[lane] = v;
    };

    stamp();
    pass(false);
    __syncthreads();  // barrier A
    cluster_publish();
    if (rank == 0) {
        if (warp == 0) {
            const int c = CS > 1 ? cluster_totals(0) : totals(0);
            if (p.max_iter > 0) {
                if (strict_tile) fit_strict(c);
                else if (lane == 0) fit_pose(c);
            }
        } else if (warp == 1) {
            const int c = CS > 1 ? cluster_totals(1) : totals(1);
            if (lane == 0) {
                s_prev[0] = c > 0 ? (double)c / (double)ns_tile : 0.0;
                s_prev[1] = c > 0 ? sqrt(s_tot[1][0] / (double)c) : 0.0;
            }
        }
    }
    cluster_fetch();
    __syncthreads();  // barrier B

    int iters = 0;
    for (int it = 0; it < p.max_iter; ++it) {
        stamp();  // [7k+0] iteration start
        if (rank == 0 && warp == kIcpWarps - 1) compose_pose();  // uses s_U of this iteration; next write is after barrier A
        pass(true);
        stamp();  // [+4] pass done (warp 0)
        __syncthreads();  // barrier A
        stamp();  // [+5] barrier A passed
        cluster_publish();
        if (rank == 0) {
            if (warp == 0) {
                // speculative: the fit for iteration it+1 runs while warp 1 decides whether to stop
                const int c = CS > 1 ? cluster_totals(0) : totals(0);
                if (it + 1 < p.max_iter) {
                    if (strict_tile) fit_strict(c);
                    else if (lane == 0) fit_pose(c);
                }
            } else if (warp == 1) {
                const int c = CS > 1 ? cluster_totals(1) : totals(1);
                if (lane == 0) {
                    const double fit = c > 0 ? (double)c / (double)ns_tile : 0.0;
                    const double rmse = c > 0 ? sqrt(s_tot[1][0] / (double)c) : 0.0;
                    s_stop = (fabs(s_prev[0] - fit) < p.rel_fit && fabs(s_prev[1] - rmse) < p.rel_rmse) ? 1 : 0;
                    s_prev[0] = fit;
                    s_prev[1] = rmse;
                }
            }
        }
        stamp();  // [+6] fit done
        cluster_fetch();
        __syncthreads();  // barrier B
        iters = it + 1;
        if (s_stop) break;
    }
    __syncthreads();

    // outputs: pose (cluster_icp.py:161-165), world cluster = T * S (:167), correspondences
    if (rank == 0 && tid == 0) {
        if (p.ori_only) {
            s_T[3] = p.init_T[16 * (size_t)b + 3];
            s_T[7] = p.init_T[16 * (size_t)b + 7];
            s_T[11] = p.init_T[16 * (size_t)b + 11];
        }
        p.out_fit[b] = s_prev[0];
        p.out_rmse[b] = s_prev[1];
        p.out_iters[b] = iters;
        p.out_ntgt[b] = nt;
    }
    __syncthreads();
    if constexpr (CS > 1) {   // every CTA of the cluster needs the leader's final pose
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();
        if (rank != 0 && tid < 16) s_T[tid] = cluster.map_shared_rank(&s_T[0], 0)[tid];
        cluster.sync();       // the leader's shared memory must outlive the reads
        __syncthreads();
    }
    if (rank == 0 && tid < 16) p.out_T[16 * (size_t)b + tid] = s_T[tid];
    const bool aff = s_T[12] == 0.0 && s_T[13] == 0.0 && s_T[14] == 0.0 && s_T[15] == 1.0;
    for (int i = tid; i < ns; i += kIcpThreads) {
        const size_t e = 3 * (size_t)(s0 + i);
        double x = ld_coord(p.src, p.pts_dtype, e), y = ld_coord(p.src, p.pts_dtype, e + 1),
               z = ld_coord(p.src, p.pts_dtype, e + 2);
        transform_point(s_T, aff, x, y, z);
        p.out_world[e] = x; p.out_world[e + 1] = y; p.out_world[e + 2] = z;
        const int j = cj[i];
        p.out_corr[s0 + i] = j >= 0 ? __ldg(p.qi + q0 + j) : -1;
    }
}

}  // namespace aurdf

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace aurdf;

extern "C" size_t aurdf_icp_workspace_bytes(int32_t n_tiles, int64_t total_src_points, int64_t tgt_capacity) {
    if (n_tiles < 0 || total_src_points < 0 || tgt_capacity < 0) return 0;
    return make_layout(n_tiles, total_src_points, tgt_capacity).total;
}

// Debug hook of the measurement scripts (scripts/tile_latency.py): per-phase cycle stamps of tile 0.
// Thread-local, like every other piece of library state: a sweep only sees what its own thread set.
static thread_local long long *g_dbg_clock = nullptr;
extern "C" __attribute__((visibility("default"))) void aurdf_debug_set_clock_buffer(void *p) { g_dbg_clock = (long long *)p; }

namespace aurdf {
int launch_icp_small(const IcpParams &P, int n_tiles, int minb, cudaStream_t stream);
int launch_icp_small2(const IcpParams &P, int n_tiles, int minb, cudaStream_t stream);
int launch_icp_grid(const IcpParams &P, int n_tiles, cudaStream_t stream);
}

// Tuning knobs for A/B measurements, read from the environment once (immutable afterwards; C++11
// guarantees the initialisation is thread-safe).  Defaults are the measured best.
//   AURDF_ICP_SMALL        0: general kernel only; 1: icp_small_kernel (round 1); 2: icp_small2_kernel (default)
//   AURDF_ICP_SMALL_MINB   resident CTAs per SM the small-tile kernel is compiled for
//   AURDF_ICP_SMALL_SPLIT  0: every round of a tile uses the same lane split (icp_small_kernel only)
//   AURDF_ICP_STRICT_NT    tiles with at most this many masked targets use the strict pose fit (v2)
//   AURDF_ICP_GRID         0: no grid-pruned search (non-small tiles scan every target); 1 (default): icp_grid_kernel
//   AURDF_ICP_GRID_CS_NS   source points per tile above which the grid kernel runs as 8-CTA clusters
//   AURDF_ICP_GRID_SMEM_KB shared memory per CTA for a tile's sorted targets + cell table (one-CTA variant)
//   AURDF_ICP_GRID_CELL_SCALE cell size relative to the one-target-per-cell-by-volume rule
struct Tuning {
    int small, minb, split, strict_nt, grid, grid_cs_ns, grid_smem_kb;
    float grid_cell_scale;
};
static const Tuning &tuning() {
    static const Tuning t = [] {
        auto geti = [](const char *name, int dflt) {
            const char *e = getenv(name);
            return e ? atoi(e) : dflt;
        };
        Tuning v;
        v.small = geti("AURDF_ICP_SMALL", 2);
        v.minb = geti("AURDF_ICP_SMALL_MINB", 4);
        v.split = geti("AURDF_ICP_SMALL_SPLIT", 1);
        v.strict_nt = geti("AURDF_ICP_STRICT_NT", 16);
        v.grid = geti("AURDF_ICP_GRID", 1);
        v.grid_cs_ns = geti("AURDF_ICP_GRID_CS_NS", 1024);
        v.grid_smem_kb = geti("AURDF_ICP_GRID_SMEM_KB", 112);
        const char *cs = getenv("AURDF_ICP_GRID_CELL_SCALE");
        v.grid_cell_scale = cs ? (float)atof(cs) : 1.0f;
        if (!(v.grid_cell_scale > 0.05f && v.grid_cell_scale < 50.f)) v.grid_cell_scale = 1.0f;
        return v;
    }();
    return t;
}

// Residency the small-tile kernel is compiled for: 4 CTAs per SM (122 registers, no spills).  The launch is bound by
// its slowest tile and by instruction issue, not by occupancy: measured on wx200_5 / franka / allegro_hand, 4 beats
// 5 and 6 both device-resident and through the host path (profiles/r02_notes.md).
static int small_tile_residency() {
    const int forced = tuning().minb;
    return forced >= 4 && forced <= 6 ? forced : 4;
}

extern "C" int aurdf_icp_sweep_launches(void) { return tuning().small ? 5 : 4; }

// ---- optional live timing of the dominant kernel (icp_tiles_kernel) --------------------------
namespace {
struct EvPair { cudaEvent_t a, b; };
thread_local bool g_prof_on = false;
thread_local std::vector<EvPair> g_prof;
}  // namespace

extern "C" int aurdf_icp_profile_enable(int on) {
    g_prof_on = on != 0;
    return AURDF_OK;
}

extern "C" int aurdf_icp_profile_collect(double *total_ms, int32_t *n_launches) {
    double tot = 0.0;
    int n = 0;
    for (EvPair &e : g_prof) {
        float ms = 0.f;
        AURDF_CUDA_CHECK(cudaEventSynchronize(e.b));
        AURDF_CUDA_CHECK(cudaEventElapsedTime(&ms, e.a, e.b));
        tot += ms;
        ++n;
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    g_prof.clear();
    if (total_ms) *total_ms = tot;
    if (n_launches) *n_launches = n;
    return AURDF_OK;
}

extern "C" int aurdf_icp_sweep(const void *src_xyz, int pts_dtype, const int32_t *src_off, const void *tgt_xyz,
                               const int32_t *tgt_off, const int32_t *tile_frame, const void *box_xyz,
                               int box_dtype, const int32_t *box_off, const double *init_T, int32_t n_tiles,
                               int64_t total_src_points, int32_t max_src_per_tile, double box_scale, double max_corr_dist, int32_t max_iter,
                               double rel_fitness, double rel_rmse, int32_t ori_only, double *out_T,
                               double *out_world_xyz, int32_t *out_corr, double *out_fitness, double *out_rmse,
                               int32_t *out_iters, int32_t *out_ntgt, void *workspace, size_t workspace_bytes,
                               int64_t tgt_capacity, int32_t *status, aurdf_stream_t stream_) {
    aurdf::NvtxRange nvtx_range("aurdf_icp_sweep");
    cudaStream_t stream = (cudaStream_t)stream_;
    AURDF_REQUIRE(n_tiles >= 0, "aurdf_icp_sweep: n_tiles < 0");
    if (n_tiles == 0) return AURDF_OK;
    AURDF_REQUIRE(max_corr_dist > 0.0, "aurdf_icp_sweep: max_corr_dist must be > 0 (open3d raises)");
    AURDF_REQUIRE(max_iter >= 0, "aurdf_icp_sweep: max_iter < 0");
    AURDF_REQUIRE(pts_dtype == AURDF_F32 || pts_dtype == AURDF_F64, "aurdf_icp_sweep: bad pts_dtype");
    AURDF_REQUIRE(box_dtype == AURDF_F32 || box_dtype == AURDF_F64, "aurdf_icp_sweep: bad box_dtype");
    AURDF_REQUIRE(src_off && tgt_off && tile_frame && init_T, "aurdf_icp_sweep: NULL input");
    AURDF_REQUIRE(box_xyz == nullptr || box_off != nullptr, "aurdf_icp_sweep: box_xyz without box_off");
    AURDF_REQUIRE(out_T && out_world_xyz && out_corr && out_fitness && out_rmse && out_iters && out_ntgt,
                  "aurdf_icp_sweep: NULL output");
    AURDF_REQUIRE(workspace && tgt_capacity >= 0, "aurdf_icp_sweep: NULL workspace");
    AURDF_REQUIRE(((uintptr_t)workspace & 255) == 0, "aurdf_icp_sweep: workspace must be 256-byte aligned");

    AURDF_REQUIRE(total_src_points >= 0, "aurdf_icp_sweep: total_src_points < 0");
    const WsLayout L = make_layout(n_tiles, total_src_points, tgt_capacity);
    if (workspace_bytes < L.total) {
        set_error("aurdf_icp_sweep: workspace_bytes %zu < %zu", workspace_bytes, L.total);
        return AURDF_EWORKSPACE;
    }
    char *ws = (char *)workspace;
    double *box = (double *)(ws + L.box);
    int *cnt = (int *)(ws + L.cnt);
    long long *toff = (long long *)(ws + L.toff);
    int *status_int = (int *)(ws + L.status);
    double *qx = (double *)(ws + L.qx), *qy = (double *)(ws + L.qy), *qz = (double *)(ws + L.qz);
    int *qi = (int *)(ws + L.qi);
    double *pspill = (double *)(ws + L.pspill);

    // Large tiles (caller's hint) run as clusters of kClusterCtas CTAs, each owning a slice of the
    // source points.  Source points live in shared memory up to p_cap per CTA (bounded by the hint and
    // kPSmemMax); larger slices keep them in the workspace spill area (sized by total_src_points).
    const bool use_cluster = max_src_per_tile > 384;
    int p_cap = max_src_per_tile > 0 ? max_src_per_tile : 256;
    if (use_cluster) p_cap = (p_cap + kClusterCtas - 1) / kClusterCtas;
    if (p_cap > kPSmemMax) p_cap = kPSmemMax;
    p_cap = (p_cap + 1) & ~1;
    const size_t smem = (size_t)3 * kQChunk * sizeof(double) + (size_t)kQChunk * sizeof(float4) +
                        (size_t)3 * p_cap * sizeof(double) + (size_t)p_cap * sizeof(int);
    if (smem > 32 * 1024) {   // static + dynamic beyond the default 48 KiB needs the opt-in
        if (use_cluster)
            AURDF_CUDA_CHECK(cudaFuncSetAttribute(icp_tiles_kernel<kClusterCtas>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else
            AURDF_CUDA_CHECK(cudaFuncSetAttribute(icp_tiles_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }

    IcpParams P;
    P.src = src_xyz; P.pts_dtype = pts_dtype; P.src_off = src_off; P.init_T = init_T;
    P.r2 = max_corr_dist * max_corr_dist; P.max_iter = max_iter; P.rel_fit = rel_fitness; P.rel_rmse = rel_rmse;
    P.ori_only = ori_only;
    P.qx = qx; P.qy = qy; P.qz = qz; P.qi = qi; P.toff = toff; P.cnt = cnt; P.status_int = status_int;
    P.pspill = pspill; P.p_cap = p_cap;
    P.out_T = out_T; P.out_world = out_world_xyz; P.out_corr = out_corr; P.out_fit = out_fitness;
    P.out_rmse = out_rmse; P.out_iters = out_iters; P.out_ntgt = out_ntgt;
    P.dbg_clock = g_dbg_clock;
    const Tuning &tn = tuning();
    P.small_on = tn.small != 0;
    P.split_tail = tn.split;
    P.small_ns = tn.small == 2 ? kS2Ns : kSmNs;
    P.small_nt = kSmNt32;
    P.n_tiles = n_tiles;
    P.queue = status_int + 3;
    P.queue2 = status_int + 4;
    P.strict_nt = tn.strict_nt;
    // grid-pruned search for the tiles the small-tile kernel does not take: one CTA per tile, or an 8-CTA
    // cluster when the caller announces tiles of more than grid_cs_ns source points
    P.grid_on = tn.grid != 0;
    P.grid_cs = max_src_per_tile > tn.grid_cs_ns ? kClusterCtas : 1;
    P.grid_pcap = ((max_src_per_tile > 0 ? max_src_per_tile : 256) + P.grid_cs - 1) / P.grid_cs;
    if (P.grid_pcap > kGridPcapMax) P.grid_pcap = kGridPcapMax;
    P.grid_pcap = (P.grid_pcap + 3) & ~3;
    // one CTA (512 threads, the whole register file) per tile and SM: 112 KB for the grid (n_t <= ~5700); cluster variant (few, large
    // tiles): whatever one CTA per SM leaves after the source-point state and the static arrays
    P.grid_smem_bytes = P.grid_cs == 1 ? tn.grid_smem_kb * 1024 : (227 - 4) * 1024 - P.grid_pcap * 48 - 2048 * 4;
    if (P.grid_smem_bytes < 0) P.grid_smem_bytes = 0;
    P.grid_cell_scale = tn.grid_cell_scale;
    P.gs = (float4 *)(ws + L.gs); P.gends = (int *)(ws + L.gends); P.gpar = (float *)(ws + L.gpar);
    P.glist = (int *)(ws + L.glist);
    box_count_kernel<<<n_tiles, kIcpThreads, 0, stream>>>(box_xyz, box_dtype, box_off, tgt_xyz, pts_dtype, tgt_off,
                                                          tile_frame, box_scale, box, cnt);
    tile_scan_kernel<<<1, 1024, 0, stream>>>(cnt, n_tiles, (long long)tgt_capacity, toff, status_int, status, P);
    mask_fill_kernel<<<n_tiles, kIcpThreads, 0, stream>>>(tgt_xyz, pts_dtype, tgt_off, tile_frame, box, toff,
                                                          status_int, qx, qy, qz, qi);
    EvPair ev{nullptr, nullptr};
    if (g_prof_on) {
        AURDF_CUDA_CHECK(cudaEventCreate(&ev.a));
        AURDF_CUDA_CHECK(cudaEventCreate(&ev.b));
        AURDF_CUDA_CHECK(cudaEventRecord(ev.a, stream));
    }
    // small tiles first (one CTA each, all resident at once); the general kernel's CTAs for those
    // tiles exit immediately, and vice versa
    if (P.small_on) {
        const int rc = tn.small == 2 ? launch_icp_small2(P, n_tiles, small_tile_residency(), stream) : launch_icp_small(P, n_tiles, small_tile_residency(), stream);
        if (rc != AURDF_OK) return rc;
    }
    if (P.grid_on) {
        const int rc = launch_icp_grid(P, n_tiles, stream);
        if (rc != AURDF_OK) return rc;
    }
    // the brute-force general kernel is only needed without the grid kernel, or beside its cluster variant
    // (rank-deficient large tiles, slices beyond the shared-memory bound)
    if (P.grid_on && P.grid_cs == 1) {
    } else if (use_cluster) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)n_tiles * kClusterCtas);
        cfg.blockDim = dim3(kIcpThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kClusterCtas;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        AURDF_CUDA_CHECK(cudaLaunchKernelEx(&cfg, icp_tiles_kernel<kClusterCtas>, P));
    } else {
        icp_tiles_kernel<1><<<n_tiles, kIcpThreads, smem, stream>>>(P);
    }
    if (g_prof_on) {
        AURDF_CUDA_CHECK(cudaEventRecord(ev.b, stream));
        g_prof.push_back(ev);
    }
    AURDF_CUDA_CHECK(cudaGetLastError());
    return AURDF_OK;
}
