// icp_small2.cu -- second generation of the small-tile ICP kernel (n_s <= 256 source points,
// n_t <= 760 masked targets: every tile of wx200 / franka / allegro_hand).
//
// Semantics: open3d RegistrationICP point-to-point inside masked_icp (AutoURDF
// PointCloud/cluster_icp.py:118-191, call at :157-159), as icp_tiles_kernel / icp_small_kernel.
// What changed against icp_small.cu, and the measurement behind each change (profiles/r02_notes.md):
//   * persistent CTAs take tiles from a queue (no second wave of 12 tiles behind 888 slots);
//   * every warp owns a fixed set of "home" source points (dealt round-robin, so a 20-point tile still
//     uses all four warps) and does the whole per-point work of an iteration on them without a block
//     barrier: move, nearest neighbour, moment sums.  Two block barriers per iteration instead of three;
//   * nearest-neighbour cache: a point remembers where it was scanned last (anchor), its winner and a
//     lower bound rho on the distance from the anchor to every OTHER target.  As long as
//     d(p, winner) + |p - anchor| < rho the winner is provably still the float64 argmin (triangle
//     inequality, every rounding over-covered) and the scan is skipped -- 60 % of the point-iterations
//     of wx200_5.  The points that do need a scan are compacted per warp and share the warp's 32 lanes
//     (S lanes per point), so a warp with 8 misses scans a quarter as long;
//   * the pose fit works on the unnormalised covariance n*S_ba - S_b S_a^T (no division before the
//     Newton iteration; 1/n is only needed for the translation and is computed beside it);
//   * tiles whose box holds a handful of targets (rank-deficient fits) fit their poses in "strict" mode:
//     the CPU reference's arithmetic operation for operation (icp_common.cuh, namespace strict), so
//     they too are bit-comparable instead of being a different member of a one-parameter family.
// The float32 scan itself (packed f32x2 arithmetic, keyed min tracking, certified against float64) is
// the one of icp_small.cu with two independent key pairs per lane.
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "icp_common.cuh"

namespace aurdf {

namespace {
constexpr int kNT = 128;
constexpr int kWarps = kNT / 32;
constexpr uint32_t kIdxMask2 = 0x3FFu;   // low mantissa bits of a key hold the target index (< 1024)
static_assert(2 * kSmPairs <= 1024, "index field");

// Shared memory of a CTA: the float32 target pairs, the per-warp scan lists, and a pool that every tile
// lays out for itself -- 52 bytes of state per home slot (128 slots per round of source points), then the
// float64 targets if they still fit (otherwise they are read from the compacted arrays in L2).
// 6 CTAs per SM: 6 x (36096 + 1360 static + 1024 reserved) <= 233472.
constexpr size_t kSlotBytes = sizeof(float4) + 4 * sizeof(double) + sizeof(int);
constexpr size_t kPoolBytes = 21760;
constexpr size_t kSmall2SmemBytes = (size_t)kSmPairs * 2 * sizeof(float4) + (size_t)kWarps * 32 * sizeof(float4) + kPoolBytes;
static_assert((size_t)kS2Ns * kSlotBytes <= kPoolBytes, "the per-slot state of the largest small tile must fit");

// home slot t = 128 r + 32 w + l  <->  source point 128 r + 4 l + w (round r, warp w, lane l)
__device__ __forceinline__ int slot_to_point(int t) { return (t & ~127) + 4 * (t & 31) + ((t >> 5) & 3); }
__device__ __forceinline__ int point_to_slot(int i) { return (i & ~127) + 32 * (i & 3) + ((i & 127) >> 2); }
}  // namespace

template <int MINB, bool DBG>
__global__ void __launch_bounds__(kNT, MINB)
icp_small2_kernel(const IcpParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *sxy = reinterpret_cast<float4 *>(smem_raw);             // (-x0, -x1, -y0, -y1) of a target pair
    float4 *sz = sxy + kSmPairs;                                    // (-z0, -z1, bits: index of target 0, of target 1)
    float4 *slist = sz + kSmPairs;                                  // per warp: points to scan (x, y, z, slot bits)
    // the rest of the pool is laid out per tile: per-slot state (128 slots per round), then -- if they
    // still fit -- the float64 targets
    unsigned char *pool = reinterpret_cast<unsigned char *>(slist + kWarps * 32);

    __shared__ double s_part[kWarps][4][6];   // per-warp moment products M[0..3][0..5] of the current pass
    __shared__ double s_tot[16];              // their totals, M[r][c] at 4 r + c (fit warp only)
    __shared__ double s_U[16];     // current update (row-major 4x4)
    __shared__ double s_T[16];     // accumulated pose
    __shared__ double s_prev[2];   // fitness, rmse of the previous pass
    __shared__ double s_warm[18];  // singular vectors of the previous Jacobi fit (fallback only)
    __shared__ int s_stop, s_tile, s_strict;
    __shared__ float s_amax[kWarps];
    __shared__ __align__(8) uint64_t s_bar;

    if (p.status_int[0]) return;   // compacted-target capacity exceeded: leave outputs untouched

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
    }
    uint32_t bar_phase = 0;

    // debug hook: (clock, id) pairs of thread 0 while it works on tile 0 (scripts/tile_latency.py)
    int dbg_n = 0;
    bool dbg_on = false;
    auto stamp = [&](int id) {
        if constexpr (!DBG) return;
        if (dbg_on && tid == 0 && dbg_n < 2040) {
            p.dbg_clock[2 * dbg_n] = clock64();
            p.dbg_clock[2 * dbg_n + 1] = id;
            ++dbg_n;
        }
    };

    for (;;) {
        __syncthreads();   // the previous tile is completely done (shared memory is reused)
        if (tid == 0) s_tile = atomicAdd(p.queue, 1);
        __syncthreads();
        const int b = s_tile;
        if (b >= p.n_tiles) break;
        const int s0 = p.src_off[b];
        const int ns = p.src_off[b + 1] - s0;
        const int nt = p.cnt[b];
        if (!(ns <= p.small_ns && nt <= p.small_nt)) continue;   // the general kernel owns this tile
        if constexpr (DBG) dbg_on = b == 0;
        const long long q0 = p.toff[b];
        const double *gqx = p.qx + q0, *gqy = p.qy + q0, *gqz = p.qz + q0;
        const int rounds_ = (ns + kNT - 1) / kNT, nslot = rounds_ * kNT, nte = (nt + 1) & ~1;
        float4 *sanc = reinterpret_cast<float4 *>(pool);            // cache: anchor (float32, about the origin), rho
        double *spx = reinterpret_cast<double *>(sanc + nslot);     // current source points, by home slot
        double *spy = spx + nslot;
        double *spz = spy + nslot;
        double *sbd = spz + nslot;                                  // exact squared distance to the nearest target
        int *scj = reinterpret_cast<int *>(sbd + nslot);            // nearest target (compacted index) or -1
        double *sqx = reinterpret_cast<double *>(scj + nslot);      // float64 targets, if the pool has room for them
        double *sqy = sqx + nte;
        double *sqz = sqy + nte;
        const bool q64s = (size_t)nslot * kSlotBytes + (size_t)nte * 24 <= kPoolBytes;
        const double *qxp = q64s ? sqx : gqx, *qyp = q64s ? sqy : gqy, *qzp = q64s ? sqz : gqz;
        const int rounds = rounds_;

        if (tid < 16) {
            s_T[tid] = p.init_T[16 * (size_t)b + tid];
            s_U[tid] = (tid % 5 == 0) ? 1.0 : 0.0;
        }
        if (tid == 0) {
            s_stop = 0;
            s_strict = nt <= p.strict_nt ? 1 : 0;
        }
        __syncthreads();

        // float64 targets: three bulk copies (TMA engine) on one mbarrier
        if (q64s && nt > 0) {
            const uint32_t bytes = (uint32_t)(((nt + 1) & ~1) * sizeof(double));
            if (tid == 0) {
                fence_proxy_async();
                mbar_arrive_expect_tx(&s_bar, 3 * bytes);
                bulk_g2s(sqx, gqx, bytes, &s_bar);
                bulk_g2s(sqy, gqy, bytes, &s_bar);
                bulk_g2s(sqz, gqz, bytes, &s_bar);
            }
        }

        // P <- T0 * S into the home slots (overlaps the bulk copies); empty cache
        {
            const bool aff0 = s_T[12] == 0.0 && s_T[13] == 0.0 && s_T[14] == 0.0 && s_T[15] == 1.0;
            for (int t = tid; t < rounds * kNT; t += kNT) {
                const int i = slot_to_point(t);
                double x = 0.0, y = 0.0, z = 0.0;
                if (i < ns) {
                    const size_t e = 3 * (size_t)(s0 + i);
                    x = ld_coord(p.src, p.pts_dtype, e); y = ld_coord(p.src, p.pts_dtype, e + 1);
                    z = ld_coord(p.src, p.pts_dtype, e + 2);
                    transform_point(s_T, aff0, x, y, z);
                }
                spx[t] = x; spy[t] = y; spz[t] = z;
                sbd[t] = 0.0;
                scj[t] = -1;
                sanc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (q64s && nt > 0) {
            mbar_wait(&s_bar, bar_phase);
            bar_phase ^= 1;
        }

        // moments are accumulated about the tile's first target point (kills cancellation); the
        // float32 filter works in the same frame
        const double ox = nt > 0 ? qxp[0] : 0.0, oy = nt > 0 ? qyp[0] : 0.0, oz = nt > 0 ? qzp[0] : 0.0;

        // most lanes a point may get: the scan reads up to 8 S pairs past the end (software pipelining)
        const int npairs = (nt + 1) >> 1;
        int S_cap = 1;
        while (S_cap < 32 && npairs + 16 * S_cap <= kSmPairs) S_cap *= 2;

        // float32 copies of the targets, negated (the scan adds), two per entry; the odd tail and the
        // read-ahead padding are points no source can match.  aq = largest coordinate magnitude, scales
        // the error bound.
        float aq = 0.f;
        {
            float amax = 0.f;
            const int nfill = min(kSmPairs, npairs + 8 * S_cap);
            for (int jj = tid; jj < nfill; jj += kNT) {
                const int j0 = 2 * jj, j1 = j0 + 1;
                float ax = 1e18f, ay = 0.f, az = 0.f, bx = 1e18f, by = 0.f, bz = 0.f;
                if (j0 < nt) {
                    ax = (float)(qxp[j0] - ox); ay = (float)(qyp[j0] - oy); az = (float)(qzp[j0] - oz);
                    amax = fmaxf(amax, fmaxf(fabsf(ax), fmaxf(fabsf(ay), fabsf(az))));
                }
                if (j1 < nt) {
                    bx = (float)(qxp[j1] - ox); by = (float)(qyp[j1] - oy); bz = (float)(qzp[j1] - oz);
                    amax = fmaxf(amax, fmaxf(fabsf(bx), fmaxf(fabsf(by), fabsf(bz))));
                }
                sxy[jj] = make_float4(-ax, -bx, -ay, -by);
                sz[jj] = make_float4(-az, -bz, __uint_as_float((uint32_t)j0), __uint_as_float((uint32_t)j1));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            if (lane == 0) s_amax[warp] = amax;
            __syncthreads();
#pragma unroll
            for (int w = 0; w < kWarps; ++w) aq = fmaxf(aq, s_amax[w]);
        }

        // exact squared distance in the reference's operation order (nanoflann L2 adaptor, no FMA)
        auto exact_d2 = [&](double x, double y, double z, int j) {
            const double dx = __dsub_rn(x, qxp[j]), dy = __dsub_rn(y, qyp[j]), dz = __dsub_rn(z, qzp[j]);
            return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        };

        // ---- phase A (warp-local): move the home points, find every point's nearest target ----
        auto pass = [&](bool apply) {
            for (int r = 0; r < rounds; ++r) {
                const int t = r * kNT + tid;
                const bool active = r * kNT + 4 * lane + warp < ns;
                // everything the cache test needs is loaded before the move (independent of it): the anchor,
                // the previous winner and its coordinates
                const float4 an = sanc[t];
                const int j1c = max(scj[t], 0);
                double qx1 = 0.0, qy1 = 0.0, qz1 = 0.0;
                if (nt > 0) { qx1 = qxp[j1c]; qy1 = qyp[j1c]; qz1 = qzp[j1c]; }   // tile-uniform branch
                double x = 0.0, y = 0.0, z = 0.0;
                if (active) {
                    x = spx[t]; y = spy[t]; z = spz[t];
                    if (apply) {
                        transform_point(s_U, true, x, y, z);
                        spx[t] = x; spy[t] = y; spz[t] = z;
                    }
                }
                if (r == 0) stamp(1);   // P update done
                const float fx = (float)(x - ox), fy = (float)(y - oy), fz = (float)(z - oz);
                bool miss = active && nt > 0;
                {
                    // cache test: the winner of the last scan is still the float64 argmin if
                    // d(p, winner) + |p - anchor| < rho (rho: lower bound on the distance from the anchor
                    // to every other target).  float32 arithmetic, every rounding over-covered 100-fold.
                    const double dx = __dsub_rn(x, qx1), dy = __dsub_rn(y, qy1), dz = __dsub_rn(z, qz1);
                    const double D1 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    const float ex = fx - an.x, ey = fy - an.y, ez = fz - an.z;
                    const float del = sqrtf(fmaf(ez, ez, fmaf(ey, ey, ex * ex)));
                    const float amag = fmaxf(aq, fmaxf(fabsf(fx), fmaxf(fabsf(fy), fabsf(fz))));
                    const float lhs = (sqrtf((float)D1) + del) * 1.0001f + 1e-6f * (amag + del);
                    if (miss && an.w > 0.f && lhs < an.w) {
                        miss = false;
                        sbd[t] = D1;
                    }
                }
                if (r == 0) stamp(2);   // cache test done
                const unsigned mb = __ballot_sync(0xffffffffu, miss);
                const int n = __popc(mb);
                if (n == 0) continue;   // warp-uniform
                if (miss) slist[warp * 32 + __popc(mb & lt_mask)] = make_float4(fx, fy, fz, __int_as_float(t));
                __syncwarp();
                int ls = 0;
                while ((2 << ls) <= S_cap && n * (2 << ls) <= 32) ++ls;
                const int S = 1 << ls;                 // lanes per point: the warp's misses share its 32 lanes
                const int sub = lane & (S - 1);
                const int e = lane >> ls;
                const bool valid = e < n;
                const float4 ent = slist[warp * 32 + (valid ? e : 0)];
                const int te = __float_as_int(ent.w);
                // float32 pre-filter.  key = (bits of the float32 squared distance, low 10 bits replaced
                // by the target index): unsigned order = distance order up to 2^-13 relative, ties and
                // near-ties fall to the exact scan below through the certificate.
                uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;
                if (valid) {
                    const float2 fx2 = make_float2(ent.x, ent.x), fy2 = make_float2(ent.y, ent.y), fz2 = make_float2(ent.z, ent.z);
                    uint32_t n1 = 0xFFFFFFFFu, n2 = 0xFFFFFFFFu;   // second, independent key pair (odd pairs)
                    auto pair = [&](const float4 qxy, const float4 qz, uint32_t &k1, uint32_t &k2) {
                        const float2 dx = __fadd2_rn(fx2, make_float2(qxy.x, qxy.y));
                        const float2 dy = __fadd2_rn(fy2, make_float2(qxy.z, qxy.w));
                        const float2 dz = __fadd2_rn(fz2, make_float2(qz.x, qz.y));
                        float2 d = __fmul2_rn(dx, dx);
                        d = __ffma2_rn(dy, dy, d);
                        d = __ffma2_rn(dz, dz, d);
                        const uint32_t a = (__float_as_uint(d.x) & ~kIdxMask2) | __float_as_uint(qz.z);
                        const uint32_t c = (__float_as_uint(d.y) & ~kIdxMask2) | __float_as_uint(qz.w);
                        const uint32_t lo = min(a, c), hi = max(a, c);
                        k2 = __vimin3_u32(k2, hi, max(k1, lo));
                        k1 = min(k1, lo);
                    };
                    // SS > 0: compile-time lane stride (addresses fold into immediates); 0: runtime S
                    auto scan = [&](auto stride_c) {
                        constexpr int SS = decltype(stride_c)::value;
                        const int st = SS ? SS : S;
                        const int trips2 = ((npairs + st - 1) / st + 1) >> 1;   // trips of two pairs per lane
                        const float4 *pxy = sxy + sub;
                        float4 a0 = pxy[0], a1 = pxy[st];
                        float4 b0 = pxy[kSmPairs], b1 = pxy[kSmPairs + st];   // sz = sxy + kSmPairs
#pragma unroll 2
                        for (int q = 0; q < trips2; ++q) {
                            pxy += 2 * st;
                            const float4 na = pxy[0], nb = pxy[st];
                            const float4 c0 = pxy[kSmPairs], c1 = pxy[kSmPairs + st];
                            pair(a0, b0, m1, m2);
                            pair(a1, b1, n1, n2);
                            a0 = na; a1 = nb; b0 = c0; b1 = c1;
                        }
                    };
                    if (S == 1) scan(std::integral_constant<int, 1>{});
                    else if (S == 2) scan(std::integral_constant<int, 2>{});
                    else if (S == 4) scan(std::integral_constant<int, 4>{});
                    else scan(std::integral_constant<int, 0>{});
                    m2 = __vimin3_u32(m2, n2, max(m1, n1));
                    m1 = min(m1, n1);
                }
                if (r == 0) stamp(3);   // float32 scan done
                for (int o = S >> 1; o > 0; o >>= 1) {
                    const uint32_t om1 = __shfl_xor_sync(0xffffffffu, m1, o), om2 = __shfl_xor_sync(0xffffffffu, m2, o);
                    m2 = __vimin3_u32(m2, om2, max(m1, om1));
                    m1 = min(m1, om1);
                }
                double x2 = 0.0, y2 = 0.0, z2 = 0.0;
                if (valid) { x2 = spx[te]; y2 = spy[te]; z2 = spz[te]; }
                bool need_exact = valid;
                const int j1 = (int)(m1 & kIdxMask2);
                if (valid && m1 != 0xFFFFFFFFu && j1 < nt) {
                    // true float32 distances: best <= m1hi, every other target >= m2lo.  The float32
                    // distance itself is within tau/4 of the real one (icp_sweep.cu), tau taken at the
                    // larger value.
                    const float m1hi = __uint_as_float(m1 | kIdxMask2);
                    const float m2lo = __uint_as_float(m2 & ~kIdxMask2), m2hi = __uint_as_float(m2 | kIdxMask2);
                    const float u = 5.9604645e-8f;
                    const float amag = fmaxf(aq, fmaxf(fabsf(ent.x), fmaxf(fabsf(ent.y), fabsf(ent.z))));
                    const float dl = 4.f * u * amag;
                    const float tau = 16.f * (dl * sqrtf(m2hi) * 1.001f + dl * dl + u * m2hi);
                    if (nt == 1 || (m2lo - m1hi > 2.f * tau && m2hi < INFINITY)) {
                        need_exact = false;
                        if (sub == 0) {
                            scj[te] = j1;
                            sbd[te] = exact_d2(x2, y2, z2, j1);
                            // every other target is at least sqrt(m2lo - tau) away from this position
                            // (d - err(d) grows with d once d > 64 dl^2, so the bound taken at the second
                            // best covers the farther ones)
                            float rho = 0.f;
                            if (nt == 1) rho = 1e30f;
                            else if (m2lo > 128.f * dl * dl && m2lo - tau > 0.f) rho = sqrtf(m2lo - tau) * 0.9999f;
                            sanc[te] = make_float4(ent.x, ent.y, ent.z, rho);
                        }
                    }
                }
                if (r == 0) stamp(4);   // merge + certificate + exact distance done
                // exact rescan of the uncertified points of this warp (~1e-3 of them; duplicates always):
                // float64, the reference's operation order, strict '<' in ascending index order
                if (__builtin_expect(__any_sync(0xffffffffu, need_exact) != 0, 0)) {
                    double bd = INFINITY;
                    int bj = -1;
                    if (need_exact) {
                        for (int j = sub; j < nt; j += S) {
                            const double d = exact_d2(x2, y2, z2, j);
                            if (d < bd) { bd = d; bj = j; }
                        }
                    }
                    for (int o = S >> 1; o > 0; o >>= 1) {   // (d, j) lexicographic min across the S lanes
                        const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
                        if (oj >= 0 && (od < bd || (od == bd && oj < bj) || bj < 0)) { bd = od; bj = oj; }
                    }
                    if (need_exact && sub == 0) {
                        scj[te] = bj;
                        sbd[te] = bd;
                        sanc[te] = make_float4(ent.x, ent.y, ent.z, 0.f);   // no bound: scan again next time
                    }
                }
                __syncwarp();   // list entries consumed, results visible to the warp
            }
        };

        // ---- phase B (warp-local): moment sums of the warp's home points on the FP64 tensor cores ----
        // The 16 moments + sum d^2 are one small product M = U^T W over the matched pairs, with
        // U = (1, bx, by, bz) (zero row when the point has no match) and W = (1, ax, ay, az, d^2);
        // a = source point, b = matched target, both about the origin o.  M[0][0] is the inlier count,
        // M[0][1..3] = sum a, M[1..3][0] = sum b, M[1..3][1..3] = sum b a^T, M[0][4] = sum d^2.
        // One mma.m8n8k4.f64 adds four points: A[row][k] = U_row(point k), B[k][col] = W_col(point k).
        const int gid = lane >> 2, tig = lane & 3;
        auto reduce = [&]() {
            const double *pu = gid == 1 ? qxp : (gid == 2 ? qyp : qzp);
            const double *pw = gid == 1 ? spx : (gid == 2 ? spy : (gid == 3 ? spz : sbd));
            const double oc = gid == 1 ? ox : (gid == 2 ? oy : (gid == 3 ? oz : 0.0));
            // operand = ((value - oc) * mul + add1) * matched: row / column 0 is the constant 1, rows >= 4 and
            // columns >= 5 are 0, an unmatched point contributes nothing.  Pure arithmetic: lanes of one warp
            // must not take different branches around the mma.
            const double mul_u = (gid >= 1 && gid < 4) ? 1.0 : 0.0, mul_w = (gid >= 1 && gid < 5) ? 1.0 : 0.0;
            const double add1 = gid == 0 ? 1.0 : 0.0;
            double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            for (int r = 0; r < rounds; ++r) {
                const int base = r * kNT + 32 * warp + tig;
                // all loads of the round first (the slots past the warp's last home point hold j = -1), then
                // the eight products back to back: the load latencies are paid once, not once per group
                int jj[8];
                double okf[8], u[8], w[8];
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    jj[g] = scj[base + 4 * g];
                    okf[g] = sbd[base + 4 * g];
                    w[g] = pw[base + 4 * g];
                }
                stamp(30);   // reduce: first loads issued
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const bool ok = jj[g] >= 0 && okf[g] < p.r2;
                    u[g] = pu[max(jj[g], 0)];
                    okf[g] = ok ? 1.0 : 0.0;
                }
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const double uu = fma(u[g] - oc, mul_u, add1) * okf[g];
                    const double ww = fma(w[g] - oc, mul_w, add1) * okf[g];   // column 4 = d^2 (oc = 0)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(acc[g & 1][0]), "+d"(acc[g & 1][1]) : "d"(uu), "d"(ww));
                }
            }
            stamp(31);   // reduce: products done
            const double d0 = acc[0][0] + acc[1][0], d1 = acc[0][1] + acc[1][1];
            // this lane holds M[gid][2 tig] and M[gid][2 tig + 1]
            if (gid < 4 && tig < 3) {
                s_part[warp][gid][2 * tig] = d0;
                s_part[warp][gid][2 * tig + 1] = d1;
            }
        };

        // ---- pose fit (warp 0) ----
        bool have_warm = false;
        // fast path, lane 0: Newton on SO(3) from the unnormalised covariance; false = not certifiable
        auto fit_fast = [&]() {
            const double *t = s_tot;   // t[4 r + c] = sum u_r w_c
            double Um[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
            if (t[0] > 0.0) {
                const double n = t[0];
                double sigma[3][3], R[3][3];
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) sigma[r][cc] = n * t[4 * (r + 1) + cc + 1] - t[4 * (r + 1)] * t[cc + 1];
                stamp(20);   // totals loaded, covariance formed
                {
                    if (__builtin_expect(!kabsch_rotation_newton4(sigma, R), 0)) {   // reflection / rank-deficient / large step: Jacobi SVD, warm-started
                        // (copies: the out-of-line call takes addresses, and sigma / R must stay in registers
                        // on the fast path)
                        double sg2[3][3], R2[3][3];
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
                    stamp(21);   // rotation fitted
                    const double inv = rcp_raw2(n);
                    const double ma[3] = {t[1] * inv, t[2] * inv, t[3] * inv};
                    const double mb[3] = {t[4] * inv, t[8] * inv, t[12] * inv};
                    const double mua[3] = {ma[0] + ox, ma[1] + oy, ma[2] + oz};
                    const double mub[3] = {mb[0] + ox, mb[1] + oy, mb[2] + oz};
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        Um[4 * r + 0] = R[r][0]; Um[4 * r + 1] = R[r][1]; Um[4 * r + 2] = R[r][2];
                        Um[4 * r + 3] = mub[r] - (R[r][0] * mua[0] + R[r][1] * mua[1] + R[r][2] * mua[2]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 12; ++k) s_U[k] = Um[k];
        };
        // strict path, whole warp 0: the CPU reference's two-pass sums in ascending source index, one
        // accumulator per lane, then its Jacobi SVD on lane 0 (icp_common.cuh, namespace strict)
        auto fit_strict = [&]() {
            // Sums over the matched points in ascending source index.  Sixteen points at a time: lane l stages
            // the addends of point c0 + l in shared memory (one padded row per accumulator), then accumulator
            // lane k adds row k in order -- its loads are independent, only the adds form a chain.
            double *stage = reinterpret_cast<double *>(slist);   // [9][17]; the other warps wait at barrier B
            auto ordered_sum = [&](int nacc, auto addends, int &count) {
                double acc = 0.0;
                count = 0;
                for (int c0 = 0; c0 < ns; c0 += 16) {
                    const int i = c0 + (lane & 15);
                    int t = 0, j = 0;
                    bool ok = false;
                    if (lane < 16 && i < ns) {
                        t = point_to_slot(i);
                        j = scj[t];
                        ok = j >= 0 && sbd[t] < p.r2;
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, ok);
                    count += __popc(m);
                    if (ok) addends(t, j, stage + lane);
                    __syncwarp();
                    if (lane < nacc) {
                        double v[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) v[q] = stage[lane * 17 + q];
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            if ((m >> q) & 1u) acc = strict::add(acc, v[q]);
                    }
                    __syncwarp();
                }
                return acc;
            };
            int c = 0;
            const double sum1 = ordered_sum(6, [&](int t, int j, double *row) {
                row[0] = spx[t]; row[17] = spy[t]; row[34] = spz[t];
                row[51] = qxp[j]; row[68] = qyp[j]; row[85] = qzp[j]; }, c);
            double Um[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
            if (c > 0) {   // warp-uniform
                const double one_over_n = strict::dvd(1.0, (double)c);
                const double mean = strict::mul(sum1, one_over_n);   // lanes 0-2: source mean, 3-5: target mean
                double ms[3], md[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    ms[k] = __shfl_sync(0xffffffffu, mean, k);
                    md[k] = __shfl_sync(0xffffffffu, mean, 3 + k);
                }
                int c2 = 0;
                double sg = ordered_sum(9, [&](int t, int j, double *row) {
                    const double a[3] = {strict::sub(spx[t], ms[0]), strict::sub(spy[t], ms[1]), strict::sub(spz[t], ms[2])};
                    const double bb[3] = {strict::sub(qxp[j], md[0]), strict::sub(qyp[j], md[1]), strict::sub(qzp[j], md[2])};
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) row[17 * (3 * r + cc)] = strict::mul(bb[r], a[cc]); }, c2);
                sg = strict::mul(sg, one_over_n);            // lanes 0-8: covariance entry (lane / 3, lane % 3)
                double sigma[3][3];
#pragma unroll
                for (int k = 0; k < 9; ++k) sigma[k / 3][k % 3] = __shfl_sync(0xffffffffu, sg, k);
                stamp(22);   // strict: ordered sums done
                if (lane == 0) strict::pose_from_sigma(sigma, ms, md, Um);
                stamp(23);   // strict: Jacobi SVD + pose done
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

        // The serial roles rotate with the tile: the warps of the CTAs resident on one SM map onto its four
        // schedulers by warp index, so a fixed "warp 0 fits the pose" piles every resident tile's serial section on
        // one scheduler (measured: 2.1x the instructions of the average scheduler, profiles/r02a_kernels.json).
        const int fit_w = b & 3, conv_w = (b + 1) & 3, comp_w = (b + 3) & 3;
        // it = -1 is open3d's initial correspondence pass (no update applied); one copy of every phase
        // keeps the loop body small enough for the instruction caches.
        int iters = 0;
#pragma unroll 1
        for (int it = -1; it < p.max_iter; ++it) {
            const bool apply = it >= 0;
            stamp(0);   // iteration start
            if (apply && warp == comp_w) compose_pose();   // uses s_U of this iteration; its next write is after barrier A
            pass(apply);
            stamp(5);   // pass done
            reduce();
            stamp(7);   // moment sums done
            __syncthreads();   // barrier A: partial sums visible
            stamp(8);
            if (warp == fit_w) {
                // speculative: the fit for iteration it+1 runs while the next warp decides whether to stop
                if (it + 1 < p.max_iter) {
                    if (lane < 16) {
                        double v = s_part[0][lane >> 2][lane & 3];
#pragma unroll
                        for (int w = 1; w < kWarps; ++w) v += s_part[w][lane >> 2][lane & 3];
                        s_tot[lane] = v;
                    }
                    __syncwarp();
                    if (__builtin_expect(s_strict != 0, 0)) fit_strict();   // rare: kept off the hot instruction stream
                    else if (lane == 0) fit_fast();
                }
            } else if (warp == conv_w) {
                if (lane == 0) {
                    double cnt = s_part[0][0][0], d2 = s_part[0][0][4];
#pragma unroll
                    for (int w = 1; w < kWarps; ++w) { cnt += s_part[w][0][0]; d2 += s_part[w][0][4]; }
                    const int c = (int)cnt;
                    const double fit = c > 0 ? (double)c / (double)ns : 0.0;
                    const double rmse = c > 0 ? sqrt(d2 / (double)c) : 0.0;
                    s_stop = (apply && fabs(s_prev[0] - fit) < p.rel_fit && fabs(s_prev[1] - rmse) < p.rel_rmse) ? 1 : 0;
                    s_prev[0] = fit;
                    s_prev[1] = rmse;
                }
            }
            stamp(9);   // fit done, update stored
            __syncthreads();   // barrier B: update and stop flag visible
            if (apply) {
                iters = it + 1;
                if (s_stop) break;
            }
        }
        __syncthreads();

        // outputs: pose (cluster_icp.py:161-165), world cluster = T * S (:167), correspondences
        if (tid == 0) {
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
        if (tid < 16) p.out_T[16 * (size_t)b + tid] = s_T[tid];
        const bool aff = s_T[12] == 0.0 && s_T[13] == 0.0 && s_T[14] == 0.0 && s_T[15] == 1.0;
        for (int i = tid; i < ns; i += kNT) {
            const int t = point_to_slot(i);
            const size_t e = 3 * (size_t)(s0 + i);
            double x = ld_coord(p.src, p.pts_dtype, e), y = ld_coord(p.src, p.pts_dtype, e + 1),
                   z = ld_coord(p.src, p.pts_dtype, e + 2);
            transform_point(s_T, aff, x, y, z);
            p.out_world[e] = x; p.out_world[e + 1] = y; p.out_world[e + 2] = z;
            const int j = scj[t];
            p.out_corr[s0 + i] = (j >= 0 && sbd[t] < p.r2) ? __ldg(p.qi + q0 + j) : -1;
        }
    }
}

// host side: persistent CTAs, MINB per SM, tiles from the queue
template <int MINB, bool DBG>
static int launch_variant2(const IcpParams &P, int n_tiles, cudaStream_t stream) {
    int dev = 0;
    const int sms = current_device_sms(&dev);
    // MINB x (35.3 KB + static) per SM only fits with the carve-out at its maximum (set once per device and variant)
    if (once_per_device(4 + MINB + (DBG ? 4 : 0), dev))
        AURDF_CUDA_CHECK(cudaFuncSetAttribute(icp_small2_kernel<MINB, DBG>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    const int grid = n_tiles < sms * MINB ? n_tiles : sms * MINB;
    icp_small2_kernel<MINB, DBG><<<grid, kNT, kSmall2SmemBytes, stream>>>(P);
    return AURDF_OK;
}

int launch_icp_small2(const IcpParams &P, int n_tiles, int minb, cudaStream_t stream) {
    if (P.dbg_clock) return launch_variant2<5, true>(P, n_tiles, stream);
    if (minb == 4) return launch_variant2<4, false>(P, n_tiles, stream);
    if (minb == 5) return launch_variant2<5, false>(P, n_tiles, stream);
    return launch_variant2<6, false>(P, n_tiles, stream);
}

}  // namespace aurdf
