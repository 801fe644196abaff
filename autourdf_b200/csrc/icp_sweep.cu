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
// Compile with -fmad=false: distances and point transforms must round exactly like the
// float64 CPU reference (no FMA contraction), otherwise near-tie correspondences flip.
#include <math.h>
#include <stdio.h>

#include <vector>

#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace aurdf {

constexpr int kIcpThreads = 256;
constexpr int kIcpWarps = kIcpThreads / 32;
constexpr int kQChunk = 1024;     // target points staged per shared-memory chunk (24 KB f64 SoA + 16 KB float4)
constexpr int kClusterCtas = 8;   // CTAs per tile in the cluster variant (large tiles)
constexpr int kPSmemMax = 2048;   // most source points a tile may keep in shared memory

struct WsLayout {
    size_t box, cnt, toff, qx, qy, qz, qi, pspill, status, total;
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
    L.status = take(4 * sizeof(int));
    L.qx = take((size_t)cap * sizeof(double));
    L.qy = take((size_t)cap * sizeof(double));
    L.qz = take((size_t)cap * sizeof(double));
    L.qi = take((size_t)cap * sizeof(int));
    L.pspill = take((size_t)total_src * 3 * sizeof(double));
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
                 int *__restrict__ status_int, int *__restrict__ status_user) {
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
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
    if (tid == 0) {
        long long total = s_carry;
        toff[B] = total;
        int over = total > capacity ? 1 : 0;
        status_int[0] = over;
        status_int[1] = (int)(total & 0xffffffffLL);
        status_int[2] = (int)(total >> 32);
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
// 3x3 SVD by two-sided Jacobi (Eigen JacobiSVD semantics): A = U diag(S) V^T,
// S sorted descending and non-negative.  Static indices keep everything in registers.
// This is the serial section of every ICP iteration (one lane), so it is written for latency:
// three reciprocal square roots per rotation and no division or square root.
// ------------------------------------------------------------------------------------------

// 1/sqrt(x) to ~1 ulp: MUFU.RSQ64H seed (rsqrt.approx.f64, ~20 bits) + two Newton steps.
__device__ __forceinline__ double fast_rsqrt(double x) {
    if (!(x > 1e-290 && x < 1e290)) return rsqrt(x);  // subnormal / huge / NaN: library path
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = y * (1.5 - hx * y * y);
    y = y * (1.5 - hx * y * y);
    return y;
}

template <int P, int Q>
__device__ __forceinline__ void jacobi_pair(double (&W)[3][3], double (&U)[3][3], double (&V)[3][3],
                                            double &maxdiag, bool &finished) {
    const double tiny = 2.2250738585072014e-308;
    const double thr = fmax(tiny, 4.440892098500626e-16 * maxdiag);
    if (!(fabs(W[P][Q]) > thr || fabs(W[Q][P]) > thr)) return;
    finished = false;
    // 2x2 block on (Q,P), Q < P
    const double m00 = W[Q][Q], m01 = W[Q][P], m10 = W[P][Q], m11 = W[P][P];
    // step 1: rotation R1 = [c1 s1; -s1 c1] that makes the block symmetric: (c1,s1) = (t,d)/|(t,d)|
    const double t = m00 + m11, d = m10 - m01;
    double c1 = 1.0, s1 = 0.0;
    const double n1 = t * t + d * d;
    if (fabs(d) >= tiny && n1 > 1e-290) {
        const double r = fast_rsqrt(n1);
        c1 = t * r;
        s1 = d * r;
    }
    const double a00 = c1 * m00 + s1 * m10, a01 = c1 * m01 + s1 * m11, a11 = -s1 * m01 + c1 * m11;
    // step 2: symmetric Jacobi J = [c2 s2; -s2 c2] with tan = t2 the small root of
    // t^2 - 2 tau t - 1 = 0, tau = h / (2 a01), h = a00 - a11:
    //   (c2, s2) = (|h| + w, -sgn(h) 2 a01) / norm,  w = sqrt(h^2 + 4 a01^2)
    double c2 = 1.0, s2 = 0.0;
    if (fabs(a01) >= tiny) {
        const double h = a00 - a11, b2 = 2.0 * a01;
        const double q = h * h + b2 * b2;
        if (q > 1e-290) {
            const double w = q * fast_rsqrt(q);
            const double cx = fabs(h) + w, sx = (h >= 0 ? -b2 : b2);
            const double r2 = fast_rsqrt(cx * cx + sx * sx);
            c2 = cx * r2;
            s2 = sx * r2;
        }
    }
    const double cl = c2 * c1 + s2 * s1, sl = c2 * s1 - s2 * c1;
#pragma unroll
    for (int j = 0; j < 3; ++j) {  // rows (Q,P) <- L * rows
        const double x = W[Q][j], y = W[P][j];
        W[Q][j] = cl * x + sl * y;
        W[P][j] = -sl * x + cl * y;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // cols (Q,P) <- cols * J ; U <- U L^T ; V <- V J
        double x = W[i][Q], y = W[i][P];
        W[i][Q] = c2 * x - s2 * y;
        W[i][P] = s2 * x + c2 * y;
        x = U[i][Q];
        y = U[i][P];
        U[i][Q] = cl * x + sl * y;
        U[i][P] = -sl * x + cl * y;
        x = V[i][Q];
        y = V[i][P];
        V[i][Q] = c2 * x - s2 * y;
        V[i][P] = s2 * x + c2 * y;
    }
    maxdiag = fmax(maxdiag, fmax(fabs(W[P][P]), fabs(W[Q][Q])));
}

template <int I, int K>
__device__ __forceinline__ void swap_cols(double (&S)[3], double (&U)[3][3], double (&V)[3][3]) {
    double t = S[I];
    S[I] = S[K];
    S[K] = t;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        t = U[r][I]; U[r][I] = U[r][K]; U[r][K] = t;
        t = V[r][I]; V[r][I] = V[r][K]; V[r][K] = t;
    }
}

__device__ __forceinline__ double det3(const double (&M)[3][3]) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

// Kabsch rotation of a 3x3 covariance (Eigen umeyama without scaling): R = U diag(1,1,s) V^T.
// warm (18 doubles: U then V, row-major) carries the singular vectors of the previous ICP
// iteration of the same tile: W = U^T sigma V is then already nearly diagonal and the Jacobi
// iteration converges in about two sweeps instead of six.  warm is updated in place.
// (No pre-scaling by max|sigma|: the convergence threshold is relative and covariances of
// metre-scale clouds are nowhere near the float64 range limits.)
__device__ void kabsch_rotation(const double (&sigma)[3][3], double (&R)[3][3], double *warm, bool have_warm) {
    double W[3][3], U[3][3], V[3][3];
    if (have_warm) {
        double SV[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                U[i][j] = warm[3 * i + j];
                V[i][j] = warm[9 + 3 * i + j];
            }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) SV[i][j] = sigma[i][0] * V[0][j] + sigma[i][1] * V[1][j] + sigma[i][2] * V[2][j];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) W[i][j] = U[0][i] * SV[0][j] + U[1][i] * SV[1][j] + U[2][i] * SV[2][j];
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                W[i][j] = sigma[i][j];
                U[i][j] = V[i][j] = (i == j) ? 1.0 : 0.0;
            }
    }
    double maxdiag = fmax(fabs(W[0][0]), fmax(fabs(W[1][1]), fabs(W[2][2])));
    bool finished = false;
    for (int sweep = 0; sweep < 64 && !finished; ++sweep) {
        finished = true;
        jacobi_pair<1, 0>(W, U, V, maxdiag, finished);
        jacobi_pair<2, 0>(W, U, V, maxdiag, finished);
        jacobi_pair<2, 1>(W, U, V, maxdiag, finished);
    }
    double S[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double a = fabs(W[i][i]);
        S[i] = a;
        if (a != 0.0 && W[i][i] < 0.0) {
#pragma unroll
            for (int r = 0; r < 3; ++r) U[r][i] = -U[r][i];
        }
    }
    // sort descending (3-element network equivalent to Eigen's selection sort)
    if (S[1] > S[0] && S[1] >= S[2]) swap_cols<0, 1>(S, U, V);
    else if (S[2] > S[0] && S[2] > S[1]) swap_cols<0, 2>(S, U, V);
    if (S[2] > S[1]) swap_cols<1, 2>(S, U, V);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            warm[3 * i + j] = U[i][j];
            warm[9 + 3 * i + j] = V[i][j];
        }
    const double sgn = (det3(U) * det3(V) < 0) ? -1.0 : 1.0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) R[r][c] = U[r][0] * V[c][0] + U[r][1] * V[c][1] + sgn * U[r][2] * V[c][2];
}

// 1/x to ~2^-40: MUFU.RCP64H seed + one Newton step.  Only used inside Newton iterations that
// self-correct, never for a value that is output.
__device__ __forceinline__ double fast_rcp(double x) {
    if (!(fabs(x) > 1e-290 && fabs(x) < 1e290)) return 1.0 / x;
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = y * (2.0 - x * y);
    return y;
}

// 1/x to ~1 ulp: MUFU.RCP64H seed + two Newton steps (a true division is ~25 dependent
// instructions; this is ~6).
__device__ __forceinline__ double rcp_nr2(double x) {
    if (!(fabs(x) > 1e-290 && fabs(x) < 1e290)) return 1.0 / x;
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = y * (2.0 - x * y);
    y = y + y * (1.0 - x * y);
    return y;
}

// Kabsch rotation for the common case, written for a short dependent chain (this is the
// serial section of every ICP iteration).  The source points are re-posed every iteration, so
// the optimal rotation R = argmax tr(R^T sigma) is near the identity.  R is optimal and proper
// iff A = R^T sigma is symmetric positive definite (then R is the polar factor U V^T, which is
// what umeyama returns when det(sigma) > 0).  Newton on SO(3): with S = sym(A) and
// k = axial(A - A^T), solve (tr(S) I - S) w = k, rotate by the Cayley transform of w (an exact
// rotation for any w), repeat; quadratic convergence, 2-4 steps.  Returns false when the result
// cannot be certified (no convergence, A not positive definite: reflection or rank-deficient
// input) and the caller falls back to the Jacobi SVD.
__device__ bool kabsch_rotation_newton(const double (&sigma)[3][3], double (&R)[3][3]) {
    double A[3][3];
    double scale = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            A[i][j] = sigma[i][j];
            R[i][j] = (i == j) ? 1.0 : 0.0;
            scale = fmax(scale, fabs(sigma[i][j]));
        }
    if (!(scale > 1e-280 && scale < 1e280)) return false;
    const double tol = 1e-16 * scale;  // below the rounding floor of A: only exact stationarity exits here
    bool converged = false;
    for (int step = 0; step < 8; ++step) {
        const double kx = A[2][1] - A[1][2], ky = A[0][2] - A[2][0], kz = A[1][0] - A[0][1];
        if (fmax(fabs(kx), fmax(fabs(ky), fabs(kz))) <= tol) {
            converged = true;
            break;
        }
        // G = tr(S) I - S (symmetric), S = sym(A)
        const double s01 = 0.5 * (A[0][1] + A[1][0]), s02 = 0.5 * (A[0][2] + A[2][0]), s12 = 0.5 * (A[1][2] + A[2][1]);
        const double g00 = A[1][1] + A[2][2], g11 = A[0][0] + A[2][2], g22 = A[0][0] + A[1][1];
        const double g01 = -s01, g02 = -s02, g12 = -s12;
        // w = G^-1 k by the adjugate
        const double c00 = g11 * g22 - g12 * g12, c01 = g02 * g12 - g01 * g22, c02 = g01 * g12 - g02 * g11;
        const double c11 = g00 * g22 - g02 * g02, c12 = g01 * g02 - g00 * g12, c22 = g00 * g11 - g01 * g01;
        const double det = g00 * c00 + g01 * c01 + g02 * c02;
        if (!(fabs(det) > 1e-280)) return false;
        const double rdet = fast_rcp(det);
        // v = w / 2
        const double vx = 0.5 * rdet * (c00 * kx + c01 * ky + c02 * kz);
        const double vy = 0.5 * rdet * (c01 * kx + c11 * ky + c12 * kz);
        const double vz = 0.5 * rdet * (c02 * kx + c12 * ky + c22 * kz);
        const double vv = vx * vx + vy * vy + vz * vz;
        if (!(vv < 1.0)) return false;  // more than 90 degrees in one step: not the near-identity case
        // Cayley: E = ((1 - vv) I + 2 v v^T + 2 [v]x) / (1 + vv), an exact rotation
        const double rden = rcp_nr2(1.0 + vv);
        const double a = (1.0 - vv) * rden, b2 = 2.0 * rden;
        double E[3][3];
        E[0][0] = a + b2 * vx * vx; E[0][1] = b2 * (vx * vy - vz); E[0][2] = b2 * (vx * vz + vy);
        E[1][0] = b2 * (vx * vy + vz); E[1][1] = a + b2 * vy * vy; E[1][2] = b2 * (vy * vz - vx);
        E[2][0] = b2 * (vx * vz - vy); E[2][1] = b2 * (vy * vz + vx); E[2][2] = a + b2 * vz * vz;
        double An[3][3], Rn[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                An[i][j] = E[0][i] * A[0][j] + E[1][i] * A[1][j] + E[2][i] * A[2][j];  // E^T A
                Rn[i][j] = R[i][0] * E[0][j] + R[i][1] * E[1][j] + R[i][2] * E[2][j];  // R E
            }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                A[i][j] = An[i][j];
                R[i][j] = Rn[i][j];
            }
        // quadratic convergence: a step of |v| < 1e-7 leaves an error of ~|v|^2 <= 1e-14
        if (vv < 1e-14) {
            converged = true;
            break;
        }
    }
    if (!converged) return false;
    // certify the maximum: sym(A) positive definite with a margin (Sylvester), which also
    // rejects det(sigma) <= 0 and near rank-deficient covariances
    const double m1 = A[0][0];
    const double m2 = A[0][0] * A[1][1] - A[0][1] * A[1][0];
    const double m3 = det3(A);
    const double eps = 1e-9;
    return m1 > eps * scale && m2 > eps * scale * scale && m3 > eps * scale * scale * scale;
}

// ------------------------------------------------------------------------------------------
// 4. fused per-tile ICP
// ------------------------------------------------------------------------------------------
struct IcpParams {
    const void *src;
    int pts_dtype;
    const int *src_off;
    const double *init_T;
    double r2;
    int max_iter;
    double rel_fit, rel_rmse;
    int ori_only;
    const double *qx, *qy, *qz;
    const int *qi;
    const long long *toff;
    const int *cnt;
    const int *status_int;
    double *pspill;
    int p_cap;  // source points that fit in shared memory
    double *out_T, *out_world;
    int *out_corr;
    double *out_fit, *out_rmse;
    int *out_iters, *out_ntgt;
    long long *dbg_clock;  // optional: per-phase cycle stamps of tile 0 (debug builds of the bench only)
};

// x' = ((m0 x + m1 y) + m2 z) + m3, each operation rounded (open3d PointCloud::Transform)
__device__ __forceinline__ double affine_row(const double *m, double x, double y, double z) {
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[0], x), __dmul_rn(m[1], y)), __dmul_rn(m[2], z)), m[3]);
}

__device__ __forceinline__ void transform_point(const double *T, bool affine, double &x, double &y, double &z) {
    const double nx = affine_row(T, x, y, z), ny = affine_row(T + 4, x, y, z), nz = affine_row(T + 8, x, y, z);
    if (affine) {  // last row (0,0,0,1): w == 1 exactly, the division is a bit-exact no-op
        x = nx; y = ny; z = nz;
    } else {
        const double w = affine_row(T + 12, x, y, z);
        x = nx / w; y = ny / w; z = nz / w;
    }
}

// Sum 16 per-lane values over the 32 lanes of a warp with 16 shuffles (a transposing
// butterfly: every step halves the number of values a lane carries).  Afterwards v[0] holds
// the warp total of value number (lane >> 1).
__device__ __forceinline__ void warp_sum16(double (&v)[16], int lane) {
#pragma unroll
    for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            const double send = hi ? v[k] : v[k + half];
            const double keep = hi ? v[k + half] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

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
    const int per = (ns_tile + CS - 1) / CS;
    const int lo = min(ns_tile, rank * per);
    const int s0 = p.src_off[b] + lo;                    // this CTA's slice of the source points
    const int ns = min(ns_tile, lo + per) - lo;
    const int nt = p.cnt[b];
    const long long q0 = p.toff[b];
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
            if (!kabsch_rotation_newton(sigma, R)) {
                kabsch_rotation(sigma, R, s_warm, have_warm);  // reflection / rank-deficient / large step
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
            if (lane == 0 && p.max_iter > 0) fit_pose(c);
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
                if (lane == 0 && it + 1 < p.max_iter) fit_pose(c);
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

static long long *g_dbg_clock = nullptr;
extern "C" __attribute__((visibility("default"))) void aurdf_debug_set_clock_buffer(void *p) { g_dbg_clock = (long long *)p; }

extern "C" int aurdf_icp_sweep_launches(void) { return 4; }

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

    box_count_kernel<<<n_tiles, kIcpThreads, 0, stream>>>(box_xyz, box_dtype, box_off, tgt_xyz, pts_dtype, tgt_off,
                                                          tile_frame, box_scale, box, cnt);
    tile_scan_kernel<<<1, 1024, 0, stream>>>(cnt, n_tiles, (long long)tgt_capacity, toff, status_int, status);
    mask_fill_kernel<<<n_tiles, kIcpThreads, 0, stream>>>(tgt_xyz, pts_dtype, tgt_off, tile_frame, box, toff,
                                                          status_int, qx, qy, qz, qi);
    IcpParams P;
    P.src = src_xyz; P.pts_dtype = pts_dtype; P.src_off = src_off; P.init_T = init_T;
    P.r2 = max_corr_dist * max_corr_dist; P.max_iter = max_iter; P.rel_fit = rel_fitness; P.rel_rmse = rel_rmse;
    P.ori_only = ori_only;
    P.qx = qx; P.qy = qy; P.qz = qz; P.qi = qi; P.toff = toff; P.cnt = cnt; P.status_int = status_int;
    P.pspill = pspill; P.p_cap = p_cap;
    P.out_T = out_T; P.out_world = out_world_xyz; P.out_corr = out_corr; P.out_fit = out_fitness;
    P.out_rmse = out_rmse; P.out_iters = out_iters; P.out_ntgt = out_ntgt;
    P.dbg_clock = g_dbg_clock;
    EvPair ev{nullptr, nullptr};
    if (g_prof_on) {
        AURDF_CUDA_CHECK(cudaEventCreate(&ev.a));
        AURDF_CUDA_CHECK(cudaEventCreate(&ev.b));
        AURDF_CUDA_CHECK(cudaEventRecord(ev.a, stream));
    }
    if (use_cluster) {
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
