// icp_common.cuh -- pieces shared by the per-tile ICP kernels (icp_sweep.cu, icp_small.cu):
// kernel parameters, the exactly-rounded SE(3) point transform, and the 3x3 Kabsch pose fit.
#pragma once

#include "common.cuh"

namespace aurdf {

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
static __device__ __noinline__ void kabsch_rotation(const double (&sigma)[3][3], double (&R)[3][3], double *warm, bool have_warm) {
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
static __device__ bool kabsch_rotation_newton(const double (&sigma)[3][3], double (&R)[3][3]) {
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

// Raw reciprocal (MUFU.RCP64H seed + Newton) without range checks: inf / NaN inputs propagate and
// are caught by the caller's final tests.
__device__ __forceinline__ double rcp_raw1(double x) {   // ~2^-40
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y * (2.0 - x * y);
}
__device__ __forceinline__ double rcp_raw2(double x) {   // ~1 ulp
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = y * (2.0 - x * y);
    return y + y * (1.0 - x * y);
}

// Same fixed point as kabsch_rotation_newton, written for the serial section of the small-tile
// kernel: short code (it is re-fetched every ICP iteration), ~60 live registers, no divisions.
//   * the accumulated rotation is a quaternion: a Cayley step with vector v is the (unnormalised)
//     quaternion (1, v), so R = E1 E2 ... is q <- q (x) (1, v) and one conversion at the end;
//   * A <- (1 + v.v) E^T A is applied column by column without forming E:
//     A_j <- (1 - v.v) A_j + 2 (v (v.A_j) - v x A_j); the positive factor (1 + v.v) changes neither
//     the polar factor nor the Newton direction, so it is never divided out;
//   * a correction with |v| < 1e-8 only enters q (its effect on A is below the rounding floor), and it
//     reuses the previous adjugate (chord step) when the previous step was already below 1e-5.
// Typical late-ICP cost: one full step plus one short finish.  No range checks on the way:
// degenerate input turns into inf / NaN and fails the final tests, the caller then falls back to
// the Jacobi SVD.
static __device__ __forceinline__ bool kabsch_rotation_newton4(const double (&sigma)[3][3], double (&R)[3][3]) {
    double A[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) A[i][j] = sigma[i][j];
    const double tol = 1e-16 * (fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]));
    double qw = 1.0, qx = 0.0, qy = 0.0, qz = 0.0;
    double c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0, hrdet = 0;   // adjugate of G, 0.5 / det(G)
    double vv = 1.0;
    bool converged = false;
#pragma unroll 1
    for (int step = 0; step < 8; ++step) {
        const double kx = A[2][1] - A[1][2], ky = A[0][2] - A[2][0], kz = A[1][0] - A[0][1];
        if (fmax(fabs(kx), fmax(fabs(ky), fabs(kz))) <= tol) {
            converged = true;
            break;
        }
        if (!(vv < 1e-10)) {   // full Newton step: G = tr(S) I - S, S = sym(A); chord step otherwise
            const double g01 = -0.5 * (A[0][1] + A[1][0]), g02 = -0.5 * (A[0][2] + A[2][0]), g12 = -0.5 * (A[1][2] + A[2][1]);
            const double g00 = A[1][1] + A[2][2], g11 = A[0][0] + A[2][2], g22 = A[0][0] + A[1][1];
            c00 = g11 * g22 - g12 * g12; c01 = g02 * g12 - g01 * g22; c02 = g01 * g12 - g02 * g11;
            c11 = g00 * g22 - g02 * g02; c12 = g01 * g02 - g00 * g12; c22 = g00 * g11 - g01 * g01;
            hrdet = 0.5 * rcp_raw1(g00 * c00 + g01 * c01 + g02 * c02);
        }
        const double vx = hrdet * (c00 * kx + c01 * ky + c02 * kz);
        const double vy = hrdet * (c01 * kx + c11 * ky + c12 * kz);
        const double vz = hrdet * (c02 * kx + c12 * ky + c22 * kz);
        vv = vx * vx + vy * vy + vz * vz;
        if (!(vv < 1.0)) return false;   // more than 90 degrees in one step (or NaN): not the near-identity case
        {   // q <- q (x) (1, v)
            const double w = qw, x = qx, y = qy, z = qz;
            qw = w - (x * vx + y * vy + z * vz);
            qx = x + (w * vx + (y * vz - z * vy));
            qy = y + (w * vy + (z * vx - x * vz));
            qz = z + (w * vz + (x * vy - y * vx));
        }
        if (vv < 1e-16) {   // A would change by less than its rounding error
            converged = true;
            break;
        }
        const double omv = 1.0 - vv;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double m0 = A[0][j], m1 = A[1][j], m2 = A[2][j];
            const double d = vx * m0 + vy * m1 + vz * m2;
            A[0][j] = omv * m0 + 2.0 * (vx * d - (vy * m2 - vz * m1));
            A[1][j] = omv * m1 + 2.0 * (vy * d - (vz * m0 - vx * m2));
            A[2][j] = omv * m2 + 2.0 * (vz * d - (vx * m1 - vy * m0));
        }
        if (vv < 1e-13) {   // a step of |v| < 3e-7 leaves a residual of ~|v|^2 <= 1e-13
            converged = true;
            break;
        }
    }
    if (!converged) return false;
    // certify the maximum: sym(A) positive definite with a margin relative to tr(A) (= the nuclear
    // norm of sigma at the optimum), which also rejects det(sigma) <= 0 and near rank-deficient input
    const double tr = A[0][0] + A[1][1] + A[2][2];
    const double e1 = 1e-9 * tr, e2 = e1 * tr, e3 = e2 * tr;
    const double m2 = A[0][0] * A[1][1] - A[0][1] * A[1][0];
    if (!(A[0][0] > e1 && m2 > e2 && det3(A) > e3)) return false;
    const double s = 2.0 * rcp_raw2(qw * qw + qx * qx + qy * qy + qz * qz);
    R[0][0] = 1.0 - s * (qy * qy + qz * qz); R[0][1] = s * (qx * qy - qz * qw); R[0][2] = s * (qx * qz + qy * qw);
    R[1][0] = s * (qx * qy + qz * qw); R[1][1] = 1.0 - s * (qx * qx + qz * qz); R[1][2] = s * (qy * qz - qx * qw);
    R[2][0] = s * (qx * qz - qy * qw); R[2][1] = s * (qy * qz + qx * qw); R[2][2] = 1.0 - s * (qx * qx + qy * qy);
    return true;
}

// ------------------------------------------------------------------------------------------
// "Strict" pose fit: the CPU reference's arithmetic, operation for operation.
//
// When fewer than three non-collinear points are matched the covariance has rank <= 1, the
// optimal rotation is a one-parameter family and the member an SVD returns is decided by the last
// bits of its input and by the order of its own operations.  For such tiles (a box that holds a
// handful of target points) the ICP kernels therefore leave the fast path and reproduce
// oracle/icp_oracle.c (kabsch_sv / orc_svd3: Eigen umeyama + two-sided Jacobi as restated there)
// exactly: sums over the matched pairs in ascending source index, two passes (means, then the
// covariance of the demeaned points), every operation rounded on its own (no FMA), true division
// and square root, the same rotation formulas, sweep order, thresholds and sort.
// ------------------------------------------------------------------------------------------
namespace strict {
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double sqr(double a) { return __dsqrt_rn(a); }

// rows P,Q of M <- [c s; -s c] from the left;  cols P,Q of M <- M [c s; -s c]  (static indices: registers)
template <int P, int Q>
__device__ __forceinline__ void rot_rows(double (&M)[3][3], double c, double s) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double x = M[P][j], y = M[Q][j];
        M[P][j] = add(mul(c, x), mul(s, y));
        M[Q][j] = add(mul(-s, x), mul(c, y));
    }
}
template <int P, int Q>
__device__ __forceinline__ void rot_cols(double (&M)[3][3], double c, double s) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double x = M[i][P], y = M[i][Q];
        M[i][P] = sub(mul(c, x), mul(s, y));
        M[i][Q] = add(mul(s, x), mul(c, y));
    }
}

// one two-sided Jacobi step on the index pair (Q, P), Q < P, exactly as in orc_svd3
template <int P, int Q>
__device__ __forceinline__ void jacobi_step(double (&W)[3][3], double (&U)[3][3], double (&V)[3][3], double &maxdiag,
                                            bool &finished) {
    const double precision = 2.0 * 2.220446049250313e-16;
    const double tiny = 2.2250738585072014e-308;
    const double thr = fmax(tiny, mul(precision, maxdiag));
    if (!(fabs(W[P][Q]) > thr || fabs(W[Q][P]) > thr)) return;
    finished = false;
    const double m00 = W[Q][Q], m01 = W[Q][P], m10 = W[P][Q], m11 = W[P][P];
    const double t = add(m00, m11), d = sub(m10, m01);
    double c1, s1;
    if (fabs(d) < tiny) { c1 = 1.0; s1 = 0.0; }
    else {
        const double u = dvd(t, d), tmp = sqr(add(1.0, mul(u, u)));
        s1 = dvd(1.0, tmp); c1 = dvd(u, tmp);
    }
    const double a00 = add(mul(c1, m00), mul(s1, m10)), a01 = add(mul(c1, m01), mul(s1, m11));
    const double a11 = add(mul(-s1, m01), mul(c1, m11));
    double c2, s2;
    if (fabs(a01) < tiny) { c2 = 1.0; s2 = 0.0; }
    else {
        const double tau = dvd(sub(a00, a11), mul(2.0, a01)), w = sqr(add(mul(tau, tau), 1.0));
        const double tn = (tau >= 0) ? dvd(-1.0, add(tau, w)) : dvd(-1.0, sub(tau, w));
        c2 = dvd(1.0, sqr(add(mul(tn, tn), 1.0)));
        s2 = mul(tn, c2);
    }
    const double cl = add(mul(c2, c1), mul(s2, s1));
    const double sl = sub(mul(c2, s1), mul(s2, c1));
    rot_rows<Q, P>(W, cl, sl);
    rot_cols<Q, P>(W, c2, s2);
    rot_cols<Q, P>(U, cl, -sl);
    rot_cols<Q, P>(V, c2, s2);
    maxdiag = fmax(maxdiag, fmax(fabs(W[P][P]), fabs(W[Q][Q])));
}

template <int I, int K>
__device__ __forceinline__ void swap_sv(double (&S)[3], double (&U)[3][3], double (&V)[3][3]) {
    double t = S[I]; S[I] = S[K]; S[K] = t;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        t = U[r][I]; U[r][I] = U[r][K]; U[r][K] = t;
        t = V[r][I]; V[r][I] = V[r][K]; V[r][K] = t;
    }
}

// A = U diag(S) V^T exactly as orc_svd3 computes it
__device__ __forceinline__ void svd3(const double (&A)[3][3], double (&U)[3][3], double (&S)[3], double (&V)[3][3]) {
    double W[3][3];
    double scale = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            U[i][j] = V[i][j] = (i == j) ? 1.0 : 0.0;
            const double a = fabs(A[i][j]);
            if (a > scale) scale = a;
        }
    if (scale == 0.0 || !(scale == scale)) scale = 1.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) W[i][j] = dvd(A[i][j], scale);
    double maxdiag = fmax(fabs(W[0][0]), fmax(fabs(W[1][1]), fabs(W[2][2])));
    bool finished = false;
#pragma unroll 1
    for (int sweeps = 0; !finished && sweeps < 64; ++sweeps) {
        finished = true;
        jacobi_step<1, 0>(W, U, V, maxdiag, finished);
        jacobi_step<2, 0>(W, U, V, maxdiag, finished);
        jacobi_step<2, 1>(W, U, V, maxdiag, finished);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double a = fabs(W[i][i]);
        S[i] = a;
        if (a != 0.0 && W[i][i] < 0.0) {
#pragma unroll
            for (int r = 0; r < 3; ++r) U[r][i] = -U[r][i];
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) S[i] = mul(S[i], scale);
    // selection sort, descending, stopping at the first zero maximum (orc_svd3), with static indices
    {
        const int k = (S[1] > S[0]) ? ((S[2] > S[1]) ? 2 : 1) : ((S[2] > S[0]) ? 2 : 0);
        const double sk = k == 0 ? S[0] : (k == 1 ? S[1] : S[2]);
        if (sk != 0.0) {
            if (k == 1) swap_sv<0, 1>(S, U, V);
            else if (k == 2) swap_sv<0, 2>(S, U, V);
            if (S[2] > S[1]) {
                if (S[2] != 0.0) swap_sv<1, 2>(S, U, V);
            }
        }
    }
}

__device__ __forceinline__ double det3s(const double (&M)[3][3]) {   // the oracle's cofactor expansion, each op rounded
    const double a = mul(M[0][0], sub(mul(M[1][1], M[2][2]), mul(M[1][2], M[2][1])));
    const double b = mul(M[0][1], sub(mul(M[1][0], M[2][2]), mul(M[1][2], M[2][0])));
    const double c = mul(M[0][2], sub(mul(M[1][0], M[2][1]), mul(M[1][1], M[2][0])));
    return add(sub(a, b), c);
}

// update U (row-major 3x4 into Um[12]) from the two-pass covariance sigma and the means; kabsch_sv of the oracle
static __device__ __noinline__ void pose_from_sigma(const double (&sigma_in)[3][3], const double (&ms)[3], const double (&md)[3],
                                                    double *Um) {
    double sigma[3][3], U[3][3], S[3], V[3][3], R[3][3];   // local copies: everything below stays in registers
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) sigma[i][j] = sigma_in[i][j];
    svd3(sigma, U, S, V);
    const double D[3] = {1.0, 1.0, (mul(det3s(U), det3s(V)) < 0) ? -1.0 : 1.0};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) s = add(s, mul(mul(U[r][k], D[k]), V[c][k]));
            R[r][c] = s;
        }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double rm = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) rm = add(rm, mul(R[r][k], ms[k]));
        Um[4 * r + 0] = R[r][0];
        Um[4 * r + 1] = R[r][1];
        Um[4 * r + 2] = R[r][2];
        Um[4 * r + 3] = sub(md[r], rm);
    }
}
}  // namespace strict

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
    int small_on;          // small tiles are taken by the small-tile kernel, the general kernel skips them
    int split_tail;        // icp_small_kernel: widen the lane split of a tile's last, partly filled round
    int small_ns, small_nt;   // class bounds of the small-tile kernel in use (source points, masked targets)
    int n_tiles;
    int *queue;            // small-tile kernel v2: next tile to take (zeroed by tile_scan_kernel)
    int *queue2;           // grid kernel, one-CTA variant: its own tile queue (zeroed by tile_scan_kernel)
    int *glist;            // workspace: the tiles the grid kernel owns (written by tile_scan_kernel, count in status_int[5])
    int strict_nt;         // tiles with at most this many masked targets fit their poses in strict mode
    // grid-pruned search (icp_grid.cu): medium / large tiles
    int grid_on;           // 0: every non-small tile runs in the brute-force general kernel
    int grid_cs;           // CTAs per tile of the grid kernel launched for this sweep (1 or 8)
    int grid_pcap;         // most source points one of its CTAs keeps in shared memory
    int grid_smem_bytes;   // shared memory of a CTA set aside for the tile's sorted targets + cell table
    float grid_cell_scale; // cell size relative to the one-target-per-cell-by-volume rule
    float4 *gs;            // workspace: cell-sorted float32 targets (x, y, z about the tile origin, compacted index)
    int *gends;            // workspace: end offset of every grid cell, tile b at toff[b] + 2 b
    float *gpar;           // workspace: 8 floats per tile (grid origin, cell size, 1 / cell size, max |coordinate|, cells per axis)
};

// Tile classes.  A "small" tile keeps its whole state in < 30 KB of shared memory, so that 7 tiles
// are resident per SM (all 900 tiles of wx200_5 start at once): icp_small.cu.  Anything larger
// runs in the general chunk-streaming / cluster kernel of icp_sweep.cu.
constexpr int kSmNs = 320;     // most source points of a small tile
constexpr int kSmPairs = 384;  // float32 target pairs in shared memory, incl. the scan's read-ahead padding
constexpr int kSmNt32 = 2 * kSmPairs - 8;   // most masked targets of a small tile (760)
constexpr int kSmNt64 = 384;   // most targets whose float64 copy also lives in shared memory
__device__ __forceinline__ bool tile_is_small(int ns, int nt) { return ns <= kSmNs && nt <= kSmNt32; }
constexpr int kS2Ns = 384;     // small-tile kernel v2: most source points (three rounds of 128 home slots)

// Tile classes of one sweep: small (icp_small*.cu), grid (icp_grid.cu), general (icp_tiles_kernel).  Every kernel
// evaluates the same predicate and leaves the tiles it does not own.
__device__ __forceinline__ bool tile_uses_small(const IcpParams &p, int ns_tile, int nt) {
    return p.small_on && ns_tile <= p.small_ns && nt <= p.small_nt;
}
__device__ __forceinline__ bool tile_uses_grid(const IcpParams &p, int ns_tile, int nt) {
    if (!p.grid_on || tile_uses_small(p, ns_tile, nt)) return false;
    if (p.grid_cs == 1) return true;   // one CTA per tile: everything the small-tile kernel leaves
    // cluster variant: rank-deficient tiles (strict pose fit: one CTA sums in source order) and slices beyond
    // the shared-memory bound stay with the general kernel
    return nt > p.strict_nt && (ns_tile + p.grid_cs - 1) / p.grid_cs <= p.grid_pcap;
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

}  // namespace aurdf
