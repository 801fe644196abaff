/*
 * oracle/icp_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
 *
 * CPU restatement (plain C, float64, -ffp-contract=off) of the cluster-ICP hot
 * path of jl6017/AutoURDF.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.
 *
 * PARITY UNPINNED at the third-party boundary: the arithmetic of this path
 * lives in open3d==0.18.0 (requirements.txt:5), which is not vendored in
 * /root/reference and not installable here; the reference holds no tests,
 * golden vectors or fixtures.  What IS pinned: the outer
 * sweep (box / mask / ori / outputs) is checked against the reference's own
 * PointCloud/cluster_icp.py executed with this restatement standing in for
 * open3d (tests/golden/make_golden.py), and every numeric piece is checked
 * against independent implementations (scipy cKDTree, numpy LAPACK SVD, scipy
 * Rotation.align_vectors).
 *
 * What is restated, and where it comes from:
 *   orc_aabb_box        PointCloud/cluster_icp.py:133-140  (box, centre, 1.2x inflate;
 *                       float32 arithmetic when the predicted cluster is float32,
 *                       which is what mlp_reg.py:121 hands over)
 *   orc_mask            PointCloud/cluster_icp.py:142-148  (strict > / < on all 3 axes,
 *                       order-preserving boolean-mask gather)
 *   orc_icp_p2p         open3d 0.18.0 pipelines/registration/Registration.cpp
 *                       RegistrationICP + GetRegistrationResultAndCorrespondences,
 *                       call site PointCloud/cluster_icp.py:157-159 (also link.py:113-117,
 *                       Sim/evaluation.py:358-362)
 *   orc_kabsch          TransformationEstimationPointToPoint::ComputeTransformation ->
 *                       Eigen 3.4 umeyama(src, dst, with_scaling=false)
 *   orc_svd3            Eigen JacobiSVD semantics (full U,V, singular values sorted
 *                       descending, non-negative)
 *   orc_transform_pts   open3d geometry::PointCloud::Transform (4x4 * [x y z 1], / w),
 *                       call sites cluster_icp.py:167, :96-98, mlp_reg.py:211-213
 *   orc_masked_icp_sweep  PointCloud/cluster_icp.py:118-191 (masked_icp), batched over
 *                       (frame, cluster) tiles, OpenMP over tiles.
 *
 * Nearest neighbour: squared L2 accumulated as ((0+dx*dx)+dy*dy)+dz*dz in double
 * (nanoflann L2_Simple_Adaptor order); accepted iff d2 < r*r (strict; SearchHybrid with
 * max_nn=1).  Exact ties resolve to the lowest target index (nanoflann's order on exact
 * ties is traversal dependent and cannot be pinned; inputs avoid duplicate points).
 * Two NN back-ends give identical answers: brute force, and an exact k-d tree (the
 * algorithmic analogue of the reference's nanoflann tree; used for the CPU baseline).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* 4x4 helpers (row-major double[16])                                         */
/* ------------------------------------------------------------------------- */

/* C = A * B, each entry summed left to right over k. */
ORC_API void orc_mat4_mul(const double *A, const double *B, double *C) {
    double R[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = A[i * 4 + 0] * B[0 * 4 + j];
            s = s + A[i * 4 + 1] * B[1 * 4 + j];
            s = s + A[i * 4 + 2] * B[2 * 4 + j];
            s = s + A[i * 4 + 3] * B[3 * 4 + j];
            R[i * 4 + j] = s;
        }
    memcpy(C, R, sizeof R);
}

/* out = (T * [p,1]).xyz / w ; in-place allowed.  open3d PointCloud::Transform. */
ORC_API void orc_transform_pts(const double *T, const double *in, int n, double *out) {
    for (int i = 0; i < n; ++i) {
        double x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];
        double nx = ((T[0] * x + T[1] * y) + T[2] * z) + T[3];
        double ny = ((T[4] * x + T[5] * y) + T[6] * z) + T[7];
        double nz = ((T[8] * x + T[9] * y) + T[10] * z) + T[11];
        double nw = ((T[12] * x + T[13] * y) + T[14] * z) + T[15];
        out[3 * i] = nx / nw;
        out[3 * i + 1] = ny / nw;
        out[3 * i + 2] = nz / nw;
    }
}

/* General 4x4 inverse (Gauss-Jordan, partial pivoting) -- np.linalg.inv stand-in for
 * the local-frame move at mlp_reg.py:211 / cluster_icp.py:96.  Returns 0 on success. */
ORC_API int orc_mat4_inv(const double *A, double *Ainv) {
    double M[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            M[i][j] = A[i * 4 + j];
            M[i][j + 4] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r)
            if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
        if (M[piv][c] == 0.0) return -1;
        if (piv != c)
            for (int j = 0; j < 8; ++j) { double t = M[c][j]; M[c][j] = M[piv][j]; M[piv][j] = t; }
        double d = M[c][c];
        for (int j = 0; j < 8; ++j) M[c][j] /= d;
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            double f = M[r][c];
            if (f == 0.0) continue;
            for (int j = 0; j < 8; ++j) M[r][j] -= f * M[c][j];
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) Ainv[i * 4 + j] = M[i][j + 4];
    return 0;
}

/* ------------------------------------------------------------------------- */
/* 3x3 SVD: two-sided Jacobi (Eigen JacobiSVD semantics)                      */
/* ------------------------------------------------------------------------- */

/* rows p,q of M <- [c s; -s c] applied from the left:  x' = c x + s y ; y' = -s x + c y */
static void rot_rows(double M[3][3], int p, int q, double c, double s) {
    for (int j = 0; j < 3; ++j) {
        double x = M[p][j], y = M[q][j];
        M[p][j] = c * x + s * y;
        M[q][j] = -s * x + c * y;
    }
}
/* cols p,q of M <- M * [c s; -s c]:  x' = c x - s y ; y' = s x + c y */
static void rot_cols(double M[3][3], int p, int q, double c, double s) {
    for (int i = 0; i < 3; ++i) {
        double x = M[i][p], y = M[i][q];
        M[i][p] = c * x - s * y;
        M[i][q] = s * x + c * y;
    }
}

/* A = U * diag(S) * V^T ; U,V orthogonal (det may be -1), S sorted descending, S >= 0. */
ORC_API void orc_svd3(const double *A, double *Uo, double *So, double *Vo) {
    double W[3][3], U[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double scale = 0.0;
    for (int i = 0; i < 9; ++i) {
        double a = fabs(A[i]);
        if (a > scale) scale = a;
    }
    if (scale == 0.0 || !(scale == scale)) scale = 1.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) W[i][j] = A[i * 3 + j] / scale;

    const double precision = 2.0 * DBL_EPSILON;
    const double tiny = DBL_MIN;
    double maxdiag = fmax(fabs(W[0][0]), fmax(fabs(W[1][1]), fabs(W[2][2])));
    int finished = 0, sweeps = 0;
    while (!finished && sweeps < 64) {
        finished = 1;
        ++sweeps;
        for (int p = 1; p < 3; ++p)
            for (int q = 0; q < p; ++q) {
                double thr = fmax(tiny, precision * maxdiag);
                if (fabs(W[p][q]) > thr || fabs(W[q][p]) > thr) {
                    finished = 0;
                    /* 2x2 block m = [[W_qq W_qp],[W_pq W_pp]] on indices (q,p), q<p.
                     * Step 1: rotation that makes it symmetric. */
                    double m00 = W[q][q], m01 = W[q][p], m10 = W[p][q], m11 = W[p][p];
                    double t = m00 + m11, d = m10 - m01;
                    double c1, s1;
                    if (fabs(d) < tiny) { c1 = 1.0; s1 = 0.0; }
                    else {
                        double u = t / d, tmp = sqrt(1.0 + u * u);
                        s1 = 1.0 / tmp; c1 = u / tmp;
                    }
                    /* left-apply R1 = [c1 s1; -s1 c1] on the 2x2 */
                    double a00 = c1 * m00 + s1 * m10, a01 = c1 * m01 + s1 * m11;
                    double a11 = -s1 * m01 + c1 * m11;
                    /* Step 2: symmetric Jacobi on [[a00 a01],[a01 a11]]: J^T A J diagonal,
                     * J = [c2 s2; -s2 c2]. */
                    double c2, s2;
                    if (fabs(a01) < tiny) { c2 = 1.0; s2 = 0.0; }
                    else {
                        /* small root of t^2 - 2 tau t - 1 = 0, t = s2/c2 */
                        double tau = (a00 - a11) / (2.0 * a01), w = sqrt(tau * tau + 1.0);
                        double tn = (tau >= 0) ? -1.0 / (tau + w) : -1.0 / (tau - w);
                        c2 = 1.0 / sqrt(tn * tn + 1.0);
                        s2 = tn * c2;
                    }
                    /* left rotation L = J^T * R1 = [cl sl; -sl cl] */
                    double cl = c2 * c1 + s2 * s1;
                    double sl = c2 * s1 - s2 * c1;
                    /* W <- L W J ;  U <- U L^T ;  V <- V J   (A/scale = U W V^T invariant) */
                    rot_rows(W, q, p, cl, sl);
                    rot_cols(W, q, p, c2, s2);
                    rot_cols(U, q, p, cl, -sl);
                    rot_cols(V, q, p, c2, s2);
                    maxdiag = fmax(maxdiag, fmax(fabs(W[p][p]), fabs(W[q][q])));
                }
            }
    }
    /* singular values = |diag| ; fold sign into U */
    double S[3];
    for (int i = 0; i < 3; ++i) {
        double a = fabs(W[i][i]);
        S[i] = a;
        if (a != 0.0 && W[i][i] < 0.0)
            for (int r = 0; r < 3; ++r) U[r][i] = -U[r][i];
    }
    for (int i = 0; i < 3; ++i) S[i] *= scale;
    /* sort descending (selection, like Eigen: swap columns of U and V) */
    for (int i = 0; i < 3; ++i) {
        int k = i;
        for (int j = i + 1; j < 3; ++j)
            if (S[j] > S[k]) k = j;
        if (S[k] == 0.0) break;
        if (k != i) {
            double t = S[i]; S[i] = S[k]; S[k] = t;
            for (int r = 0; r < 3; ++r) {
                t = U[r][i]; U[r][i] = U[r][k]; U[r][k] = t;
                t = V[r][i]; V[r][i] = V[r][k]; V[r][k] = t;
            }
        }
    }
    for (int i = 0; i < 3; ++i) {
        So[i] = S[i];
        for (int j = 0; j < 3; ++j) { Uo[i * 3 + j] = U[i][j]; Vo[i * 3 + j] = V[i][j]; }
    }
}

/* ------------------------------------------------------------------------- */
/* Kabsch / umeyama (no scaling)                                              */
/* ------------------------------------------------------------------------- */

/* P: current source points (ns x 3), Q: target points (nt x 3), corr[i] = j or -1.
 * Writes the 4x4 update (row-major).  Identity when there is no correspondence.
 * Sums run over correspondences in ascending source index (single-thread order of
 * GetRegistrationResultAndCorrespondences). */
static void kabsch_sv(const double *P, int ns, const double *Q, const int *corr, double *Uout, double *Sout);

ORC_API void orc_kabsch(const double *P, int ns, const double *Q, const int *corr, double *Uout) {
    kabsch_sv(P, ns, Q, corr, Uout, NULL);
}

/* Sout (optional): the three singular values of the covariance (0,0,0 without correspondences) */
static void kabsch_sv(const double *P, int ns, const double *Q, const int *corr, double *Uout, double *Sout) {
    for (int i = 0; i < 16; ++i) Uout[i] = (i % 5 == 0) ? 1.0 : 0.0;
    if (Sout) Sout[0] = Sout[1] = Sout[2] = 0.0;
    int c = 0;
    double ms[3] = {0, 0, 0}, md[3] = {0, 0, 0};
    for (int i = 0; i < ns; ++i) {
        int j = corr[i];
        if (j < 0) continue;
        ++c;
        for (int d = 0; d < 3; ++d) { ms[d] += P[3 * i + d]; md[d] += Q[3 * j + d]; }
    }
    if (c == 0) return;
    const double one_over_n = 1.0 / (double)c;
    for (int d = 0; d < 3; ++d) { ms[d] *= one_over_n; md[d] *= one_over_n; }
    double sigma[9] = {0};
    for (int i = 0; i < ns; ++i) {
        int j = corr[i];
        if (j < 0) continue;
        double a[3], b[3];
        for (int d = 0; d < 3; ++d) { a[d] = P[3 * i + d] - ms[d]; b[d] = Q[3 * j + d] - md[d]; }
        for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc) sigma[r * 3 + cc] += b[r] * a[cc];
    }
    for (int i = 0; i < 9; ++i) sigma[i] *= one_over_n;
    double U[9], S[3], V[9];
    orc_svd3(sigma, U, S, V);
    if (Sout) { Sout[0] = S[0]; Sout[1] = S[1]; Sout[2] = S[2]; }
    double detU = U[0] * (U[4] * U[8] - U[5] * U[7]) - U[1] * (U[3] * U[8] - U[5] * U[6]) + U[2] * (U[3] * U[7] - U[4] * U[6]);
    double detV = V[0] * (V[4] * V[8] - V[5] * V[7]) - V[1] * (V[3] * V[8] - V[5] * V[6]) + V[2] * (V[3] * V[7] - V[4] * V[6]);
    double D[3] = {1.0, 1.0, (detU * detV < 0) ? -1.0 : 1.0};
    double R[9];
    for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += U[r * 3 + k] * D[k] * V[cc * 3 + k];
            R[r * 3 + cc] = s;
        }
    for (int r = 0; r < 3; ++r) {
        double rm = 0.0;
        for (int k = 0; k < 3; ++k) rm += R[r * 3 + k] * ms[k];
        Uout[r * 4 + 0] = R[r * 3 + 0];
        Uout[r * 4 + 1] = R[r * 3 + 1];
        Uout[r * 4 + 2] = R[r * 3 + 2];
        Uout[r * 4 + 3] = md[r] - rm;
    }
}

/* ------------------------------------------------------------------------- */
/* nearest neighbour: brute force and exact k-d tree                          */
/* ------------------------------------------------------------------------- */

static inline double sqdist3(const double *a, const double *b) {
    double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    double r = dx * dx;
    r = r + dy * dy;
    r = r + dz * dz;
    return r;
}

/* argmin_j |p - Q_j|^2, lowest j on exact ties.  Returns -1 when nt == 0. */
ORC_API int orc_nn_brute(const double *p, const double *Q, int nt, double *d2out) {
    int best = -1;
    double bd = INFINITY;
    for (int j = 0; j < nt; ++j) {
        double d = sqdist3(p, Q + 3 * j);
        if (d < bd) { bd = d; best = j; }
    }
    *d2out = bd;
    return best;
}

typedef struct {
    int lo, hi;      /* range in perm[] (leaf) */
    int left, right; /* children, -1 for leaf */
    int dim;
    double split;
} kdnode;

typedef struct {
    const double *Q;
    int n;
    int *perm;
    kdnode *nodes;
    int nnodes;
} kdtree;

#define KD_LEAF 10

static void kd_select(const double *Q, int *perm, int lo, int hi, int k, int dim) {
    /* quickselect on perm[lo..hi) so that perm[k] holds the k-th smallest coord */
    while (hi - lo > 1) {
        double pivot = Q[3 * perm[lo + (hi - lo) / 2] + dim];
        int i = lo, j = hi - 1;
        while (i <= j) {
            while (Q[3 * perm[i] + dim] < pivot) ++i;
            while (Q[3 * perm[j] + dim] > pivot) --j;
            if (i <= j) { int t = perm[i]; perm[i] = perm[j]; perm[j] = t; ++i; --j; }
        }
        if (k <= j) hi = j + 1;
        else if (k >= i) lo = i;
        else return;
    }
}

static int kd_build_rec(kdtree *t, int lo, int hi) {
    int id = t->nnodes++;
    kdnode *nd = &t->nodes[id];
    nd->lo = lo; nd->hi = hi; nd->left = nd->right = -1; nd->dim = 0; nd->split = 0.0;
    if (hi - lo <= KD_LEAF) return id;
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = lo; i < hi; ++i)
        for (int d = 0; d < 3; ++d) {
            double v = t->Q[3 * t->perm[i] + d];
            if (v < mn[d]) mn[d] = v;
            if (v > mx[d]) mx[d] = v;
        }
    int dim = 0;
    if (mx[1] - mn[1] > mx[dim] - mn[dim]) dim = 1;
    if (mx[2] - mn[2] > mx[dim] - mn[dim]) dim = 2;
    int mid = lo + (hi - lo) / 2;
    kd_select(t->Q, t->perm, lo, hi, mid, dim);
    double split = t->Q[3 * t->perm[mid] + dim];
    int l = kd_build_rec(t, lo, mid);
    int r = kd_build_rec(t, mid, hi);
    nd = &t->nodes[id]; /* nodes[] is preallocated, pointer stays valid; re-read for clarity */
    nd->dim = dim; nd->split = split; nd->left = l; nd->right = r;
    return id;
}

static int kd_init(kdtree *t, const double *Q, int n) {
    t->Q = Q; t->n = n; t->nnodes = 0; t->perm = NULL; t->nodes = NULL;
    if (n <= 0) return 0;
    t->perm = (int *)malloc(sizeof(int) * (size_t)n);
    t->nodes = (kdnode *)malloc(sizeof(kdnode) * (size_t)(2 * n + 2));
    if (!t->perm || !t->nodes) return -1;
    for (int i = 0; i < n; ++i) t->perm[i] = i;
    kd_build_rec(t, 0, n);
    return 0;
}
static void kd_free(kdtree *t) { free(t->perm); free(t->nodes); t->perm = NULL; t->nodes = NULL; }

static void kd_search_rec(const kdtree *t, int id, const double *p, int *best, double *bd) {
    const kdnode *nd = &t->nodes[id];
    if (nd->left < 0) {
        for (int i = nd->lo; i < nd->hi; ++i) {
            int j = t->perm[i];
            double d = sqdist3(p, t->Q + 3 * j);
            if (d < *bd || (d == *bd && j < *best)) { *bd = d; *best = j; }
        }
        return;
    }
    double diff = p[nd->dim] - nd->split;
    int near = diff < 0 ? nd->left : nd->right;
    int far = diff < 0 ? nd->right : nd->left;
    kd_search_rec(t, near, p, best, bd);
    /* every point on the far side has |dx_dim| >= |diff|, hence d2 >= diff*diff under
     * monotone rounding; visit it unless it is strictly worse (keeps the tie rule exact) */
    if (!(diff * diff > *bd)) kd_search_rec(t, far, p, best, bd);
}

static int kd_nn(const kdtree *t, const double *p, double *d2out) {
    int best = -1;
    double bd = INFINITY;
    if (t->n > 0) kd_search_rec(t, 0, p, &best, &bd);
    *d2out = bd;
    return best;
}

/* batch NN, for tests: idx[i], d2[i] for each of ns query points */
ORC_API void orc_nn_batch(const double *P, int ns, const double *Q, int nt, int use_kdtree, int *idx, double *d2) {
    if (use_kdtree) {
        kdtree t;
        kd_init(&t, Q, nt);
        for (int i = 0; i < ns; ++i) idx[i] = kd_nn(&t, P + 3 * i, &d2[i]);
        kd_free(&t);
    } else {
        for (int i = 0; i < ns; ++i) idx[i] = orc_nn_brute(P + 3 * i, Q, nt, &d2[i]);
    }
}

/* ------------------------------------------------------------------------- */
/* open3d RegistrationICP, point-to-point                                     */
/* ------------------------------------------------------------------------- */

/* Parallel structure of the CPU run.  0 (default): OpenMP over tiles, each tile single-threaded
 * and summed in ascending source index (deterministic; what the parity tests use).
 * 1: the reference's own structure -- tiles one after the other like the Python loop at
 * cluster_icp.py:131, OpenMP over the source points inside the correspondence search like
 * open3d's GetRegistrationResultAndCorrespondences (per-thread partial sums, so the summation
 * order is run dependent, as it is in open3d). */
static int g_inner_par = 0;
ORC_API void orc_set_reference_threading(int on) { g_inner_par = on ? 1 : 0; }

static void correspond(const double *P, int ns, const double *Q, int nt, const kdtree *tree, double r2,
                       int *corr, double *fit, double *rmse) {
    double err2 = 0.0;
    int c = 0;
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : err2, c) schedule(static) if (g_inner_par)
#endif
    for (int i = 0; i < ns; ++i) {
        double d2;
        int j = tree ? kd_nn(tree, P + 3 * i, &d2) : orc_nn_brute(P + 3 * i, Q, nt, &d2);
        if (j >= 0 && d2 < r2) { corr[i] = j; err2 += d2; ++c; }
        else corr[i] = -1;
    }
    if (c == 0) { *fit = 0.0; *rmse = 0.0; }
    else { *fit = (double)c / (double)ns; *rmse = sqrt(err2 / (double)c); }
}

ORC_API int orc_icp_p2p_cond(const double *src, int ns, const double *tgt, int nt, double max_corr, const double *T0,
                             int max_iter, double rel_fit, double rel_rmse, int use_kdtree, double *T_out, int *corr,
                             double *fit_out, double *rmse_out, int *iters_out, double *P_out, double *cond_out);

/* Returns 0, or -1 on bad arguments (open3d raises when max_corr <= 0), -2 on OOM.
 * corr[i]: index into tgt or -1.  P_out (optional): the incrementally updated points. */
ORC_API int orc_icp_p2p(const double *src, int ns, const double *tgt, int nt, double max_corr, const double *T0,
                        int max_iter, double rel_fit, double rel_rmse, int use_kdtree, double *T_out, int *corr,
                        double *fit_out, double *rmse_out, int *iters_out, double *P_out) {
    return orc_icp_p2p_cond(src, ns, tgt, nt, max_corr, T0, max_iter, rel_fit, rel_rmse, use_kdtree, T_out, corr,
                            fit_out, rmse_out, iters_out, P_out, NULL);
}

/* Same, also reporting cond_out = min over the Kabsch fits of sigma_2/sigma_1 of the covariance
 * (test diagnostic: a value near 0 means some fit had a rank<=1 covariance, where the optimal
 * rotation is not unique and every SVD implementation returns a different member). */
ORC_API int orc_icp_p2p_cond(const double *src, int ns, const double *tgt, int nt, double max_corr, const double *T0,
                             int max_iter, double rel_fit, double rel_rmse, int use_kdtree, double *T_out, int *corr,
                             double *fit_out, double *rmse_out, int *iters_out, double *P_out, double *cond_out) {
    double cond = 1.0;
    if (!(max_corr > 0.0) || ns < 0 || nt < 0) return -1;
    double T[16];
    memcpy(T, T0, sizeof T);
    double *P = (double *)malloc(sizeof(double) * 3 * (size_t)(ns > 0 ? ns : 1));
    if (!P) return -2;
    /* geometry::PointCloud pcd = source; if (!init.isIdentity()) pcd.Transform(init);
     * (transforming by an exact identity is a bit-exact no-op, so no special case) */
    orc_transform_pts(T, src, ns, P);
    kdtree tree;
    const kdtree *tp = NULL;
    if (use_kdtree) {
        if (kd_init(&tree, tgt, nt)) { free(P); return -2; }
        tp = &tree;
    }
    const double r2 = max_corr * max_corr;
    double fit, rmse;
    correspond(P, ns, tgt, nt, tp, r2, corr, &fit, &rmse);
    int it = 0;
    for (int i = 0; i < max_iter; ++i) {
        double U[16], Sv[3];
        kabsch_sv(P, ns, tgt, corr, U, Sv);
        if (fit > 0.0) {
            double ratio = Sv[0] > 0.0 ? Sv[1] / Sv[0] : 0.0;
            if (ratio < cond) cond = ratio;
        }
        orc_mat4_mul(U, T, T);
        orc_transform_pts(U, P, ns, P);
        double fit2, rmse2;
        correspond(P, ns, tgt, nt, tp, r2, corr, &fit2, &rmse2);
        int stop = fabs(fit - fit2) < rel_fit && fabs(rmse - rmse2) < rel_rmse;
        fit = fit2; rmse = rmse2;
        it = i + 1;
        if (stop) break;
    }
    if (use_kdtree) kd_free(&tree);
    memcpy(T_out, T, sizeof T);
    if (P_out) memcpy(P_out, P, sizeof(double) * 3 * (size_t)ns);
    free(P);
    *fit_out = fit; *rmse_out = rmse; *iters_out = it;
    if (cond_out) *cond_out = cond;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* outer sweep: cluster_icp.py:118-191                                        */
/* ------------------------------------------------------------------------- */

/* box[0..2] = lo xyz, box[3..5] = hi xyz, as double.  is_f32: pts are float and the
 * arithmetic runs in float (numpy keeps float32 through np.mean / scalar multiply).
 * n == 0 -> empty box (lo=+inf, hi=-inf). */
ORC_API void orc_aabb_box(const void *pts, int n, int is_f32, double scale, double *box) {
    if (n <= 0) {
        for (int d = 0; d < 3; ++d) { box[d] = INFINITY; box[3 + d] = -INFINITY; }
        return;
    }
    if (is_f32) {
        const float *p = (const float *)pts;
        for (int d = 0; d < 3; ++d) {
            float lo = p[d], hi = p[d];
            for (int i = 1; i < n; ++i) {
                float v = p[3 * i + d];
                if (v < lo) lo = v;
                if (v > hi) hi = v;
            }
            volatile float sum = lo + hi;
            float c = sum / 2.0f;
            volatile float size = hi - lo;
            volatile float hs = (float)(0.5 * scale) * size;
            volatile float blo = c - hs, bhi = c + hs;
            box[d] = (double)blo;
            box[3 + d] = (double)bhi;
        }
    } else {
        const double *p = (const double *)pts;
        for (int d = 0; d < 3; ++d) {
            double lo = p[d], hi = p[d];
            for (int i = 1; i < n; ++i) {
                double v = p[3 * i + d];
                if (v < lo) lo = v;
                if (v > hi) hi = v;
            }
            double c = (lo + hi) / 2.0;
            double size = hi - lo;
            double hs = (0.5 * scale) * size;
            box[d] = c - hs;
            box[3 + d] = c + hs;
        }
    }
}

/* order-preserving strict-inside gather; returns the count, writes original indices */
ORC_API int orc_mask(const double *cloud, int M, const double *box, int *idx_out) {
    int c = 0;
    for (int i = 0; i < M; ++i) {
        const double *p = cloud + 3 * i;
        if (p[0] > box[0] && p[0] < box[3] && p[1] > box[1] && p[1] < box[4] && p[2] > box[2] && p[2] < box[5])
            idx_out[c++] = i;
    }
    return c;
}

/* Batched masked_icp over B (frame, cluster) tiles.
 *   src       packed local clusters (sum n_s x 3 f64), src_off[B+1]
 *   tgt       packed frame clouds   (sum M x 3 f64),   tgt_off[F+1], tile_frame[B]
 *   box_pts   packed predicted world clusters (f32 or f64), box_off[B+1]
 *   init_T    B x 16 (f64 values; the reference hands f32-valued matrices, widened exactly)
 *   outputs   out_T B x 16, out_world like src, out_corr (sum n_s; index into the frame's
 *             cloud, -1 = none), out_fit/out_rmse/out_iters/out_ntgt per tile.
 * Returns 0 or the first non-zero status of a tile. */
ORC_API int orc_masked_icp_sweep(const double *src, const int *src_off, const double *tgt, const int *tgt_off,
                                 const int *tile_frame, const void *box_pts, int box_is_f32, const int *box_off,
                                 const double *init_T, int B, double box_scale, double max_corr, int max_iter,
                                 double rel_fit, double rel_rmse, int ori_only, int use_kdtree, int nthreads,
                                 double *out_T, double *out_world, int *out_corr, double *out_fit,
                                 double *out_rmse, int *out_iters, int *out_ntgt, double *out_cond) {
    int status = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) if (!g_inner_par)
#endif
    for (int b = 0; b < B; ++b) {
        const int s0 = src_off[b], ns = src_off[b + 1] - s0;
        const int f = tile_frame[b];
        const int t0 = tgt_off[f], M = tgt_off[f + 1] - t0;
        const int b0 = box_off[b], nb = box_off[b + 1] - b0;
        double box[6];
        const void *bp = box_is_f32 ? (const void *)((const float *)box_pts + 3 * (size_t)b0)
                                    : (const void *)((const double *)box_pts + 3 * (size_t)b0);
        orc_aabb_box(bp, nb, box_is_f32, box_scale, box);
        int *midx = (int *)malloc(sizeof(int) * (size_t)(M > 0 ? M : 1));
        int nt = orc_mask(tgt + 3 * (size_t)t0, M, box, midx);
        double *Q = (double *)malloc(sizeof(double) * 3 * (size_t)(nt > 0 ? nt : 1));
        for (int j = 0; j < nt; ++j)
            for (int d = 0; d < 3; ++d) Q[3 * j + d] = tgt[3 * ((size_t)t0 + midx[j]) + d];
        int *corr = (int *)malloc(sizeof(int) * (size_t)(ns > 0 ? ns : 1));
        double T[16], fit = 0, rmse = 0, cond = 1.0;
        int iters = 0;
        int rc = orc_icp_p2p_cond(src + 3 * (size_t)s0, ns, Q, nt, max_corr, init_T + 16 * (size_t)b, max_iter,
                                  rel_fit, rel_rmse, use_kdtree, T, corr, &fit, &rmse, &iters, NULL, &cond);
        if (rc != 0) {
#ifdef _OPENMP
#pragma omp critical
#endif
            { if (status == 0) status = rc; }
        } else {
            if (ori_only) { /* cluster_icp.py:161-163 */
                T[3] = init_T[16 * (size_t)b + 3];
                T[7] = init_T[16 * (size_t)b + 7];
                T[11] = init_T[16 * (size_t)b + 11];
            }
            memcpy(out_T + 16 * (size_t)b, T, sizeof T);
            orc_transform_pts(T, src + 3 * (size_t)s0, ns, out_world + 3 * (size_t)s0); /* :167 */
            for (int i = 0; i < ns; ++i) out_corr[s0 + i] = corr[i] < 0 ? -1 : midx[corr[i]];
            out_fit[b] = fit; out_rmse[b] = rmse; out_iters[b] = iters; out_ntgt[b] = nt;
            if (out_cond) out_cond[b] = cond;
        }
        free(midx); free(Q); free(corr);
    }
    return status;
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline sets the thread count explicitly */
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
