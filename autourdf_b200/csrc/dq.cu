// dq.cu -- batched dual-quaternion / quaternion SE(3) library (sm_100a), forward and backward.
//
// One kernel, one thread per batch element, for the 11 functions of AutoURDF
// PointCloud/dq_func.py:4-257 and the four pytorch3d 0.7.7 rotation_conversions functions
// they are built on (dq_func.py:2).  Arithmetic runs in the tensor dtype with the reference's
// expression order; compiled with -fmad=false so float32 results round like the eager torch
// ops of the reference (each multiply/add rounded on its own).
// Real-first quaternions (w, x, y, z), Hamilton product.  All ops are elementwise and
// HBM-bound (<= 0.5 FLOP/B).
//
// Backward (aurdf_dq_op_bwd): the reference differentiates through these functions inside
// train() (PointCloud/mlp_reg.py:60-84 -> loss.backward() at :114-116; `--r q` uses
// matrix_to_quaternion / quaternion_to_matrix, `--r dq` transform_to_dualquat /
// dualquat_to_transform).  Every operator body is written once, templated on the scalar type; the
// backward kernel instantiates it with dual numbers (value, derivative) and seeds one input
// component at a time, which gives the exact Jacobian column of the very same expression --
// including the branch the forward pass took (arg-max candidate of matrix_to_quaternion, the
// max(.,0.1) and eps clamps, the sign standardisation), i.e. the subgradient torch autograd uses.
// An element has at most 19 inputs and the batch is the number of clusters, so the 19 passes
// are free.
#include <math.h>

#include "common.cuh"

namespace aurdf {

using ::sqrt;   // the scalar overloads, next to the dual-number one below

// ---- dual numbers ------------------------------------------------------------------------
template <typename T>
struct Dual {
    T v, d;
    __device__ __forceinline__ Dual() : v(0), d(0) {}
    __device__ __forceinline__ Dual(T v_) : v(v_), d(0) {}
    __device__ __forceinline__ Dual(T v_, T d_) : v(v_), d(d_) {}
};
template <typename T> __device__ __forceinline__ Dual<T> operator+(Dual<T> a, Dual<T> b) { return {a.v + b.v, a.d + b.d}; }
template <typename T> __device__ __forceinline__ Dual<T> operator-(Dual<T> a, Dual<T> b) { return {a.v - b.v, a.d - b.d}; }
template <typename T> __device__ __forceinline__ Dual<T> operator-(Dual<T> a) { return {-a.v, -a.d}; }
template <typename T> __device__ __forceinline__ Dual<T> operator*(Dual<T> a, Dual<T> b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
template <typename T> __device__ __forceinline__ Dual<T> operator/(Dual<T> a, Dual<T> b) {
    const T q = a.v / b.v;
    return {q, (a.d - q * b.d) / b.v};
}
template <typename T> __device__ __forceinline__ bool operator>(Dual<T> a, Dual<T> b) { return a.v > b.v; }
template <typename T> __device__ __forceinline__ bool operator<(Dual<T> a, Dual<T> b) { return a.v < b.v; }
template <typename T> __device__ __forceinline__ Dual<T> sqrt(Dual<T> a) {
    const T s = ::sqrt(a.v);
    return {s, a.d / ((T)2 * s)};
}

template <typename S> struct scalar_of { using type = S; };
template <typename T> struct scalar_of<Dual<T>> { using type = T; };

template <typename S>
struct Q4 {
    S w, x, y, z;
};

template <typename S>
__device__ __forceinline__ Q4<S> ldq(const S *p) { return {p[0], p[1], p[2], p[3]}; }
template <typename S>
__device__ __forceinline__ void stq(S *p, const Q4<S> &q) { p[0] = q.w; p[1] = q.x; p[2] = q.y; p[3] = q.z; }

// pytorch3d quaternion_raw_multiply
template <typename S>
__device__ __forceinline__ Q4<S> qmul(const Q4<S> &a, const Q4<S> &b) {
    Q4<S> o;
    o.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    o.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    o.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    o.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    return o;
}
// pytorch3d quaternion_invert: q * (1,-1,-1,-1), no normalisation
template <typename S>
__device__ __forceinline__ Q4<S> qinv(const Q4<S> &q) { return {q.w, -q.x, -q.y, -q.z}; }

// pytorch3d quaternion_to_matrix (row-major 3x3 into R[9])
template <typename S>
__device__ __forceinline__ void q2m(const Q4<S> &q, S *R) {
    const S r = q.w, i = q.x, j = q.y, k = q.z;
    const S one = S(1), two = S(2);
    const S two_s = two / (((r * r + i * i) + j * j) + k * k);
    R[0] = one - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r); R[2] = two_s * (i * k + j * r);
    R[3] = two_s * (i * j + k * r); R[4] = one - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
    R[6] = two_s * (i * k - j * r); R[7] = two_s * (j * k + i * r); R[8] = one - two_s * (i * i + j * j);
}

// pytorch3d _sqrt_positive_part: sqrt(x) where x > 0, else 0 (with a zero subgradient there)
template <typename S>
__device__ __forceinline__ S sqrt_pos(S x) { return x > S(0) ? sqrt(x) : S(0); }

// pytorch3d matrix_to_quaternion (+ standardize: non-negative real part)
template <typename S>
__device__ __forceinline__ Q4<S> m2q(const S *m) {
    const S m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
    const S one = S(1);
    const S qa0 = sqrt_pos(one + m00 + m11 + m22), qa1 = sqrt_pos(one + m00 - m11 - m22);
    const S qa2 = sqrt_pos(one - m00 + m11 - m22), qa3 = sqrt_pos(one - m00 - m11 + m22);
    // argmax, first occurrence
    int best = 0;
    S bv = qa0;
    if (qa1 > bv) { bv = qa1; best = 1; }
    if (qa2 > bv) { bv = qa2; best = 2; }
    if (qa3 > bv) { bv = qa3; best = 3; }
    const S tenth = S((typename scalar_of<S>::type)0.1);
    const S den = S(2) * (bv > tenth ? bv : tenth);
    Q4<S> o;
    if (best == 0) o = {qa0 * qa0, m21 - m12, m02 - m20, m10 - m01};
    else if (best == 1) o = {m21 - m12, qa1 * qa1, m10 + m01, m02 + m20};
    else if (best == 2) o = {m02 - m20, m10 + m01, qa2 * qa2, m12 + m21};
    else o = {m10 - m01, m20 + m02, m21 + m12, qa3 * qa3};
    o.w = o.w / den; o.x = o.x / den; o.y = o.y / den; o.z = o.z / den;
    if (o.w < S(0)) { o.w = -o.w; o.x = -o.x; o.y = -o.y; o.z = -o.z; }
    return o;
}

template <typename T>
__device__ __forceinline__ T eps_of();
template <>
__device__ __forceinline__ float eps_of<float>() { return 1.1920928955078125e-07f; }
template <>
__device__ __forceinline__ double eps_of<double>() { return 2.220446049250313e-16; }

// dq_func.py:47-70
template <typename S>
__device__ __forceinline__ void quat_trans_to_dq(const Q4<S> &q, const S *t, S *dq) {
    const Q4<S> qd = {S(0), t[0], t[1], t[2]};
    const Q4<S> d = qmul(qd, q);
    const S half = S((typename scalar_of<S>::type)0.5);
    stq(dq, q);
    dq[4] = half * d.w; dq[5] = half * d.x; dq[6] = half * d.y; dq[7] = half * d.z;
}

// dq_func.py:72-98
template <typename S>
__device__ __forceinline__ void rot_trans_to_dq(const S *R, const S *t, S *dq) {
    using T = typename scalar_of<S>::type;
    Q4<S> q = m2q(R);
    const S n = sqrt(((q.w * q.w + q.x * q.x) + q.y * q.y) + q.z * q.z);
    const S e = S(eps_of<T>());
    const S d = n > e ? n : e;
    q.w = q.w / d; q.x = q.x / d; q.y = q.y / d; q.z = q.z / d;
    quat_trans_to_dq(q, t, dq);
}

// t = 2 * (q_d (x) q_r^-1).xyz   (dq_func.py:145 / :167)
template <typename S>
__device__ __forceinline__ void dq_translation(const S *dq, S *t) {
    const Q4<S> r = ldq(dq), d = ldq(dq + 4);
    const Q4<S> p = qmul(d, qinv(r));
    t[0] = S(2) * p.x; t[1] = S(2) * p.y; t[2] = S(2) * p.z;
}

template <typename S>
__device__ __forceinline__ void write_transform(const S *R, const S *t, S *M) {
    M[0] = R[0]; M[1] = R[1]; M[2] = R[2]; M[3] = t[0];
    M[4] = R[3]; M[5] = R[4]; M[6] = R[5]; M[7] = t[1];
    M[8] = R[6]; M[9] = R[7]; M[10] = R[8]; M[11] = t[2];
    M[12] = S(0); M[13] = S(0); M[14] = S(0); M[15] = S(1);
}

// element sizes of (in0, in1, out0, out1) per operator
struct DqShape { int i0, i1, o0, o1; };
__host__ __device__ inline DqShape dq_shape(int op) {
    switch (op) {
        case AURDF_DQ_TRANSFORM_FROM_ROT_TRANS: return {9, 3, 16, 0};
        case AURDF_DQ_QUATERNION_CONJUGATE: return {4, 0, 4, 0};
        case AURDF_DQ_QUAT_TRANS_TO_DUALQUAT: return {4, 3, 8, 0};
        case AURDF_DQ_ROT_TRANS_TO_DUALQUAT: return {9, 3, 8, 0};
        case AURDF_DQ_TRANSFORM_TO_DUALQUAT: return {16, 0, 8, 0};
        case AURDF_DQ_DUALQUAT_TO_QUAT_TRANS: return {8, 0, 4, 3};
        case AURDF_DQ_DUALQUAT_TO_ROT_TRANS: return {8, 0, 9, 3};
        case AURDF_DQ_DUALQUAT_TO_TRANSFORM: return {8, 0, 16, 0};
        case AURDF_DQ_DUALQUAT_MULTIPLY: return {8, 8, 8, 0};
        case AURDF_DQ_DUALQUAT_INVERT: return {8, 0, 8, 0};
        case AURDF_DQ_POINT_TO_DUALQUAT: return {3, 0, 8, 0};
        case AURDF_Q_RAW_MULTIPLY: return {4, 4, 4, 0};
        case AURDF_Q_INVERT: return {4, 0, 4, 0};
        case AURDF_Q_TO_MATRIX: return {4, 0, 9, 0};
        case AURDF_MATRIX_TO_Q: return {9, 0, 4, 0};
        default: return {0, 0, 0, 0};
    }
}

// one batch element of operator `op`: a (in0), b (in1) -> o0, o1
template <typename S>
__device__ __forceinline__ void dq_apply(int op, const S *a, const S *b, S *o0, S *o1) {
    using T = typename scalar_of<S>::type;
    switch (op) {
        case AURDF_DQ_TRANSFORM_FROM_ROT_TRANS: {
            write_transform(a, b, o0);
        } break;
        case AURDF_DQ_QUATERNION_CONJUGATE:
        case AURDF_Q_INVERT: {
            stq(o0, qinv(ldq(a)));
        } break;
        case AURDF_DQ_QUAT_TRANS_TO_DUALQUAT: {
            quat_trans_to_dq(ldq(a), b, o0);
        } break;
        case AURDF_DQ_ROT_TRANS_TO_DUALQUAT: {
            rot_trans_to_dq(a, b, o0);
        } break;
        case AURDF_DQ_TRANSFORM_TO_DUALQUAT: {
            const S R[9] = {a[0], a[1], a[2], a[4], a[5], a[6], a[8], a[9], a[10]};
            const S t[3] = {a[3], a[7], a[11]};
            rot_trans_to_dq(R, t, o0);
        } break;
        case AURDF_DQ_DUALQUAT_TO_QUAT_TRANS: {  // q = q_r (x) q_d, exactly as written upstream (:144)
            stq(o0, qmul(ldq(a), ldq(a + 4)));
            dq_translation(a, o1);
        } break;
        case AURDF_DQ_DUALQUAT_TO_ROT_TRANS: {
            q2m(ldq(a), o0);
            dq_translation(a, o1);
        } break;
        case AURDF_DQ_DUALQUAT_TO_TRANSFORM: {
            S R[9], t[3];
            q2m(ldq(a), R);
            dq_translation(a, t);
            write_transform(R, t, o0);
        } break;
        case AURDF_DQ_DUALQUAT_MULTIPLY: {
            const Q4<S> ar = ldq(a), ad = ldq(a + 4), br = ldq(b), bd = ldq(b + 4);
            stq(o0, qmul(ar, br));
            const Q4<S> u = qmul(ar, bd), v = qmul(ad, br);
            const Q4<S> d = {u.w + v.w, u.x + v.x, u.y + v.y, u.z + v.z};
            stq(o0 + 4, d);
        } break;
        case AURDF_DQ_DUALQUAT_INVERT: {  // dq_func.py:213-236
            const Q4<S> r = ldq(a), d = ldq(a + 4);
            const S nrm = sqrt(((r.w * r.w + r.x * r.x) + r.y * r.y) + r.z * r.z);
            const S n2 = nrm * nrm;
            const S e = S(eps_of<T>());
            const S den = n2 > e ? n2 : e;
            const Q4<S> rc = qinv(r), dc = qinv(d);
            const S dot = (((r.w * d.w + r.x * d.x) + r.y * d.y) + r.z * d.z) / (den * den);
            o0[0] = rc.w / den; o0[1] = rc.x / den; o0[2] = rc.y / den; o0[3] = rc.z / den;
            o0[4] = dc.w / den - (S(2) * rc.w) * dot;
            o0[5] = dc.x / den - (S(2) * rc.x) * dot;
            o0[6] = dc.y / den - (S(2) * rc.y) * dot;
            o0[7] = dc.z / den - (S(2) * rc.z) * dot;
        } break;
        case AURDF_DQ_POINT_TO_DUALQUAT: {
            o0[0] = S(1); o0[1] = S(0); o0[2] = S(0); o0[3] = S(0); o0[4] = S(0); o0[5] = a[0]; o0[6] = a[1]; o0[7] = a[2];
        } break;
        case AURDF_Q_RAW_MULTIPLY: {
            stq(o0, qmul(ldq(a), ldq(b)));
        } break;
        case AURDF_Q_TO_MATRIX: {
            q2m(ldq(a), o0);
        } break;
        case AURDF_MATRIX_TO_Q: {
            stq(o0, m2q(a));
        } break;
        default: break;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
dq_op_kernel(int op, const T *__restrict__ in0, const T *__restrict__ in1, T *__restrict__ out0,
             T *__restrict__ out1, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DqShape sh = dq_shape(op);
    T a[16], b[8], o0[16], o1[9];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = k < sh.i0 ? in0[sh.i0 * i + k] : (T)0;
#pragma unroll
    for (int k = 0; k < 8; ++k) b[k] = k < sh.i1 ? in1[sh.i1 * i + k] : (T)0;
    dq_apply<T>(op, a, b, o0, o1);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (k < sh.o0) out0[sh.o0 * i + k] = o0[k];
#pragma unroll
    for (int k = 0; k < 9; ++k)
        if (k < sh.o1) out1[sh.o1 * i + k] = o1[k];
}

// vector-Jacobian product: gin_k = sum_m gout_m d out_m / d in_k, one Jacobian column per dual pass
template <typename T>
__global__ void __launch_bounds__(128)
dq_op_bwd_kernel(int op, const T *__restrict__ in0, const T *__restrict__ in1, const T *__restrict__ gout0,
                 const T *__restrict__ gout1, T *__restrict__ gin0, T *__restrict__ gin1, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DqShape sh = dq_shape(op);
    T a[16], b[8], g0[16], g1[9];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = k < sh.i0 ? in0[sh.i0 * i + k] : (T)0;
#pragma unroll
    for (int k = 0; k < 8; ++k) b[k] = k < sh.i1 ? in1[sh.i1 * i + k] : (T)0;
#pragma unroll
    for (int k = 0; k < 16; ++k) g0[k] = (k < sh.o0 && gout0) ? gout0[sh.o0 * i + k] : (T)0;
#pragma unroll
    for (int k = 0; k < 9; ++k) g1[k] = (k < sh.o1 && gout1) ? gout1[sh.o1 * i + k] : (T)0;
#pragma unroll 1
    for (int e = 0; e < sh.i0 + sh.i1; ++e) {
        Dual<T> da[16], db[8], d0[16], d1[9];
#pragma unroll
        for (int k = 0; k < 16; ++k) da[k] = Dual<T>(a[k], (k == e) ? (T)1 : (T)0);
#pragma unroll
        for (int k = 0; k < 8; ++k) db[k] = Dual<T>(b[k], (k + sh.i0 == e) ? (T)1 : (T)0);
#pragma unroll
        for (int k = 0; k < 16; ++k) d0[k] = Dual<T>();
#pragma unroll
        for (int k = 0; k < 9; ++k) d1[k] = Dual<T>();
        dq_apply<Dual<T>>(op, da, db, d0, d1);
        T s = (T)0;
#pragma unroll
        for (int k = 0; k < 16; ++k) s += k < sh.o0 ? g0[k] * d0[k].d : (T)0;
#pragma unroll
        for (int k = 0; k < 9; ++k) s += k < sh.o1 ? g1[k] * d1[k].d : (T)0;
        if (e < sh.i0) {
            if (gin0) gin0[sh.i0 * i + e] = s;
        } else if (gin1) {
            gin1[sh.i1 * i + (e - sh.i0)] = s;
        }
    }
}

}  // namespace aurdf

using namespace aurdf;

static bool dq_binary(int op) { return dq_shape(op).i1 > 0; }
static bool dq_two_out(int op) { return dq_shape(op).o1 > 0; }

extern "C" int aurdf_dq_op(int op, const void *in0, const void *in1, void *out0, void *out1, int64_t n, int dtype,
                           aurdf_stream_t stream_) {
    aurdf::NvtxRange nvtx_range("aurdf_dq_op");
    cudaStream_t stream = (cudaStream_t)stream_;
    AURDF_REQUIRE(op >= 0 && op <= AURDF_MATRIX_TO_Q, "aurdf_dq_op: unknown op");
    AURDF_REQUIRE(n >= 0, "aurdf_dq_op: n < 0");
    AURDF_REQUIRE(dtype == AURDF_F32 || dtype == AURDF_F64, "aurdf_dq_op: bad dtype");
    if (n == 0) return AURDF_OK;
    AURDF_REQUIRE(in0 && out0 && (!dq_binary(op) || in1) && (!dq_two_out(op) || out1), "aurdf_dq_op: NULL pointer");
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (dtype == AURDF_F32)
        dq_op_kernel<float><<<grid, 256, 0, stream>>>(op, (const float *)in0, (const float *)in1, (float *)out0, (float *)out1, n);
    else
        dq_op_kernel<double><<<grid, 256, 0, stream>>>(op, (const double *)in0, (const double *)in1, (double *)out0, (double *)out1, n);
    AURDF_CUDA_CHECK(cudaGetLastError());
    return AURDF_OK;
}

extern "C" int aurdf_dq_op_bwd(int op, const void *in0, const void *in1, const void *gout0, const void *gout1, void *gin0,
                               void *gin1, int64_t n, int dtype, aurdf_stream_t stream_) {
    aurdf::NvtxRange nvtx_range("aurdf_dq_op_bwd");
    cudaStream_t stream = (cudaStream_t)stream_;
    AURDF_REQUIRE(op >= 0 && op <= AURDF_MATRIX_TO_Q, "aurdf_dq_op_bwd: unknown op");
    AURDF_REQUIRE(n >= 0, "aurdf_dq_op_bwd: n < 0");
    AURDF_REQUIRE(dtype == AURDF_F32 || dtype == AURDF_F64, "aurdf_dq_op_bwd: bad dtype");
    if (n == 0) return AURDF_OK;
    AURDF_REQUIRE(in0 && (!dq_binary(op) || in1), "aurdf_dq_op_bwd: NULL input");
    AURDF_REQUIRE(gout0 || gout1, "aurdf_dq_op_bwd: no output gradient");
    AURDF_REQUIRE(gin0 || gin1, "aurdf_dq_op_bwd: no input gradient requested");
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (dtype == AURDF_F32)
        dq_op_bwd_kernel<float><<<grid, 128, 0, stream>>>(op, (const float *)in0, (const float *)in1, (const float *)gout0,
                                                          (const float *)gout1, (float *)gin0, (float *)gin1, n);
    else
        dq_op_bwd_kernel<double><<<grid, 128, 0, stream>>>(op, (const double *)in0, (const double *)in1, (const double *)gout0,
                                                           (const double *)gout1, (double *)gin0, (double *)gin1, n);
    AURDF_CUDA_CHECK(cudaGetLastError());
    return AURDF_OK;
}
