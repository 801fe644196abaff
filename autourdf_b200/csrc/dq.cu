// dq.cu -- batched dual-quaternion / quaternion SE(3) library (sm_100a).
//
// One kernel, one thread per batch element, for the 11 functions of AutoURDF
// PointCloud/dq_func.py:4-257 and the four pytorch3d 0.7.7 rotation_conversions functions
// they are built on (dq_func.py:2).  Arithmetic runs in the tensor dtype with the reference's
// expression order; compile with -fmad=false so float32 results round like the eager torch
// ops of the reference (each multiply/add rounded on its own).
// Real-first quaternions (w, x, y, z), Hamilton product.  All ops are elementwise and
// HBM-bound (<= 0.5 FLOP/B).
#include <math.h>

#include "common.cuh"

namespace aurdf {

template <typename T>
struct Q4 {
    T w, x, y, z;
};

template <typename T>
__device__ __forceinline__ Q4<T> ldq(const T *p) { return {p[0], p[1], p[2], p[3]}; }
template <typename T>
__device__ __forceinline__ void stq(T *p, const Q4<T> &q) { p[0] = q.w; p[1] = q.x; p[2] = q.y; p[3] = q.z; }

// pytorch3d quaternion_raw_multiply
template <typename T>
__device__ __forceinline__ Q4<T> qmul(const Q4<T> &a, const Q4<T> &b) {
    Q4<T> o;
    o.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    o.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    o.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    o.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    return o;
}
// pytorch3d quaternion_invert: q * (1,-1,-1,-1), no normalisation
template <typename T>
__device__ __forceinline__ Q4<T> qinv(const Q4<T> &q) { return {q.w, -q.x, -q.y, -q.z}; }

// pytorch3d quaternion_to_matrix (row-major 3x3 into R[9])
template <typename T>
__device__ __forceinline__ void q2m(const Q4<T> &q, T *R) {
    const T r = q.w, i = q.x, j = q.y, k = q.z;
    const T two_s = (T)2.0 / (((r * r + i * i) + j * j) + k * k);
    R[0] = (T)1 - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r); R[2] = two_s * (i * k + j * r);
    R[3] = two_s * (i * j + k * r); R[4] = (T)1 - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
    R[6] = two_s * (i * k - j * r); R[7] = two_s * (j * k + i * r); R[8] = (T)1 - two_s * (i * i + j * j);
}

template <typename T>
__device__ __forceinline__ T sqrt_pos(T x) { return x > (T)0 ? sqrt(x) : (T)0; }

// pytorch3d matrix_to_quaternion (+ standardize: non-negative real part)
template <typename T>
__device__ __forceinline__ Q4<T> m2q(const T *m) {
    const T m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
    const T one = (T)1;
    const T qa0 = sqrt_pos(one + m00 + m11 + m22), qa1 = sqrt_pos(one + m00 - m11 - m22);
    const T qa2 = sqrt_pos(one - m00 + m11 - m22), qa3 = sqrt_pos(one - m00 - m11 + m22);
    // argmax, first occurrence
    int best = 0;
    T bv = qa0;
    if (qa1 > bv) { bv = qa1; best = 1; }
    if (qa2 > bv) { bv = qa2; best = 2; }
    if (qa3 > bv) { bv = qa3; best = 3; }
    const T den = (T)2 * (bv > (T)0.1 ? bv : (T)0.1);
    Q4<T> o;
    if (best == 0) o = {qa0 * qa0, m21 - m12, m02 - m20, m10 - m01};
    else if (best == 1) o = {m21 - m12, qa1 * qa1, m10 + m01, m02 + m20};
    else if (best == 2) o = {m02 - m20, m10 + m01, qa2 * qa2, m12 + m21};
    else o = {m10 - m01, m20 + m02, m21 + m12, qa3 * qa3};
    o.w /= den; o.x /= den; o.y /= den; o.z /= den;
    if (o.w < (T)0) { o.w = -o.w; o.x = -o.x; o.y = -o.y; o.z = -o.z; }
    return o;
}

template <typename T>
__device__ __forceinline__ T eps_of();
template <>
__device__ __forceinline__ float eps_of<float>() { return 1.1920928955078125e-07f; }
template <>
__device__ __forceinline__ double eps_of<double>() { return 2.220446049250313e-16; }

// dq_func.py:47-70
template <typename T>
__device__ __forceinline__ void quat_trans_to_dq(const Q4<T> &q, const T *t, T *dq) {
    const Q4<T> qd = {(T)0, t[0], t[1], t[2]};
    const Q4<T> d = qmul(qd, q);
    stq(dq, q);
    dq[4] = (T)0.5 * d.w; dq[5] = (T)0.5 * d.x; dq[6] = (T)0.5 * d.y; dq[7] = (T)0.5 * d.z;
}

// dq_func.py:72-98
template <typename T>
__device__ __forceinline__ void rot_trans_to_dq(const T *R, const T *t, T *dq) {
    Q4<T> q = m2q(R);
    const T n = sqrt(((q.w * q.w + q.x * q.x) + q.y * q.y) + q.z * q.z);
    const T d = n > eps_of<T>() ? n : eps_of<T>();
    q.w /= d; q.x /= d; q.y /= d; q.z /= d;
    quat_trans_to_dq(q, t, dq);
}

// t = 2 * (q_d (x) q_r^-1).xyz   (dq_func.py:145 / :167)
template <typename T>
__device__ __forceinline__ void dq_translation(const T *dq, T *t) {
    const Q4<T> r = ldq(dq), d = ldq(dq + 4);
    const Q4<T> p = qmul(d, qinv(r));
    t[0] = (T)2 * p.x; t[1] = (T)2 * p.y; t[2] = (T)2 * p.z;
}

template <typename T>
__device__ __forceinline__ void write_transform(const T *R, const T *t, T *M) {
    M[0] = R[0]; M[1] = R[1]; M[2] = R[2]; M[3] = t[0];
    M[4] = R[3]; M[5] = R[4]; M[6] = R[5]; M[7] = t[1];
    M[8] = R[6]; M[9] = R[7]; M[10] = R[8]; M[11] = t[2];
    M[12] = (T)0; M[13] = (T)0; M[14] = (T)0; M[15] = (T)1;
}

template <typename T>
__global__ void __launch_bounds__(256)
dq_op_kernel(int op, const T *__restrict__ in0, const T *__restrict__ in1, T *__restrict__ out0,
             T *__restrict__ out1, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    switch (op) {
        case AURDF_DQ_TRANSFORM_FROM_ROT_TRANS: {
            write_transform(in0 + 9 * i, in1 + 3 * i, out0 + 16 * i);
        } break;
        case AURDF_DQ_QUATERNION_CONJUGATE:
        case AURDF_Q_INVERT: {
            stq(out0 + 4 * i, qinv(ldq(in0 + 4 * i)));
        } break;
        case AURDF_DQ_QUAT_TRANS_TO_DUALQUAT: {
            quat_trans_to_dq(ldq(in0 + 4 * i), in1 + 3 * i, out0 + 8 * i);
        } break;
        case AURDF_DQ_ROT_TRANS_TO_DUALQUAT: {
            rot_trans_to_dq(in0 + 9 * i, in1 + 3 * i, out0 + 8 * i);
        } break;
        case AURDF_DQ_TRANSFORM_TO_DUALQUAT: {
            const T *M = in0 + 16 * i;
            const T R[9] = {M[0], M[1], M[2], M[4], M[5], M[6], M[8], M[9], M[10]};
            const T t[3] = {M[3], M[7], M[11]};
            rot_trans_to_dq(R, t, out0 + 8 * i);
        } break;
        case AURDF_DQ_DUALQUAT_TO_QUAT_TRANS: {  // q = q_r (x) q_d, exactly as written upstream (:144)
            const T *dq = in0 + 8 * i;
            stq(out0 + 4 * i, qmul(ldq(dq), ldq(dq + 4)));
            dq_translation(dq, out1 + 3 * i);
        } break;
        case AURDF_DQ_DUALQUAT_TO_ROT_TRANS: {
            const T *dq = in0 + 8 * i;
            q2m(ldq(dq), out0 + 9 * i);
            dq_translation(dq, out1 + 3 * i);
        } break;
        case AURDF_DQ_DUALQUAT_TO_TRANSFORM: {
            const T *dq = in0 + 8 * i;
            T R[9], t[3];
            q2m(ldq(dq), R);
            dq_translation(dq, t);
            write_transform(R, t, out0 + 16 * i);
        } break;
        case AURDF_DQ_DUALQUAT_MULTIPLY: {
            const T *a = in0 + 8 * i, *b = in1 + 8 * i;
            const Q4<T> ar = ldq(a), ad = ldq(a + 4), br = ldq(b), bd = ldq(b + 4);
            stq(out0 + 8 * i, qmul(ar, br));
            const Q4<T> u = qmul(ar, bd), v = qmul(ad, br);
            const Q4<T> d = {u.w + v.w, u.x + v.x, u.y + v.y, u.z + v.z};
            stq(out0 + 8 * i + 4, d);
        } break;
        case AURDF_DQ_DUALQUAT_INVERT: {  // dq_func.py:213-236
            const T *dq = in0 + 8 * i;
            const Q4<T> r = ldq(dq), d = ldq(dq + 4);
            const T nrm = sqrt(((r.w * r.w + r.x * r.x) + r.y * r.y) + r.z * r.z);
            const T n2 = nrm * nrm;
            const T den = n2 > eps_of<T>() ? n2 : eps_of<T>();
            const Q4<T> rc = qinv(r), dc = qinv(d);
            const T dot = (((r.w * d.w + r.x * d.x) + r.y * d.y) + r.z * d.z) / (den * den);
            T *o = out0 + 8 * i;
            o[0] = rc.w / den; o[1] = rc.x / den; o[2] = rc.y / den; o[3] = rc.z / den;
            o[4] = dc.w / den - ((T)2 * rc.w) * dot;
            o[5] = dc.x / den - ((T)2 * rc.x) * dot;
            o[6] = dc.y / den - ((T)2 * rc.y) * dot;
            o[7] = dc.z / den - ((T)2 * rc.z) * dot;
        } break;
        case AURDF_DQ_POINT_TO_DUALQUAT: {
            const T *p = in0 + 3 * i;
            T *o = out0 + 8 * i;
            o[0] = (T)1; o[1] = (T)0; o[2] = (T)0; o[3] = (T)0; o[4] = (T)0; o[5] = p[0]; o[6] = p[1]; o[7] = p[2];
        } break;
        case AURDF_Q_RAW_MULTIPLY: {
            stq(out0 + 4 * i, qmul(ldq(in0 + 4 * i), ldq(in1 + 4 * i)));
        } break;
        case AURDF_Q_TO_MATRIX: {
            q2m(ldq(in0 + 4 * i), out0 + 9 * i);
        } break;
        case AURDF_MATRIX_TO_Q: {
            stq(out0 + 4 * i, m2q(in0 + 9 * i));
        } break;
        default: break;
    }
}

}  // namespace aurdf

using namespace aurdf;

extern "C" int aurdf_dq_op(int op, const void *in0, const void *in1, void *out0, void *out1, int64_t n, int dtype,
                           aurdf_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AURDF_REQUIRE(op >= 0 && op <= AURDF_MATRIX_TO_Q, "aurdf_dq_op: unknown op");
    AURDF_REQUIRE(n >= 0, "aurdf_dq_op: n < 0");
    AURDF_REQUIRE(dtype == AURDF_F32 || dtype == AURDF_F64, "aurdf_dq_op: bad dtype");
    if (n == 0) return AURDF_OK;
    const bool binary = op == AURDF_DQ_TRANSFORM_FROM_ROT_TRANS || op == AURDF_DQ_QUAT_TRANS_TO_DUALQUAT ||
                        op == AURDF_DQ_ROT_TRANS_TO_DUALQUAT || op == AURDF_DQ_DUALQUAT_MULTIPLY ||
                        op == AURDF_Q_RAW_MULTIPLY;
    const bool two_out = op == AURDF_DQ_DUALQUAT_TO_QUAT_TRANS || op == AURDF_DQ_DUALQUAT_TO_ROT_TRANS;
    AURDF_REQUIRE(in0 && out0 && (!binary || in1) && (!two_out || out1), "aurdf_dq_op: NULL pointer");
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (dtype == AURDF_F32)
        dq_op_kernel<float><<<grid, 256, 0, stream>>>(op, (const float *)in0, (const float *)in1, (float *)out0, (float *)out1, n);
    else
        dq_op_kernel<double><<<grid, 256, 0, stream>>>(op, (const double *)in0, (const double *)in1, (double *)out0, (double *)out1, n);
    AURDF_CUDA_CHECK(cudaGetLastError());
    return AURDF_OK;
}
