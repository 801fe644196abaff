/*
 * aurdf.h -- C ABI of the B200-native cluster-registration engine (libaurdf.so).
 *
 * Drop-in boundary for ONE hot path of jl6017/AutoURDF: the per-frame, per-cluster
 * point-to-point ICP sweep and the SE(3) / dual-quaternion helpers around it.
 * Reference files cited below are relative to the AutoURDF repository root.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - unless a function name ends in _host, every pointer is a DEVICE pointer on the
 *     current CUDA device and the call is asynchronous on `stream`;
 *   - the library allocates nothing on the device-pointer entry points: the caller owns
 *     every buffer, including `workspace`;
 *   - return value 0 = success, negative = AURDF_E* below; nothing throws; the message of
 *     the last failure on this thread is available from aurdf_last_error_string();
 *   - points are packed xyz triples (AoS, as numpy (n,3) arrays are), CSR int32 offsets
 *     delimit ragged groups; poses are row-major 4x4 doubles; quaternions are real-first.
 */
#ifndef AURDF_H
#define AURDF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AURDF_VERSION 100

#if defined(__GNUC__)
#define AURDF_API __attribute__((visibility("default")))
#else
#define AURDF_API
#endif

typedef void *aurdf_stream_t; /* a cudaStream_t */

enum {
    AURDF_OK = 0,
    AURDF_EINVAL = -1,     /* bad argument (NULL pointer, negative size, max_corr_dist <= 0 ...) */
    AURDF_EWORKSPACE = -2, /* workspace_bytes smaller than aurdf_icp_workspace_bytes() */
    AURDF_ECUDA = -3,      /* a CUDA runtime call failed */
    AURDF_ECAPACITY = -4,  /* compacted-target capacity exceeded (host path only; see status[]) */
    AURDF_ENOMEM = -5
};

enum { AURDF_F32 = 0, AURDF_F64 = 1 };

AURDF_API int aurdf_version(void);
AURDF_API const char *aurdf_last_error_string(void);

/* ---------------------------------------------------------------------------------------
 * Cluster-ICP sweep: replaces masked_icp(), PointCloud/cluster_icp.py:118-191, including
 * the open3d registration_icp() call at :157-159, batched over B (frame, cluster) tiles.
 *
 * Per tile b (frame f = tile_frame[b]):
 *   box      = AABB of box_xyz[box_off[b]..box_off[b+1]) inflated by box_scale about its
 *              centre (cluster_icp.py:133-140; float32 arithmetic when box_dtype is F32,
 *              as numpy does for the float32 prediction handed over by mlp_reg.py:121);
 *   target   = points of frame f strictly inside the box, original order (:142-148);
 *   ICP      = open3d 0.18 RegistrationICP, point-to-point, float64: init init_T[b],
 *              max_corr_dist, max_iter, |d fitness| < rel_fitness && |d rmse| < rel_rmse;
 *   out_T[b] = fitted pose (translation reset to the init's when ori_only, :161-163);
 *   out_world_xyz = out_T[b] applied to the tile's source points (:167);
 *   out_corr[i]   = index, in frame f's cloud, of the final correspondence of source
 *                   point i, or -1; out_fitness/out_rmse/out_iters/out_ntgt per tile.
 *
 * box_xyz == NULL disables the mask (every point of frame f is a target): that is the plain
 * registration_icp() of link.py:113-117 and Sim/evaluation.py:358-362.
 * pts_dtype is the storage type of src_xyz and tgt_xyz; all arithmetic is float64.
 * tgt_capacity is the number of compacted target points the workspace can hold (the sum
 * over tiles of the masked-target counts, each rounded up to even).  status (device
 * int32[4], optional): [0] = 1 if the capacity was exceeded (outputs are then untouched),
 * [1..2] = low/high 32 bits of the capacity that would have been needed.
 * total_src_points = src_off[n_tiles] (the host knows it; the offsets live on the device).
 * max_src_per_tile: upper bound on any tile's source count (0 = unknown), used only to
 * size shared memory; tiles above the shared-memory bound use the workspace spill area.
 * ------------------------------------------------------------------------------------- */
AURDF_API size_t aurdf_icp_workspace_bytes(int32_t n_tiles, int64_t total_src_points, int64_t tgt_capacity);

AURDF_API int aurdf_icp_sweep(const void *src_xyz, int pts_dtype, const int32_t *src_off,
                    const void *tgt_xyz, const int32_t *tgt_off, const int32_t *tile_frame,
                    const void *box_xyz, int box_dtype, const int32_t *box_off,
                    const double *init_T, int32_t n_tiles, int64_t total_src_points,
                    int32_t max_src_per_tile,
                    double box_scale, double max_corr_dist, int32_t max_iter,
                    double rel_fitness, double rel_rmse, int32_t ori_only,
                    double *out_T, double *out_world_xyz, int32_t *out_corr,
                    double *out_fitness, double *out_rmse, int32_t *out_iters, int32_t *out_ntgt,
                    void *workspace, size_t workspace_bytes, int64_t tgt_capacity,
                    int32_t *status, aurdf_stream_t stream);

/* Number of kernels one aurdf_icp_sweep() call launches (for launch accounting). */
AURDF_API int aurdf_icp_sweep_launches(void);

/* Live timing of the dominant kernel (the fused per-tile ICP: icp_small_kernel for tiles of up to
 * 320 source / 760 target points plus icp_tiles_kernel for the rest): while enabled, every
 * aurdf_icp_sweep() on this thread brackets those two launches with CUDA events on the caller's
 * stream; collect() waits for them, returns the summed duration and launch count, and resets.
 * Do not enable it while the stream is being captured into a CUDA graph. */
AURDF_API int aurdf_icp_profile_enable(int on);
AURDF_API int aurdf_icp_profile_collect(double *total_ms, int32_t *n_launches);

/* ---------------------------------------------------------------------------------------
 * Host-buffer path: same operator with HOST pointers.  A context owns streams, pinned
 * staging and growable device buffers; the call copies inputs host->device, runs the sweep,
 * copies results device->host and returns when they are in place.  This is what a
 * reference-side binding calls from the numpy world of mlp_reg.py:325.  A large batch whose
 * tile_frame is non-decreasing is cut into frame blocks on separate streams so that copies
 * overlap kernels; a single-frame call with pageable buffers is one copy each way.
 * ------------------------------------------------------------------------------------- */
typedef struct aurdf_ctx aurdf_ctx;

AURDF_API int aurdf_ctx_create(int device, aurdf_ctx **out);
AURDF_API void aurdf_ctx_destroy(aurdf_ctx *ctx);

AURDF_API int aurdf_icp_sweep_host(aurdf_ctx *ctx,
                         const void *src_xyz, int pts_dtype, const int32_t *src_off,
                         const void *tgt_xyz, const int32_t *tgt_off, const int32_t *tile_frame,
                         const void *box_xyz, int box_dtype, const int32_t *box_off,
                         const double *init_T, int32_t n_tiles, int32_t n_frames,
                         double box_scale, double max_corr_dist, int32_t max_iter,
                         double rel_fitness, double rel_rmse, int32_t ori_only,
                         double *out_T, double *out_world_xyz, int32_t *out_corr,
                         double *out_fitness, double *out_rmse, int32_t *out_iters, int32_t *out_ntgt);

/* bytes moved host->device / device->host by the last aurdf_icp_sweep_host() on ctx */
AURDF_API void aurdf_ctx_last_copy_bytes(const aurdf_ctx *ctx, int64_t *h2d, int64_t *d2h);

/* ---------------------------------------------------------------------------------------
 * Nearest neighbour under squared L2 (float64 arithmetic, lowest index on exact ties):
 * the correspondence search of open3d GetRegistrationResultAndCorrespondences, exposed on
 * its own for B independent (query group, target group) pairs.
 *   n_queries   = query_off[n_groups] (total query points; sizes the grid)
 *   out_idx[i]  index within the group's target range, -1 if the group has no target
 *   out_d2[i]   squared distance (may be NULL)
 * ------------------------------------------------------------------------------------- */
AURDF_API int aurdf_nn_l2(const void *query_xyz, const int32_t *query_off, const void *target_xyz,
                const int32_t *target_off, int pts_dtype, int32_t n_groups, int64_t n_queries,
                int32_t *out_idx, double *out_d2, aurdf_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * float32 1-NN under L1 (norm 1) or squared L2 (norm 2): pytorch3d 0.7.7 ops.knn_points(K=1), the
 * operator inside loss.chamfer_distance(pred, y, norm=1) at PointCloud/mlp_reg.py:96 and
 * Sim/evaluation.py:81.  Distances accumulate in pytorch3d's order, the first
 * minimum wins.  out_idx is relative to the group's target range (-1 for an empty one).
 * aurdf_nn_f32_bwd is knn_points' backward: grad_p1[i] += g_i * d dist/d p1, grad_p2[idx[i]] -= the
 * same (atomic scatter; both gradient buffers must be zero-initialised by the caller, either may
 * be NULL).  workspace: aurdf_nn_f32_workspace_bytes(n_queries) bytes, 8-byte aligned.
 * ------------------------------------------------------------------------------------- */
AURDF_API size_t aurdf_nn_f32_workspace_bytes(int64_t n_queries);
AURDF_API int aurdf_nn_f32(const float *query_xyz, const int32_t *query_off, const float *target_xyz,
                           const int32_t *target_off, int32_t n_groups, int64_t n_queries,
                           int64_t n_targets, int norm, int32_t *out_idx, float *out_dist,
                           void *workspace, size_t workspace_bytes, aurdf_stream_t stream);
AURDF_API int aurdf_nn_f32_bwd(const float *p1_xyz, const int32_t *p1_off, const float *p2_xyz,
                               const int32_t *p2_off, const int32_t *idx, const float *grad_dist,
                               int32_t n_groups, int64_t n_p1, int norm, float *grad_p1,
                               float *grad_p2, aurdf_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Fused chamfer distance: pytorch3d 0.7.7 loss.chamfer_distance(x, y, norm=1|2) exactly as the reference calls
 * it (PointCloud/mlp_reg.py:96 inside the 300-epoch train() loop, ~600 calls per frame at P ~ 5000;
 * Sim/evaluation.py:81) -- both nearest-neighbour directions, the point / batch reductions and the scalar loss
 * in ONE kernel launch; the backward (knn_points' backward for both directions) in one more.
 *   x (N, P1, 3), y (N, P2, 3) contiguous float32;  point_mean / batch_mean: 1 = "mean", 0 = "sum";
 *   idx_x (N*P1) / idx_y (N*P2): nearest neighbour of every point in the other cloud (kept for the backward);
 *   loss: one float (device);  grad_loss: one float (device), the upstream gradient;
 *   grad_x / grad_y: zero-initialised by the caller, either may be NULL.
 * workspace: aurdf_chamfer_workspace_bytes(N, P1, P2) bytes, 256-byte aligned, initialised ONCE with
 * aurdf_chamfer_workspace_init (arrival counters; the kernel resets them itself) and then reused by every call
 * of the same shape; one workspace must not be used by two calls that may run concurrently.
 * ------------------------------------------------------------------------------------- */
AURDF_API size_t aurdf_chamfer_workspace_bytes(int32_t n_batch, int32_t p1, int32_t p2);
AURDF_API int aurdf_chamfer_workspace_init(void *workspace, size_t workspace_bytes, aurdf_stream_t stream);
AURDF_API int aurdf_chamfer_fwd(const float *x, const float *y, int32_t n_batch, int32_t p1, int32_t p2, int norm,
                                int point_mean, int batch_mean, int32_t *idx_x, int32_t *idx_y, float *loss,
                                void *workspace, size_t workspace_bytes, aurdf_stream_t stream);
AURDF_API int aurdf_chamfer_bwd(const float *x, const float *y, const int32_t *idx_x, const int32_t *idx_y,
                                const float *grad_loss, int32_t n_batch, int32_t p1, int32_t p2, int norm,
                                int point_mean, int batch_mean, float *grad_x, float *grad_y, aurdf_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * SE(3) apply: calculate_pc(), PointCloud/mlp_reg.py:155-170:  out = X @ R_k^T + t_k for
 * every point of group k (dtype F32 or F64 for points, poses and output alike).
 * aurdf_se3_apply_bwd is its adjoint for autograd: grad_X = g R_k, grad_T[k][:3,:3] =
 * sum g^T x, grad_T[k][:3,3] = sum g, bottom row 0 (fully overwritten; grad_xyz or grad_T may
 * be NULL to skip that half).
 * aurdf_se3_to_local: inv(T_k) @ [X;1], mlp_reg.py:211-213 and cluster_icp.py:96-98
 * (general 4x4 inverse, float64).
 * ------------------------------------------------------------------------------------- */
AURDF_API int aurdf_se3_apply(const void *xyz, const int32_t *off, const void *T, int32_t n_groups,
                    int64_t n_points, int dtype, void *out_xyz, aurdf_stream_t stream);
AURDF_API int aurdf_se3_apply_bwd(const void *grad_out, const void *xyz, const int32_t *off, const void *T,
                        int32_t n_groups, int64_t n_points, int dtype, void *grad_xyz, void *grad_T,
                        aurdf_stream_t stream);
AURDF_API int aurdf_se3_to_local(const double *xyz, const int32_t *off, const double *T, int32_t n_groups,
                       int64_t n_points, double *out_xyz, aurdf_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Cluster re-sampling: resample_cluster(), PointCloud/mlp_reg.py:172-237 (normal=False): Lloyd
 * k-means of every frame's cloud seeded with the fitted cluster origins matrices[:, :3, 3]
 * (sklearn.cluster.k_means(init=..., n_init=1), :204), then every cluster moved into its local
 * frame with inv(matrices[k]) (:207-213).  SURVEY section 8(f)-2.  float64.
 *   cloud_xyz/cloud_off  packed frame clouds (F frames);   matrices  F x K x 16
 *   out_labels (per point), out_centers F x K x 3, out_local_xyz packed like cloud_xyz but each
 *   frame's points grouped by cluster in cloud order, out_local_off F x (K+1) group offsets
 *   within the frame, out_n_iter / out_inertia per frame (may be NULL).
 * ------------------------------------------------------------------------------------- */
AURDF_API int aurdf_resample_clusters(const double *cloud_xyz, const int32_t *cloud_off,
                                      const double *matrices, int32_t n_frames, int32_t n_clusters,
                                      int32_t max_points_per_frame, int32_t max_iter, double tol,
                                      int32_t *out_labels, double *out_centers, double *out_local_xyz,
                                      int32_t *out_local_off, int32_t *out_n_iter, double *out_inertia,
                                      aurdf_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Dual-quaternion library: PointCloud/dq_func.py:4-257 and the four pytorch3d 0.7.7
 * rotation_conversions functions it imports (dq_func.py:2).  n = number of batch elements,
 * dtype F32 or F64, arithmetic in that dtype in the reference's operation order.
 * ------------------------------------------------------------------------------------- */
enum {
    AURDF_DQ_TRANSFORM_FROM_ROT_TRANS = 0, /* (R 3x3, t 3)      -> T 4x4        dq_func.py:4   */
    AURDF_DQ_QUATERNION_CONJUGATE = 1,     /* q 4               -> q 4          :29            */
    AURDF_DQ_QUAT_TRANS_TO_DUALQUAT = 2,   /* (q 4, t 3)        -> dq 8         :47            */
    AURDF_DQ_ROT_TRANS_TO_DUALQUAT = 3,    /* (R 3x3, t 3)      -> dq 8         :72            */
    AURDF_DQ_TRANSFORM_TO_DUALQUAT = 4,    /* T 4x4             -> dq 8         :100           */
    AURDF_DQ_DUALQUAT_TO_QUAT_TRANS = 5,   /* dq 8              -> (q 4, t 3)   :126           */
    AURDF_DQ_DUALQUAT_TO_ROT_TRANS = 6,    /* dq 8              -> (R 3x3, t 3) :148           */
    AURDF_DQ_DUALQUAT_TO_TRANSFORM = 7,    /* dq 8              -> T 4x4        :170           */
    AURDF_DQ_DUALQUAT_MULTIPLY = 8,        /* (dq 8, dq 8)      -> dq 8         :188           */
    AURDF_DQ_DUALQUAT_INVERT = 9,          /* dq 8              -> dq 8         :213           */
    AURDF_DQ_POINT_TO_DUALQUAT = 10,       /* p 3               -> dq 8         :238           */
    AURDF_Q_RAW_MULTIPLY = 11,             /* (q 4, q 4)        -> q 4   pytorch3d             */
    AURDF_Q_INVERT = 12,                   /* q 4               -> q 4   pytorch3d             */
    AURDF_Q_TO_MATRIX = 13,                /* q 4               -> R 3x3 pytorch3d             */
    AURDF_MATRIX_TO_Q = 14                 /* R 3x3             -> q 4   pytorch3d             */
};
/* in0/in1: inputs (in1 NULL for unary ops); out0/out1: outputs (out1 NULL unless the op
 * has two).  Contiguous, batch-major. */
AURDF_API int aurdf_dq_op(int op, const void *in0, const void *in1, void *out0, void *out1, int64_t n,
                int dtype, aurdf_stream_t stream);
/* Vector-Jacobian product of the same operator (what loss.backward() needs at
 * PointCloud/mlp_reg.py:114-116 when train() runs with --r q / --r dq, :60-84):
 * gin_k = sum_m gout_m * d out_m / d in_k, evaluated on the branch the forward pass took.
 * gout0/gout1: gradients w.r.t. out0/out1 (either may be NULL = zero); gin0/gin1: gradients
 * w.r.t. in0/in1 (either may be NULL = not wanted; gin1 is ignored for unary ops). */
AURDF_API int aurdf_dq_op_bwd(int op, const void *in0, const void *in1, const void *gout0, const void *gout1,
                    void *gin0, void *gin1, int64_t n, int dtype, aurdf_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Pairwise cluster motion-distance map: replaces CoordMap.coord_dist_map,
 * PointCloud/coord_map.py:230-307 (T x K x K Python loops, one roma call per element).
 *   matrices   (n_frames, n_coords, 4, 4) row-major doubles: the stacked matrix/{t:04}.npy files
 *              the registration loop writes (mlp_reg.py:331)
 *   diff != 0  motion between consecutive frames (n_frames - 1 steps, :253-286); diff == 0: poses
 *              themselves (:287-302)
 *   out_map    (n_coords, n_coords, steps) doubles, step fastest (np.stack(..., axis=2), :304)
 *   out_sum    (n_coords, n_coords): sum over steps of |out_map| (:305)
 * n_coords <= 384. workspace: aurdf_coord_dist_map_workspace_bytes() bytes, 256-byte aligned. */
AURDF_API size_t aurdf_coord_dist_map_workspace_bytes(int32_t n_frames, int32_t n_coords, int32_t diff);
AURDF_API int aurdf_coord_dist_map(const double *matrices, int32_t n_frames, int32_t n_coords, double bounding_box,
                         int32_t diff, double *out_map, double *out_sum, void *workspace, size_t workspace_bytes,
                         aurdf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AURDF_H */
