// host_api.cu -- host-buffer entry point of the cluster-ICP sweep.
//
// aurdf_icp_sweep_host() is what a reference-side binding calls from the numpy world of
// AutoURDF PointCloud/mlp_reg.py:325: all pointers are HOST pointers.  A context owns one
// stream, one pinned staging buffer per direction and growable device buffers, so a call is
//   pack inputs into pinned memory -> ONE async H2D copy -> 4 kernels -> ONE async D2H copy
//   -> stream synchronise -> unpack.
// The compacted-target capacity is guessed from the previous call and the call is re-run
// once (inputs already resident) if the guess was too small.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

struct aurdf_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    void *d_in = nullptr;  size_t d_in_bytes = 0;
    void *d_out = nullptr; size_t d_out_bytes = 0;
    void *d_ws = nullptr;  size_t d_ws_bytes = 0;
    void *h_in = nullptr;  size_t h_in_bytes = 0;
    void *h_out = nullptr; size_t h_out_bytes = 0;
    int64_t cap_hint = 0;
    int64_t last_h2d = 0, last_d2h = 0;
};

namespace {
using aurdf::align_up;

int grow_dev(void **p, size_t *have, size_t need) {
    if (need <= *have) return AURDF_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    need = align_up(need + need / 4, 1 << 20);
    cudaError_t e = cudaMalloc(p, need);
    if (e != cudaSuccess) { aurdf::set_error("cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e)); return AURDF_ENOMEM; }
    *have = need;
    return AURDF_OK;
}
// true when a host pointer is page-locked (cudaMallocHost / cudaHostRegister): it can be the
// source or destination of an asynchronous copy without going through the staging buffer
bool is_pinned(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

int grow_pinned(void **p, size_t *have, size_t need) {
    if (need <= *have) return AURDF_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *have = 0;
    need = align_up(need + need / 4, 1 << 20);
    cudaError_t e = cudaMallocHost(p, need);
    if (e != cudaSuccess) { aurdf::set_error("cudaMallocHost(%zu) failed: %s", need, cudaGetErrorString(e)); return AURDF_ENOMEM; }
    *have = need;
    return AURDF_OK;
}
}  // namespace

extern "C" int aurdf_ctx_create(int device, aurdf_ctx **out) {
    AURDF_REQUIRE(out != nullptr, "aurdf_ctx_create: NULL out");
    AURDF_CUDA_CHECK(cudaSetDevice(device));
    aurdf_ctx *c = new aurdf_ctx();
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return aurdf::cuda_fail(e, "cudaStreamCreateWithFlags"); }
    *out = c;
    return AURDF_OK;
}

extern "C" void aurdf_ctx_destroy(aurdf_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->d_in) cudaFree(c->d_in);
    if (c->d_out) cudaFree(c->d_out);
    if (c->d_ws) cudaFree(c->d_ws);
    if (c->h_in) cudaFreeHost(c->h_in);
    if (c->h_out) cudaFreeHost(c->h_out);
    delete c;
}

extern "C" void aurdf_ctx_last_copy_bytes(const aurdf_ctx *c, int64_t *h2d, int64_t *d2h) {
    if (h2d) *h2d = c ? c->last_h2d : 0;
    if (d2h) *d2h = c ? c->last_d2h : 0;
}

extern "C" int aurdf_icp_sweep_host(aurdf_ctx *c, const void *src_xyz, int pts_dtype, const int32_t *src_off,
                                    const void *tgt_xyz, const int32_t *tgt_off, const int32_t *tile_frame,
                                    const void *box_xyz, int box_dtype, const int32_t *box_off, const double *init_T,
                                    int32_t n_tiles, int32_t n_frames, double box_scale, double max_corr_dist,
                                    int32_t max_iter, double rel_fitness, double rel_rmse, int32_t ori_only,
                                    double *out_T, double *out_world_xyz, int32_t *out_corr, double *out_fitness,
                                    double *out_rmse, int32_t *out_iters, int32_t *out_ntgt) {
    AURDF_REQUIRE(c != nullptr, "aurdf_icp_sweep_host: NULL ctx");
    AURDF_REQUIRE(n_tiles >= 0 && n_frames >= 0, "aurdf_icp_sweep_host: negative size");
    if (n_tiles == 0) { c->last_h2d = c->last_d2h = 0; return AURDF_OK; }
    AURDF_REQUIRE(max_corr_dist > 0.0, "aurdf_icp_sweep_host: max_corr_dist must be > 0 (open3d raises)");
    AURDF_REQUIRE(pts_dtype == AURDF_F32 || pts_dtype == AURDF_F64, "aurdf_icp_sweep_host: bad pts_dtype");
    AURDF_REQUIRE(box_dtype == AURDF_F32 || box_dtype == AURDF_F64, "aurdf_icp_sweep_host: bad box_dtype");
    AURDF_REQUIRE(src_off && tgt_off && tile_frame && init_T, "aurdf_icp_sweep_host: NULL input");
    AURDF_REQUIRE(box_xyz == nullptr || box_off != nullptr, "aurdf_icp_sweep_host: box_xyz without box_off");
    AURDF_REQUIRE(out_T && out_world_xyz && out_corr && out_fitness && out_rmse && out_iters && out_ntgt,
                  "aurdf_icp_sweep_host: NULL output");
    AURDF_CUDA_CHECK(cudaSetDevice(c->device));

    const size_t psz = pts_dtype == AURDF_F32 ? 4 : 8, bsz = box_dtype == AURDF_F32 ? 4 : 8;
    const int64_t n_src = src_off[n_tiles], n_tgt = tgt_off[n_frames], n_box = box_xyz ? box_off[n_tiles] : 0;
    int max_src = 0;
    int64_t cap_upper = 0;
    for (int b = 0; b < n_tiles; ++b) {
        const int ns = src_off[b + 1] - src_off[b];
        if (ns > max_src) max_src = ns;
        const int f = tile_frame[b];
        AURDF_REQUIRE(f >= 0 && f < n_frames, "aurdf_icp_sweep_host: tile_frame out of range");
        cap_upper += (int64_t)(tgt_off[f + 1] - tgt_off[f]) + 1;
    }
    AURDF_REQUIRE((n_src == 0 || src_xyz) && (n_tgt == 0 || tgt_xyz), "aurdf_icp_sweep_host: NULL points");

    // ---- input staging layout (256-byte aligned sections) ----
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t o_src = take((size_t)n_src * 3 * psz), o_tgt = take((size_t)n_tgt * 3 * psz);
    const size_t o_box = take((size_t)n_box * 3 * bsz);
    const size_t o_soff = take((size_t)(n_tiles + 1) * 4), o_toff = take((size_t)(n_frames + 1) * 4);
    const size_t o_tf = take((size_t)n_tiles * 4), o_boff = take((size_t)(n_tiles + 1) * 4);
    const size_t o_init = take((size_t)n_tiles * 16 * 8);
    const size_t in_bytes = o;
    // ---- output layout ----
    o = 0;
    const size_t q_T = take((size_t)n_tiles * 16 * 8), q_world = take((size_t)n_src * 3 * 8);
    const size_t q_corr = take((size_t)n_src * 4), q_fit = take((size_t)n_tiles * 8), q_rmse = take((size_t)n_tiles * 8);
    const size_t q_it = take((size_t)n_tiles * 4), q_nt = take((size_t)n_tiles * 4), q_status = take(16);
    const size_t out_bytes = o;

    int rc;
    if ((rc = grow_pinned(&c->h_in, &c->h_in_bytes, in_bytes)) != AURDF_OK) return rc;
    if ((rc = grow_pinned(&c->h_out, &c->h_out_bytes, out_bytes)) != AURDF_OK) return rc;
    if ((rc = grow_dev(&c->d_in, &c->d_in_bytes, in_bytes)) != AURDF_OK) return rc;
    if ((rc = grow_dev(&c->d_out, &c->d_out_bytes, out_bytes)) != AURDF_OK) return rc;

    // host -> device: pinned caller buffers are copied straight from where they are; pageable
    // ones are packed into the context's pinned staging buffer first
    char *hi = (char *)c->h_in;
    char *di = (char *)c->d_in;
    c->last_h2d = 0;
    c->last_d2h = 0;
    auto h2d = [&](size_t off, const void *src, size_t bytes) -> int {
        if (!bytes) return AURDF_OK;
        const void *from = src;
        if (!is_pinned(src)) {
            memcpy(hi + off, src, bytes);
            from = hi + off;
        }
        AURDF_CUDA_CHECK(cudaMemcpyAsync(di + off, from, bytes, cudaMemcpyHostToDevice, c->stream));
        c->last_h2d += (int64_t)bytes;
        return AURDF_OK;
    };
    if ((rc = h2d(o_src, src_xyz, (size_t)n_src * 3 * psz)) != AURDF_OK) return rc;
    if ((rc = h2d(o_tgt, tgt_xyz, (size_t)n_tgt * 3 * psz)) != AURDF_OK) return rc;
    if ((rc = h2d(o_box, box_xyz, (size_t)n_box * 3 * bsz)) != AURDF_OK) return rc;
    if ((rc = h2d(o_soff, src_off, (size_t)(n_tiles + 1) * 4)) != AURDF_OK) return rc;
    if ((rc = h2d(o_toff, tgt_off, (size_t)(n_frames + 1) * 4)) != AURDF_OK) return rc;
    if ((rc = h2d(o_tf, tile_frame, (size_t)n_tiles * 4)) != AURDF_OK) return rc;
    if (box_xyz && (rc = h2d(o_boff, box_off, (size_t)(n_tiles + 1) * 4)) != AURDF_OK) return rc;
    if ((rc = h2d(o_init, init_T, (size_t)n_tiles * 16 * 8)) != AURDF_OK) return rc;

    int64_t cap = c->cap_hint > 0 ? c->cap_hint : 4 * n_src + 2 * (int64_t)n_tiles + 1024;
    if (cap > cap_upper) cap = cap_upper;
    if (cap < 2) cap = 2;
    char *d_o = (char *)c->d_out;
    char *ho = (char *)c->h_out;
    // device -> host targets: straight into pinned caller buffers, through staging otherwise
    struct Out { void *dst; size_t off, bytes; bool direct; };
    Out outs[7] = {{out_T, q_T, (size_t)n_tiles * 16 * 8, false},     {out_world_xyz, q_world, (size_t)n_src * 3 * 8, false},
                   {out_corr, q_corr, (size_t)n_src * 4, false},       {out_fitness, q_fit, (size_t)n_tiles * 8, false},
                   {out_rmse, q_rmse, (size_t)n_tiles * 8, false},     {out_iters, q_it, (size_t)n_tiles * 4, false},
                   {out_ntgt, q_nt, (size_t)n_tiles * 4, false}};
    for (Out &o_ : outs) o_.direct = o_.bytes && is_pinned(o_.dst);
    for (int attempt = 0; attempt < 2; ++attempt) {
        const size_t ws = aurdf_icp_workspace_bytes(n_tiles, n_src, cap);
        if ((rc = grow_dev(&c->d_ws, &c->d_ws_bytes, ws)) != AURDF_OK) return rc;
        rc = aurdf_icp_sweep(di + o_src, pts_dtype, (const int32_t *)(di + o_soff), di + o_tgt,
                             (const int32_t *)(di + o_toff), (const int32_t *)(di + o_tf),
                             box_xyz ? (const void *)(di + o_box) : nullptr, box_dtype,
                             box_xyz ? (const int32_t *)(di + o_boff) : nullptr, (const double *)(di + o_init), n_tiles,
                             n_src, max_src, box_scale, max_corr_dist, max_iter, rel_fitness, rel_rmse, ori_only,
                             (double *)(d_o + q_T), (double *)(d_o + q_world), (int32_t *)(d_o + q_corr),
                             (double *)(d_o + q_fit), (double *)(d_o + q_rmse), (int32_t *)(d_o + q_it),
                             (int32_t *)(d_o + q_nt), c->d_ws, c->d_ws_bytes, cap, (int32_t *)(d_o + q_status), c->stream);
        if (rc != AURDF_OK) return rc;
        // optimistic: queue the status and every output behind the kernels, synchronise once;
        // if the capacity guess was too small the copies are simply repeated after the re-run
        AURDF_CUDA_CHECK(cudaMemcpyAsync(ho + q_status, d_o + q_status, 16, cudaMemcpyDeviceToHost, c->stream));
        c->last_d2h += 16;
        for (Out &o_ : outs) {
            if (!o_.bytes) continue;
            AURDF_CUDA_CHECK(cudaMemcpyAsync(o_.direct ? o_.dst : (void *)(ho + o_.off), d_o + o_.off, o_.bytes,
                                             cudaMemcpyDeviceToHost, c->stream));
            c->last_d2h += (int64_t)o_.bytes;
        }
        AURDF_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        const int32_t *st = (const int32_t *)(ho + q_status);
        const int64_t need = (int64_t)(uint32_t)st[1] | ((int64_t)st[2] << 32);
        if (!st[0]) {
            c->cap_hint = need + need / 8 + 64;  // next call of the same shape fits first time
            break;
        }
        if (attempt == 1) {
            aurdf::set_error("aurdf_icp_sweep_host: compacted-target capacity %lld still too small (need %lld)",
                             (long long)cap, (long long)need);
            return AURDF_ECAPACITY;
        }
        cap = need;
        c->last_d2h = 0;
    }
    for (Out &o_ : outs)
        if (o_.bytes && !o_.direct) memcpy(o_.dst, ho + o_.off, o_.bytes);
    return AURDF_OK;
}
