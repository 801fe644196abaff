// host_api.cu -- host-buffer entry point of the cluster-ICP sweep.
//
// aurdf_icp_sweep_host() is what a reference-side binding calls from the numpy world of
// AutoURDF PointCloud/mlp_reg.py:325: all pointers are HOST pointers.  A context owns streams,
// one pinned staging buffer per direction and growable device buffers, so a call is
//   pack inputs into pinned memory -> async H2D copies -> kernels -> async D2H copies
//   -> stream synchronise -> unpack,
// cut into up to 3 contiguous frame blocks on separate streams so that the copies of one block
// overlap the kernels of another (AURDF_HOST_CHUNKS=1 restores the single-block behaviour).
// The host thread is the scarce resource of this path (every driver call costs 2-4 us), so the
// call issues as few of them as it can: the small per-tile inputs (offsets, tile_frame, initial
// poses) go up in ONE copy, the per-tile outputs of a block come back in ONE copy (the device
// layout is block-major for that), only the three point arrays and the two per-point outputs
// travel on their own -- straight from / to page-locked caller memory when it is page-locked.
// out_world_xyz / out_corr may be NULL: a caller that only wants the poses skips 85 % of the
// device-to-host bytes.  The compacted-target capacity is guessed from the previous call and a
// block is re-run once (inputs already resident) if the guess was too small.
// A context serialises its callers with a mutex (ctypes releases the GIL around the call).
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

struct aurdf_ctx {
    static constexpr int kMaxChunks = 4;
    int device = 0;
    cudaStream_t streams[kMaxChunks] = {nullptr, nullptr, nullptr, nullptr};
    void *d_in = nullptr;  size_t d_in_bytes = 0;
    void *d_out = nullptr; size_t d_out_bytes = 0;
    void *d_ws = nullptr;  size_t d_ws_bytes = 0;
    void *h_in = nullptr;  size_t h_in_bytes = 0;
    void *h_out = nullptr; size_t h_out_bytes = 0;
    int64_t cap_hint[kMaxChunks] = {0, 0, 0, 0};   // compacted-target capacity that fitted last time, per chunk
    int cap_chunks = 0, cap_tiles = 0;             // the call shape those hints belong to
    int64_t last_h2d = 0, last_d2h = 0;
    cudaEvent_t meta_ready = nullptr;              // the packed per-tile inputs have arrived (stream 0)
    std::mutex lock;
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
    cudaError_t e = cudaStreamCreateWithFlags(&c->streams[0], cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return aurdf::cuda_fail(e, "cudaStreamCreateWithFlags"); }
    *out = c;
    return AURDF_OK;
}

extern "C" void aurdf_ctx_destroy(aurdf_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (cudaStream_t st : c->streams)
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    if (c->meta_ready) cudaEventDestroy(c->meta_ready);
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

namespace {
// AURDF_HOST_CHUNKS: frame blocks of a large batch (default 3), read once
int host_chunks_wanted() {
    static const int want = [] {
        const char *e = getenv("AURDF_HOST_CHUNKS");
        int w = e ? atoi(e) : 3;
        if (w < 1) w = 1;
        if (w > aurdf_ctx::kMaxChunks) w = aurdf_ctx::kMaxChunks;
        return w;
    }();
    return want;
}
}  // namespace

static int sweep_host_locked(aurdf_ctx *c, const void *src_xyz, int pts_dtype, const int32_t *src_off,
                             const void *tgt_xyz, const int32_t *tgt_off, const int32_t *tile_frame,
                             const void *box_xyz, int box_dtype, const int32_t *box_off, const double *init_T,
                             int32_t n_tiles, int32_t n_frames, double box_scale, double max_corr_dist,
                             int32_t max_iter, double rel_fitness, double rel_rmse, int32_t ori_only,
                             double *out_T, double *out_world_xyz, int32_t *out_corr, double *out_fitness,
                             double *out_rmse, int32_t *out_iters, int32_t *out_ntgt) {
    const size_t psz = pts_dtype == AURDF_F32 ? 4 : 8, bsz = box_dtype == AURDF_F32 ? 4 : 8;
    const int64_t n_src = src_off[n_tiles], n_tgt = tgt_off[n_frames], n_box = box_xyz ? box_off[n_tiles] : 0;
    bool sorted = true;
    for (int b = 0; b < n_tiles; ++b) {
        const int f = tile_frame[b];
        AURDF_REQUIRE(f >= 0 && f < n_frames, "aurdf_icp_sweep_host: tile_frame out of range");
        if (b && f < tile_frame[b - 1]) sorted = false;
    }
    AURDF_REQUIRE((n_src == 0 || src_xyz) && (n_tgt == 0 || tgt_xyz), "aurdf_icp_sweep_host: NULL points");

    // ---- chunks: contiguous frame blocks, one stream each (tiles are frame-major in every caller of this
    // path; if they are not, or the batch is small, the call runs as a single block)
    int n_chunks = 1;
    {
        const int want = host_chunks_wanted();
        if (sorted && n_tiles >= 64 * want && n_frames >= 2 * want) n_chunks = want;
    }
    int t_lo[aurdf_ctx::kMaxChunks + 1];   // tile range of every chunk (whole frames)
    t_lo[0] = 0;
    for (int k = 1; k < n_chunks; ++k) {
        int t = (int)((int64_t)n_tiles * k / n_chunks);
        while (t < n_tiles && t > 0 && tile_frame[t] == tile_frame[t - 1]) ++t;   // do not split a frame
        t_lo[k] = t < t_lo[k - 1] ? t_lo[k - 1] : t;
    }
    t_lo[n_chunks] = n_tiles;

    // ---- device / staging layout of the inputs (256-byte aligned sections): points, then the packed meta block
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t o_src = take((size_t)n_src * 3 * psz), o_tgt = take((size_t)n_tgt * 3 * psz);
    const size_t o_box = take((size_t)n_box * 3 * bsz);
    const size_t o_meta = o;
    const size_t o_soff = take((size_t)(n_tiles + 1) * 4), o_toff = take((size_t)(n_frames + 1) * 4);
    const size_t o_tf = take((size_t)n_tiles * 4), o_boff = take((size_t)(n_tiles + 1) * 4);
    const size_t o_init = take((size_t)n_tiles * 16 * 8);
    const size_t in_bytes = o, meta_bytes = o - o_meta;
    // ---- outputs: per-point arrays, then one block per chunk holding its per-tile outputs back to back
    o = 0;
    const size_t q_world = take((size_t)n_src * 3 * 8), q_corr = take((size_t)n_src * 4);
    struct Blk { size_t base, T, fit, rmse, it, nt, status, bytes; };
    Blk blk[aurdf_ctx::kMaxChunks];
    for (int k = 0; k < n_chunks; ++k) {
        const size_t n = (size_t)(t_lo[k + 1] - t_lo[k]);
        blk[k].base = o;
        blk[k].T = take(n * 16 * 8); blk[k].fit = take(n * 8); blk[k].rmse = take(n * 8);
        blk[k].it = take(n * 4); blk[k].nt = take(n * 4); blk[k].status = take(16);
        blk[k].bytes = o - blk[k].base;
    }
    const size_t out_bytes = o;

    int rc;
    if ((rc = grow_pinned(&c->h_in, &c->h_in_bytes, in_bytes)) != AURDF_OK) return rc;
    if ((rc = grow_pinned(&c->h_out, &c->h_out_bytes, out_bytes)) != AURDF_OK) return rc;
    if ((rc = grow_dev(&c->d_in, &c->d_in_bytes, in_bytes)) != AURDF_OK) return rc;
    if ((rc = grow_dev(&c->d_out, &c->d_out_bytes, out_bytes)) != AURDF_OK) return rc;
    for (int k = 0; k < n_chunks; ++k)
        if (!c->streams[k]) AURDF_CUDA_CHECK(cudaStreamCreateWithFlags(&c->streams[k], cudaStreamNonBlocking));
    if (!c->meta_ready) AURDF_CUDA_CHECK(cudaEventCreateWithFlags(&c->meta_ready, cudaEventDisableTiming));
    if (c->cap_chunks != n_chunks || c->cap_tiles != n_tiles) {   // remembered capacities belong to another shape
        for (int k = 0; k < aurdf_ctx::kMaxChunks; ++k) c->cap_hint[k] = 0;
        c->cap_chunks = n_chunks;
        c->cap_tiles = n_tiles;
    }

    char *hi = (char *)c->h_in, *di = (char *)c->d_in, *d_o = (char *)c->d_out, *ho = (char *)c->h_out;
    c->last_h2d = 0;
    c->last_d2h = 0;
    const bool pin_src = is_pinned(src_xyz), pin_tgt = is_pinned(tgt_xyz), pin_box = is_pinned(box_xyz);
    const bool pin_world = is_pinned(out_world_xyz), pin_corr = is_pinned(out_corr);
    // One frame block of pageable buffers (the drop-in call of the reference's frame loop): the staging
    // buffer mirrors the device layout, so everything is packed first and moved by ONE copy each way.
    const bool one_copy_in = n_chunks == 1 && !pin_src && !pin_tgt && !pin_box;
    const bool one_copy_out = n_chunks == 1 && !pin_world && !pin_corr;

    // every failure below must leave no copy in flight on caller memory or on the staging buffers
    auto fail = [&](int code) {
        for (int k = 0; k < aurdf_ctx::kMaxChunks; ++k)
            if (c->streams[k]) cudaStreamSynchronize(c->streams[k]);
        return code;
    };
    auto h2d = [&](cudaStream_t st, size_t off, const void *base, bool pinned, size_t first, size_t bytes) -> int {
        if (!bytes) return AURDF_OK;
        const char *src = (const char *)base + first;
        const void *from = src;
        if (one_copy_in || !pinned) {
            memcpy(hi + off + first, src, bytes);
            from = hi + off + first;
        }
        if (one_copy_in) return AURDF_OK;   // moved by the single copy below
        AURDF_CUDA_CHECK(cudaMemcpyAsync(di + off + first, from, bytes, cudaMemcpyHostToDevice, st));
        c->last_h2d += (int64_t)bytes;
        return AURDF_OK;
    };

    // the packed meta block: offsets, tile_frame, initial poses -- one copy on stream 0, the other blocks wait for it
    memcpy(hi + o_soff, src_off, (size_t)(n_tiles + 1) * 4);
    memcpy(hi + o_toff, tgt_off, (size_t)(n_frames + 1) * 4);
    memcpy(hi + o_tf, tile_frame, (size_t)n_tiles * 4);
    if (box_xyz) memcpy(hi + o_boff, box_off, (size_t)(n_tiles + 1) * 4);
    memcpy(hi + o_init, init_T, (size_t)n_tiles * 16 * 8);
    if (!one_copy_in) {
        AURDF_CUDA_CHECK(cudaMemcpyAsync(di + o_meta, hi + o_meta, meta_bytes, cudaMemcpyHostToDevice, c->streams[0]));
        c->last_h2d += (int64_t)meta_bytes;
        if (n_chunks > 1) AURDF_CUDA_CHECK(cudaEventRecord(c->meta_ready, c->streams[0]));
    }

    struct Chunk { int t0, t1, f0, f1; int64_t s0, s1, cap, cap_upper; size_t ws_off, ws_bytes; };
    Chunk ch[aurdf_ctx::kMaxChunks];
    size_t ws_total = 0;
    for (int k = 0; k < n_chunks; ++k) {
        Chunk &q = ch[k];
        q.t0 = t_lo[k]; q.t1 = t_lo[k + 1];
        q.s0 = src_off[q.t0]; q.s1 = src_off[q.t1];
        q.f0 = q.t1 > q.t0 ? tile_frame[q.t0] : 0;
        q.f1 = q.t1 > q.t0 ? tile_frame[q.t1 - 1] + 1 : 0;
        if (n_chunks == 1) { q.f0 = 0; q.f1 = n_frames; }
        q.cap_upper = 0;
        for (int b = q.t0; b < q.t1; ++b) q.cap_upper += (int64_t)(tgt_off[tile_frame[b] + 1] - tgt_off[tile_frame[b]]) + 1;
        q.cap = c->cap_hint[k] > 0 ? c->cap_hint[k] : 4 * (q.s1 - q.s0) + 2 * (int64_t)(q.t1 - q.t0) + 1024;
        if (q.cap > q.cap_upper) q.cap = q.cap_upper;
        if (q.cap < 2) q.cap = 2;
        q.ws_bytes = align_up(aurdf_icp_workspace_bytes(q.t1 - q.t0, n_src, q.cap), 256);
        q.ws_off = ws_total;
        ws_total += q.ws_bytes;
    }
    if ((rc = grow_dev(&c->d_ws, &c->d_ws_bytes, ws_total)) != AURDF_OK) return fail(rc);

    // queue one chunk on its stream: its points, the kernels, its outputs
    auto run_chunk = [&](int k, bool copy_inputs) -> int {
        const Chunk &q = ch[k];
        cudaStream_t st = c->streams[k];
        const int nt = q.t1 - q.t0;
        if (nt <= 0) return AURDF_OK;
        int rc2;
        if (copy_inputs) {
            if ((rc2 = h2d(st, o_src, src_xyz, pin_src, (size_t)q.s0 * 3 * psz, (size_t)(q.s1 - q.s0) * 3 * psz)) != AURDF_OK) return rc2;
            if ((rc2 = h2d(st, o_tgt, tgt_xyz, pin_tgt, (size_t)tgt_off[q.f0] * 3 * psz, (size_t)(tgt_off[q.f1] - tgt_off[q.f0]) * 3 * psz)) != AURDF_OK) return rc2;
            if (box_xyz && (rc2 = h2d(st, o_box, box_xyz, pin_box, (size_t)box_off[q.t0] * 3 * bsz, (size_t)(box_off[q.t1] - box_off[q.t0]) * 3 * bsz)) != AURDF_OK) return rc2;
            if (one_copy_in) {
                AURDF_CUDA_CHECK(cudaMemcpyAsync(di, hi, in_bytes, cudaMemcpyHostToDevice, st));
                c->last_h2d += (int64_t)in_bytes;
            } else if (k > 0) {
                AURDF_CUDA_CHECK(cudaStreamWaitEvent(st, c->meta_ready, 0));
            }
        }
        int max_src_c = 0;
        for (int b = q.t0; b < q.t1; ++b) max_src_c = src_off[b + 1] - src_off[b] > max_src_c ? src_off[b + 1] - src_off[b] : max_src_c;
        // per-tile arrays are shifted to the chunk's first tile; point arrays keep their base because the
        // offsets stored in src_off / box_off / tgt_off are global
        const Blk &B = blk[k];
        rc2 = aurdf_icp_sweep(di + o_src, pts_dtype, (const int32_t *)(di + o_soff) + q.t0, di + o_tgt,
                              (const int32_t *)(di + o_toff), (const int32_t *)(di + o_tf) + q.t0,
                              box_xyz ? (const void *)(di + o_box) : nullptr, box_dtype,
                              box_xyz ? (const int32_t *)(di + o_boff) + q.t0 : nullptr,
                              (const double *)(di + o_init) + 16 * (size_t)q.t0, nt, n_src, max_src_c, box_scale,
                              max_corr_dist, max_iter, rel_fitness, rel_rmse, ori_only,
                              (double *)(d_o + B.T), (double *)(d_o + q_world), (int32_t *)(d_o + q_corr),
                              (double *)(d_o + B.fit), (double *)(d_o + B.rmse), (int32_t *)(d_o + B.it),
                              (int32_t *)(d_o + B.nt), (char *)c->d_ws + q.ws_off, q.ws_bytes, q.cap,
                              (int32_t *)(d_o + B.status), st);
        if (rc2 != AURDF_OK) return rc2;
        // optimistic: queue the status and every output behind the kernels; if the capacity guess was
        // too small the chunk is simply run again (inputs already resident)
        if (one_copy_out) {   // per-point outputs (if wanted) and the per-tile block in one copy
            const size_t first = (out_world_xyz || out_corr) ? 0 : B.base;
            AURDF_CUDA_CHECK(cudaMemcpyAsync(ho + first, d_o + first, out_bytes - first, cudaMemcpyDeviceToHost, st));
            c->last_d2h += (int64_t)(out_bytes - first);
            return AURDF_OK;
        }
        AURDF_CUDA_CHECK(cudaMemcpyAsync(ho + B.base, d_o + B.base, B.bytes, cudaMemcpyDeviceToHost, st));
        c->last_d2h += (int64_t)B.bytes;
        if (out_world_xyz && q.s1 > q.s0) {
            const size_t first = (size_t)q.s0 * 24, bytes = (size_t)(q.s1 - q.s0) * 24;
            AURDF_CUDA_CHECK(cudaMemcpyAsync(pin_world ? (void *)((char *)out_world_xyz + first) : (void *)(ho + q_world + first),
                                             d_o + q_world + first, bytes, cudaMemcpyDeviceToHost, st));
            c->last_d2h += (int64_t)bytes;
        }
        if (out_corr && q.s1 > q.s0) {
            const size_t first = (size_t)q.s0 * 4, bytes = (size_t)(q.s1 - q.s0) * 4;
            AURDF_CUDA_CHECK(cudaMemcpyAsync(pin_corr ? (void *)((char *)out_corr + first) : (void *)(ho + q_corr + first),
                                             d_o + q_corr + first, bytes, cudaMemcpyDeviceToHost, st));
            c->last_d2h += (int64_t)bytes;
        }
        return AURDF_OK;
    };

    for (int k = 0; k < n_chunks; ++k)
        if ((rc = run_chunk(k, true)) != AURDF_OK) return fail(rc);
    for (int k = 0; k < n_chunks; ++k) {
        if (cudaStreamSynchronize(c->streams[k]) != cudaSuccess) return fail(aurdf::cuda_fail(cudaGetLastError(), "cudaStreamSynchronize"));
        const int32_t *st = (const int32_t *)(ho + blk[k].status);
        int64_t need = (int64_t)(uint32_t)st[1] | ((int64_t)st[2] << 32);
        if (ch[k].t1 > ch[k].t0 && st[0]) {   // capacity guess too small: once more with the exact size
            ch[k].cap = need;
            const size_t ws = align_up(aurdf_icp_workspace_bytes(ch[k].t1 - ch[k].t0, n_src, need), 256);
            if (ws > ch[k].ws_bytes) {   // needs a workspace of its own: wait for the other chunks, then regrow
                for (int j = 0; j < n_chunks; ++j)
                    if (cudaStreamSynchronize(c->streams[j]) != cudaSuccess) return fail(aurdf::cuda_fail(cudaGetLastError(), "cudaStreamSynchronize"));
                // outputs of the other chunks are already on the host; only this chunk uses the new workspace
                void *extra = nullptr;
                cudaError_t e = cudaMalloc(&extra, ws);
                if (e != cudaSuccess) { aurdf::set_error("cudaMalloc(%zu) failed: %s", ws, cudaGetErrorString(e)); return fail(AURDF_ENOMEM); }
                const size_t keep_off = ch[k].ws_off, keep_bytes = ch[k].ws_bytes;
                void *keep_ws = c->d_ws;
                c->d_ws = extra; ch[k].ws_off = 0; ch[k].ws_bytes = ws;
                rc = run_chunk(k, false);
                cudaError_t e2 = cudaStreamSynchronize(c->streams[k]);
                c->d_ws = keep_ws; ch[k].ws_off = keep_off; ch[k].ws_bytes = keep_bytes;
                cudaFree(extra);
                if (rc != AURDF_OK) return fail(rc);
                if (e2 != cudaSuccess) return fail(aurdf::cuda_fail(e2, "cudaStreamSynchronize"));
            } else {
                if ((rc = run_chunk(k, false)) != AURDF_OK) return fail(rc);
                if (cudaStreamSynchronize(c->streams[k]) != cudaSuccess) return fail(aurdf::cuda_fail(cudaGetLastError(), "cudaStreamSynchronize"));
            }
            st = (const int32_t *)(ho + blk[k].status);
            if (st[0]) {
                aurdf::set_error("aurdf_icp_sweep_host: compacted-target capacity %lld still too small", (long long)need);
                return fail(AURDF_ECAPACITY);
            }
            need = (int64_t)(uint32_t)st[1] | ((int64_t)st[2] << 32);
        }
        // next call of the same shape fits first time; never shrinks: the frames of a sequence alternate between
        // needs, and a too-small guess costs a second run (plus, if the workspace must grow, a cudaMalloc)
        if (need + need / 8 + 64 > c->cap_hint[k]) c->cap_hint[k] = need + need / 8 + 64;
        // unpack this block's per-tile outputs while the later blocks are still running
        const size_t n = (size_t)(ch[k].t1 - ch[k].t0), t0 = (size_t)ch[k].t0;
        memcpy(out_T + 16 * t0, ho + blk[k].T, n * 16 * 8);
        memcpy(out_fitness + t0, ho + blk[k].fit, n * 8);
        memcpy(out_rmse + t0, ho + blk[k].rmse, n * 8);
        memcpy(out_iters + t0, ho + blk[k].it, n * 4);
        memcpy(out_ntgt + t0, ho + blk[k].nt, n * 4);
    }
    if (out_world_xyz && !pin_world) memcpy(out_world_xyz, ho + q_world, (size_t)n_src * 24);
    if (out_corr && !pin_corr) memcpy(out_corr, ho + q_corr, (size_t)n_src * 4);
    return AURDF_OK;
}

extern "C" int aurdf_icp_sweep_host(aurdf_ctx *c, const void *src_xyz, int pts_dtype, const int32_t *src_off,
                                    const void *tgt_xyz, const int32_t *tgt_off, const int32_t *tile_frame,
                                    const void *box_xyz, int box_dtype, const int32_t *box_off, const double *init_T,
                                    int32_t n_tiles, int32_t n_frames, double box_scale, double max_corr_dist,
                                    int32_t max_iter, double rel_fitness, double rel_rmse, int32_t ori_only,
                                    double *out_T, double *out_world_xyz, int32_t *out_corr, double *out_fitness,
                                    double *out_rmse, int32_t *out_iters, int32_t *out_ntgt) {
    aurdf::NvtxRange nvtx_range("aurdf_icp_sweep_host");
    AURDF_REQUIRE(c != nullptr, "aurdf_icp_sweep_host: NULL ctx");
    AURDF_REQUIRE(n_tiles >= 0 && n_frames >= 0, "aurdf_icp_sweep_host: negative size");
    std::lock_guard<std::mutex> guard(c->lock);
    if (n_tiles == 0) { c->last_h2d = c->last_d2h = 0; return AURDF_OK; }
    AURDF_REQUIRE(max_corr_dist > 0.0, "aurdf_icp_sweep_host: max_corr_dist must be > 0 (open3d raises)");
    AURDF_REQUIRE(pts_dtype == AURDF_F32 || pts_dtype == AURDF_F64, "aurdf_icp_sweep_host: bad pts_dtype");
    AURDF_REQUIRE(box_dtype == AURDF_F32 || box_dtype == AURDF_F64, "aurdf_icp_sweep_host: bad box_dtype");
    AURDF_REQUIRE(src_off && tgt_off && tile_frame && init_T, "aurdf_icp_sweep_host: NULL input");
    AURDF_REQUIRE(box_xyz == nullptr || box_off != nullptr, "aurdf_icp_sweep_host: box_xyz without box_off");
    AURDF_REQUIRE(out_T && out_fitness && out_rmse && out_iters && out_ntgt, "aurdf_icp_sweep_host: NULL output");
    AURDF_CUDA_CHECK(cudaSetDevice(c->device));
    return sweep_host_locked(c, src_xyz, pts_dtype, src_off, tgt_xyz, tgt_off, tile_frame, box_xyz, box_dtype, box_off,
                             init_T, n_tiles, n_frames, box_scale, max_corr_dist, max_iter, rel_fitness, rel_rmse,
                             ori_only, out_T, out_world_xyz, out_corr, out_fitness, out_rmse, out_iters, out_ntgt);
}
