// icp_grid.cu -- per-tile ICP for medium and large tiles with an exact, grid-pruned nearest-neighbour search.
//
// Semantics: open3d RegistrationICP point-to-point inside masked_icp (AutoURDF
// PointCloud/cluster_icp.py:118-191, call at :157-159; the plain call sites link.py:113-117 and
// Sim/evaluation.py:358-362 with their 10 000-point clouds), as icp_tiles_kernel.  open3d answers the
// correspondence query from a k-d tree; a brute-force scan is O(n_s n_t) per iteration and is what
// bounded the large tiles of the C5 sweep.  Here:
//
//   build_grid          (device function; its own launch, grid_build_kernel, for the cluster variant) a uniform grid
//                       over the bounding box of the tile's masked targets -- about one target per cell by
//                       volume, at most 64 cells per axis -- and a counting sort of the targets by cell into
//                       float32 records (x, y, z about the tile origin, compacted index) + the end offset of
//                       every cell.
//   icp_grid_kernel<CS> CS = 1: persistent 512-thread CTAs take the tiles the small-tile kernel leaves from a
//                       queue; CS = 8: a thread-block cluster per tile, every CTA a slice of the source points.
//                       The sorted targets and the cell table are staged in shared memory when they fit (targets
//                       first), the source points of a CTA are kept in Morton order of their cells so that a
//                       warp's queries walk overlapping blocks, one thread per source point.  A query scans a
//                       block of cells: the 3x3x3 neighbourhood of its own cell, widened to the cube that holds
//                       the ball through the previous winner.  The float32 scan keeps best and second best;
//                       targets in cells outside the block are at least as far as the nearest face of the
//                       block.  If min(second best, face distance^2) - best > 2 tau (tau over-covers every
//                       float32 rounding) the float32 winner is provably the float64 argmin and only its exact
//                       distance is evaluated, in the reference's operation order; otherwise the block is
//                       rescanned in float64 (lowest index on ties) and grown until the face distance clears the
//                       winner.  The same quantities certify a nearest-neighbour cache (anchor, winner, rho)
//                       exactly as in icp_small2.cu, so about half the points skip the scan once the pose settles.
//
// Measured alternatives (profiles/r02_notes.md): eight lanes per query, miss lists + one warp per large block,
// work items of eight candidates dealt to all threads, and one in-step scan of a warp's bounding block were all
// slower than the per-lane walk over Morton-ordered queries.
//
// Pose fit, convergence rule, pose composition and outputs are those of icp_tiles_kernel.
#include <math.h>

#include <type_traits>

#include <cooperative_groups.h>

#include "icp_common.cuh"

namespace cg = cooperative_groups;

namespace aurdf {

namespace {
constexpr int kGT = 512;
constexpr int kGW = kGT / 32;
constexpr int kGridSortMax = 2048;   // most source points of a CTA (sorted by cell at tile start)

struct GridView {
    const float4 *gs;   // cell-sorted targets of this tile
    const int *ends;    // end offset of every cell (start = end of the previous cell)
    uint32_t gs_s, ends_s;   // the same as shared-space addresses when the grid has been staged in shared memory
    float g0x, g0y, g0z, h, inv_h;
    int Gx, Gy, Gz;
};

// Grid loads.  SH = true: explicit ld.shared on a staged copy; SH = false: through the generic pointers, which point
// at the staged copies when there are any.  The explicit form measured the same (9.07 vs 9.24 ms on C5 16384x32,
// profiles/r02_notes.md), so the kernel instantiates only the generic one and can stage targets and cell table
// independently.
template <bool SH>
__device__ __forceinline__ int ld_end(const GridView &gv, int c) {
    if constexpr (SH) {
        int v;
        asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(gv.ends_s + 4u * (uint32_t)c));
        return v;
    } else {
        return gv.ends[c];
    }
}
template <bool SH>
__device__ __forceinline__ float4 ld_tgt(const GridView &gv, int k) {
    if constexpr (SH) {
        float4 v;
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(gv.gs_s + 16u * (uint32_t)k));
        return v;
    } else {
        return gv.gs[k];
    }
}

__device__ __forceinline__ int cell_of(float v, float g0, float inv_h, int G) {
    const int c = __float2int_rd((v - g0) * inv_h);
    return min(max(c, 0), G - 1);
}

struct Block {
    int lx, hx, ly, hy, lz, hz;
};

// float32 scan of a block of cells: best (m1, sorted position k1) and second best (m2) squared distance.
// The source points of a CTA are sorted by cell (see the kernel), so the lanes of a warp walk (nearly) the same rows:
// the row loop is uniform and the loads of a warp fall on the same few lines.
template <bool SH>
__device__ __forceinline__ void scan_block_f32(const GridView &gv, const Block &bk, float fx, float fy, float fz, float &m1,
                                               float &m2, int &k1) {
    m1 = INFINITY; m2 = INFINITY; k1 = -1;
    for (int cz = bk.lz; cz <= bk.hz; ++cz) {
        for (int cy = bk.ly; cy <= bk.hy; ++cy) {
            const int row = (cz * gv.Gy + cy) * gv.Gx;
            const int a = row + bk.lx;
            int k = a > 0 ? ld_end<SH>(gv, a - 1) : 0;
            const int e = ld_end<SH>(gv, row + bk.hx);
            for (; k + 1 < e; k += 2) {
                const float4 q0 = ld_tgt<SH>(gv, k), q1 = ld_tgt<SH>(gv, k + 1);
                const float dx0 = fx - q0.x, dy0 = fy - q0.y, dz0 = fz - q0.z;
                const float dx1 = fx - q1.x, dy1 = fy - q1.y, dz1 = fz - q1.z;
                const float d0 = fmaf(dz0, dz0, fmaf(dy0, dy0, dx0 * dx0));
                const float d1 = fmaf(dz1, dz1, fmaf(dy1, dy1, dx1 * dx1));
                m2 = fminf(m2, fmaxf(m1, d0));
                k1 = d0 < m1 ? k : k1;
                m1 = fminf(m1, d0);
                m2 = fminf(m2, fmaxf(m1, d1));
                k1 = d1 < m1 ? k + 1 : k1;
                m1 = fminf(m1, d1);
            }
            if (k < e) {
                const float4 q0 = ld_tgt<SH>(gv, k);
                const float dx0 = fx - q0.x, dy0 = fy - q0.y, dz0 = fz - q0.z;
                const float d0 = fmaf(dz0, dz0, fmaf(dy0, dy0, dx0 * dx0));
                m2 = fminf(m2, fmaxf(m1, d0));
                k1 = d0 < m1 ? k : k1;
                m1 = fminf(m1, d0);
            }
        }
    }
}

// Lower bound on the distance from (fx, fy, fz) to any target in a cell outside the block: such a target lies
// beyond one of the block's faces that is not a face of the grid.  eps covers the rounding of the cell function
// and of the float32 target coordinates.
__device__ __forceinline__ float face_distance(const GridView &gv, const Block &bk, float fx, float fy, float fz, float eps) {
    float bnd = INFINITY;
    if (bk.lx > 0) bnd = fminf(bnd, fx - (gv.g0x + (float)bk.lx * gv.h));
    if (bk.hx < gv.Gx - 1) bnd = fminf(bnd, (gv.g0x + (float)(bk.hx + 1) * gv.h) - fx);
    if (bk.ly > 0) bnd = fminf(bnd, fy - (gv.g0y + (float)bk.ly * gv.h));
    if (bk.hy < gv.Gy - 1) bnd = fminf(bnd, (gv.g0y + (float)(bk.hy + 1) * gv.h) - fy);
    if (bk.lz > 0) bnd = fminf(bnd, fz - (gv.g0z + (float)bk.lz * gv.h));
    if (bk.hz < gv.Gz - 1) bnd = fminf(bnd, (gv.g0z + (float)(bk.hz + 1) * gv.h) - fz);
    return bnd == INFINITY ? bnd : fmaxf(bnd - eps, 0.f);
}

__device__ __forceinline__ bool block_is_grid(const GridView &gv, const Block &bk) {
    return bk.lx == 0 && bk.ly == 0 && bk.lz == 0 && bk.hx == gv.Gx - 1 && bk.hy == gv.Gy - 1 && bk.hz == gv.Gz - 1;
}

// cells that hold the cube of half-side R about the point, united with the block
__device__ __forceinline__ bool widen_to_radius(const GridView &gv, Block &bk, float fx, float fy, float fz, float R) {
    const int lx = cell_of(fx - R, gv.g0x, gv.inv_h, gv.Gx), hx = cell_of(fx + R, gv.g0x, gv.inv_h, gv.Gx);
    const int ly = cell_of(fy - R, gv.g0y, gv.inv_h, gv.Gy), hy = cell_of(fy + R, gv.g0y, gv.inv_h, gv.Gy);
    const int lz = cell_of(fz - R, gv.g0z, gv.inv_h, gv.Gz), hz = cell_of(fz + R, gv.g0z, gv.inv_h, gv.Gz);
    const bool grew = lx < bk.lx || hx > bk.hx || ly < bk.ly || hy > bk.hy || lz < bk.lz || hz > bk.hz;
    bk.lx = min(bk.lx, lx); bk.hx = max(bk.hx, hx);
    bk.ly = min(bk.ly, ly); bk.hy = max(bk.hy, hy);
    bk.lz = min(bk.lz, lz); bk.hz = max(bk.hz, hz);
    return grew;
}

__device__ __forceinline__ void grow_block(const GridView &gv, Block &bk, int r) {
    bk.lx = max(bk.lx - r, 0); bk.hx = min(bk.hx + r, gv.Gx - 1);
    bk.ly = max(bk.ly - r, 0); bk.hy = min(bk.hy + r, gv.Gy - 1);
    bk.lz = max(bk.lz - r, 0); bk.hz = min(bk.hz + r, gv.Gz - 1);
}

// Exact float64 scan of a block (the reference's operation order, lowest compacted index on ties), grown until
// every target outside it is provably farther than the winner.  Rare path: out of line.
static __device__ __noinline__ void scan_block_exact(const GridView &gv, Block bk, const double *qx, const double *qy,
                                                     const double *qz, double x, double y, double z, float fx, float fy,
                                                     float fz, float eps, double &bd, int &bj) {
    for (;;) {
        bd = INFINITY;
        bj = -1;
        for (int cz = bk.lz; cz <= bk.hz; ++cz) {
            for (int cy = bk.ly; cy <= bk.hy; ++cy) {
                const int row = (cz * gv.Gy + cy) * gv.Gx;
                const int a = row + bk.lx;
                int k = a > 0 ? gv.ends[a - 1] : 0;
                const int e = gv.ends[row + bk.hx];
                // four candidates per trip: their twelve float64 loads (L2 latency each) are in flight together
                for (; k < e; k += 4) {
                    int j[4];
                    double qx_[4], qy_[4], qz_[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) j[u] = k + u < e ? __float_as_int(gv.gs[k + u].w) : -1;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int jj = max(j[u], 0);
                        qx_[u] = __ldg(qx + jj); qy_[u] = __ldg(qy + jj); qz_[u] = __ldg(qz + jj);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const double dx = __dsub_rn(x, qx_[u]), dy = __dsub_rn(y, qy_[u]), dz = __dsub_rn(z, qz_[u]);
                        const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        if (j[u] >= 0 && (d < bd || (d == bd && j[u] < bj))) { bd = d; bj = j[u]; }
                    }
                }
            }
        }
        if (block_is_grid(gv, bk)) return;
        const float fb = face_distance(gv, bk, fx, fy, fz, eps);
        // fb is a float32 lower bound (already deflated by eps) on the float32-frame distance; deflate once more
        // for the float32 rounding of the point itself before comparing with the float64 winner
        const double fbd = (double)fb * 0.999999 - (double)eps;
        if (bj >= 0 && fbd > 0.0 && fbd * fbd > bd) return;
        grow_block(gv, bk, bj >= 0 ? 1 : 2);
    }
}
}  // namespace

// ------------------------------------------------------------------------------------------
// grid build: one CTA per grid-class tile
// ------------------------------------------------------------------------------------------
// (all kGT threads of a CTA; nt > 0)
static __device__ __noinline__ void build_grid(const IcpParams &p, int b, int nt) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long q0 = p.toff[b];
    const double *qx = p.qx + q0, *qy = p.qy + q0, *qz = p.qz + q0;
    float4 *gs = p.gs + q0;
    int *ends = p.gends + q0 + 2 * (long long)b;
    float *gpar = p.gpar + 8 * (size_t)b;
    const double ox = __ldg(qx), oy = __ldg(qy), oz = __ldg(qz);

    __shared__ float s_red[kGW][7];
    __shared__ float s_par[8];
    __shared__ int s_G[4];
    __shared__ int s_scan[kGT];

    // 1. bounding box and largest coordinate magnitude of the float32 coordinates
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, amax = 0.f;
    for (int j = tid; j < nt; j += kGT) {
        const float f[3] = {(float)(__ldg(qx + j) - ox), (float)(__ldg(qy + j) - oy), (float)(__ldg(qz + j) - oz)};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            lo[d] = fminf(lo[d], f[d]);
            hi[d] = fmaxf(hi[d], f[d]);
            amax = fmaxf(amax, fabsf(f[d]));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    }
    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            s_red[warp][d] = lo[d];
            s_red[warp][3 + d] = hi[d];
        }
        s_red[warp][6] = amax;
    }
    __syncthreads();
    // 2. cell size: about one target per cell by volume, at most 64 cells per axis, at most n_t cells
    if (tid == 0) {
        float l[3], hgh[3], am = 0.f;
        for (int d = 0; d < 3; ++d) {
            l[d] = s_red[0][d];
            hgh[d] = s_red[0][3 + d];
            for (int w = 1; w < kGW; ++w) {
                l[d] = fminf(l[d], s_red[w][d]);
                hgh[d] = fmaxf(hgh[d], s_red[w][3 + d]);
            }
        }
        for (int w = 0; w < kGW; ++w) am = fmaxf(am, s_red[w][6]);
        const float e[3] = {hgh[0] - l[0], hgh[1] - l[1], hgh[2] - l[2]};
        const float emax = fmaxf(e[0], fmaxf(e[1], e[2]));
        int G[3] = {1, 1, 1};
        float h = 1.f;
        if (emax > 0.f) {
            const float fl = 1e-3f * emax;
            h = p.grid_cell_scale * cbrtf(fmaxf(e[0], fl) * fmaxf(e[1], fl) * fmaxf(e[2], fl) / (float)nt);
            h = fmaxf(h, emax / 64.f);
            for (;;) {
                for (int d = 0; d < 3; ++d) G[d] = min(max((int)(e[d] / h) + 1, 1), 64);
                if (G[0] * G[1] * G[2] <= nt) break;
                h *= 1.1f;
            }
        }
        s_par[0] = l[0]; s_par[1] = l[1]; s_par[2] = l[2]; s_par[3] = h; s_par[4] = 1.f / h; s_par[5] = am;
        s_G[0] = G[0]; s_G[1] = G[1]; s_G[2] = G[2];
        gpar[0] = l[0]; gpar[1] = l[1]; gpar[2] = l[2]; gpar[3] = h; gpar[4] = 1.f / h; gpar[5] = am;
        gpar[6] = __int_as_float(G[0] | (G[1] << 8) | (G[2] << 16));
        gpar[7] = 0.f;
    }
    __syncthreads();
    const float g0x = s_par[0], g0y = s_par[1], g0z = s_par[2], inv_h = s_par[4];
    const int Gx = s_G[0], Gy = s_G[1], Gz = s_G[2], ncells = Gx * Gy * Gz;
    // 3. counting sort by cell
    for (int c = tid; c < ncells; c += kGT) ends[c] = 0;
    __syncthreads();
    for (int j = tid; j < nt; j += kGT) {
        const float fx = (float)(__ldg(qx + j) - ox), fy = (float)(__ldg(qy + j) - oy), fz = (float)(__ldg(qz + j) - oz);
        const int c = (cell_of(fz, g0z, inv_h, Gz) * Gy + cell_of(fy, g0y, inv_h, Gy)) * Gx + cell_of(fx, g0x, inv_h, Gx);
        atomicAdd(ends + c, 1);
    }
    __syncthreads();
    // exclusive scan: every thread owns a contiguous run of cells
    const int per = (ncells + kGT - 1) / kGT;
    const int c0 = min(tid * per, ncells), c1 = min(c0 + per, ncells);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += ends[c];
    s_scan[tid] = sum;
    __syncthreads();
    if (warp == 0) {
        int carry = 0;
        for (int base = 0; base < kGT; base += 32) {
            const int v = s_scan[base + lane];
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += n;
            }
            s_scan[base + lane] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();
    int run = s_scan[tid];
    for (int c = c0; c < c1; ++c) {
        const int v = ends[c];
        ends[c] = run;   // start of the cell = cursor of the scatter below
        run += v;
    }
    __syncthreads();
    for (int j = tid; j < nt; j += kGT) {
        const float fx = (float)(__ldg(qx + j) - ox), fy = (float)(__ldg(qy + j) - oy), fz = (float)(__ldg(qz + j) - oz);
        const int c = (cell_of(fz, g0z, inv_h, Gz) * Gy + cell_of(fy, g0y, inv_h, Gy)) * Gx + cell_of(fx, g0x, inv_h, Gx);
        const int pos = atomicAdd(ends + c, 1);   // afterwards ends[c] = end of cell c
        gs[pos] = make_float4(fx, fy, fz, __int_as_float(j));
    }
    __syncthreads();   // the grid is complete (and visible to this CTA)
}

// the cluster variant builds its grids in a launch of its own (one CTA per tile)
__global__ void __launch_bounds__(kGT)
grid_build_kernel(const IcpParams p) {
    if (p.status_int[0]) return;
    const int b = blockIdx.x;
    const int ns_tile = p.src_off[b + 1] - p.src_off[b];
    const int nt = p.cnt[b];
    if (!tile_uses_grid(p, ns_tile, nt) || nt <= 0) return;
    build_grid(p, b, nt);
}

// ------------------------------------------------------------------------------------------
// per-tile ICP, grid search
// ------------------------------------------------------------------------------------------
template <int CS>
__device__ __forceinline__ void grid_tile(const IcpParams &p, const int b) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *sanc = reinterpret_cast<float4 *>(smem_raw);            // cache: anchor (float32 frame), rho
    double *spx_s = reinterpret_cast<double *>(sanc + p.grid_pcap);   // current source points of this CTA's slice
    double *spy_s = spx_s + p.grid_pcap;
    double *spz_s = spy_s + p.grid_pcap;
    int *scj_s = reinterpret_cast<int *>(spz_s + p.grid_pcap);      // nearest target (compacted index) or -1
    int *sord = scj_s + p.grid_pcap;                                // slot -> index of the point within the slice (cell order)
    unsigned *skey = reinterpret_cast<unsigned *>(sord + p.grid_pcap);   // sort keys, kGridSortMax of them (setup only)

    __shared__ double s_part[kGW][16];
    __shared__ int s_cnt[kGW];
    __shared__ double s_tot[2][16];
    __shared__ double s_U[16];
    __shared__ double s_T[16];
    __shared__ double s_prev[2];
    __shared__ double s_warm[18];
    __shared__ int s_stop;
    __shared__ double s_cl[CS][16];
    __shared__ int s_clc[CS];

    if (p.status_int[0]) return;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int rank = 0;
    if constexpr (CS > 1) rank = (int)cg::this_cluster().block_rank();
    const int ns_tile = p.src_off[b + 1] - p.src_off[b];
    const int nt = p.cnt[b];
    if (!tile_uses_grid(p, ns_tile, nt)) return;   // another kernel owns this tile (the whole cluster leaves)
    if constexpr (CS > 1) cg::this_cluster().sync();   // every CTA of the cluster runs before any touches a peer's shared memory
    // CS == 1 owns every tile the small-tile kernel leaves, rank-deficient ones included (strict pose fit)
    const bool strict_tile = nt <= p.strict_nt;
    const int per = (ns_tile + CS - 1) / CS;
    const int lo = min(ns_tile, rank * per);
    const int s0 = p.src_off[b] + lo;
    const int ns = min(ns_tile, lo + per) - lo;
    const long long q0 = p.toff[b];
    const double *qx = p.qx + q0, *qy = p.qy + q0, *qz = p.qz + q0;

    // source-point state: shared memory when the slice fits, else the workspace spill area (no cache then)
    const bool in_smem = ns <= p.grid_pcap;
    double *px, *py, *pz;
    int *cj;
    if (in_smem) {
        px = spx_s; py = spy_s; pz = spz_s; cj = scj_s;
    } else {
        px = p.pspill + 3 * (size_t)s0; py = px + ns; pz = py + ns; cj = p.out_corr + s0;
    }

    if constexpr (CS == 1) {
        if (nt > 0) build_grid(p, b, nt);   // the cluster variant has its grids from grid_build_kernel
    }
    GridView gv;
    gv.gs = p.gs + q0;
    gv.ends = p.gends + q0 + 2 * (long long)b;
    gv.g0x = gv.g0y = gv.g0z = 0.f; gv.h = gv.inv_h = 1.f; gv.Gx = gv.Gy = gv.Gz = 1;
    gv.gs_s = gv.ends_s = 0;
    float aq = 0.f;
    if (nt > 0) {
        const float *gp = p.gpar + 8 * (size_t)b;
        gv.g0x = gp[0]; gv.g0y = gp[1]; gv.g0z = gp[2]; gv.h = gp[3]; gv.inv_h = gp[4];
        aq = gp[5];
        const int G = __float_as_int(gp[6]);
        gv.Gx = G & 0xff; gv.Gy = (G >> 8) & 0xff; gv.Gz = (G >> 16) & 0xff;
        // the sorted targets and the cell table move to shared memory when they fit: the walk is a chain of
        // dependent, scattered loads, which shared memory serves at a fraction of the L1 latency and without the
        // one-line-per-lane gather cost
        const int ncells = gv.Gx * gv.Gy * gv.Gz;
        // sorted targets first (one load per candidate), then the cell table (two loads per row) if it still fits:
        // 8192 x ~9700 tiles keep their 155 KB of targets in shared memory and read the 39 KB table through L1
        size_t avail = (size_t)p.grid_smem_bytes;
        float4 *sg = reinterpret_cast<float4 *>(skey + kGridSortMax);   // grid_pcap is a multiple of 4: 16-byte aligned
        if ((size_t)nt * sizeof(float4) <= avail) {
            for (int k = tid; k < nt; k += kGT) sg[k] = gv.gs[k];
            gv.gs = sg;
            avail -= (size_t)nt * sizeof(float4);
            int *se = reinterpret_cast<int *>(sg + nt);
            if ((size_t)ncells * sizeof(int) <= avail) {
                for (int c = tid; c < ncells; c += kGT) se[c] = gv.ends[c];
                gv.ends = se;
            }
        }
    }

    if (tid == 0) s_stop = 0;
    if (tid < 16) {
        s_T[tid] = p.init_T[16 * (size_t)b + tid];
        s_U[tid] = (tid % 5 == 0) ? 1.0 : 0.0;
    }
    __syncthreads();
    const double ox = nt > 0 ? __ldg(qx) : 0.0, oy = nt > 0 ? __ldg(qy) : 0.0, oz = nt > 0 ? __ldg(qz) : 0.0;
    {
        const bool aff0 = s_T[12] == 0.0 && s_T[13] == 0.0 && s_T[14] == 0.0 && s_T[15] == 1.0;
        auto posed = [&](int i, double &x, double &y, double &z) __attribute__((always_inline)) {   // P = T0 * S
            const size_t e = 3 * (size_t)(s0 + i);
            x = ld_coord(p.src, p.pts_dtype, e); y = ld_coord(p.src, p.pts_dtype, e + 1); z = ld_coord(p.src, p.pts_dtype, e + 2);
            transform_point(s_T, aff0, x, y, z);
        };
        // The points of the slice are kept in Morton order of the grid cell they start in (bitonic sort of
        // (code, index) keys, once per tile): the 32 queries of a warp then sit in a compact box of cells, walk
        // overlapping blocks and load the same lines.  (Scanning the bounding block of a warp's queries once, in
        // step, was measured slower: at ~1 point per cell it holds 4x the candidates of one query's block.)
        // Strict tiles keep the source order (their sums run in ascending source index).
        const bool sorted = in_smem && nt > 0 && !strict_tile && ns > 32;
        if (sorted) {
            int n2 = 64;
            while (n2 < ns) n2 <<= 1;
            for (int i = tid; i < n2; i += kGT) {
                unsigned key = 0xffffffffu;
                if (i < ns) {
                    double x, y, z;
                    posed(i, x, y, z);
                    // Morton code of the cell (<= 6 bits per axis): consecutive points fill compact boxes of cells
                    auto spread = [](unsigned v) { v &= 0x3fu; v = (v | (v << 8)) & 0x300fu; v = (v | (v << 4)) & 0x30c3u; return (v | (v << 2)) & 0x9249u; };
                    const unsigned c = spread((unsigned)cell_of((float)(x - ox), gv.g0x, gv.inv_h, gv.Gx)) |
                                       (spread((unsigned)cell_of((float)(y - oy), gv.g0y, gv.inv_h, gv.Gy)) << 1) |
                                       (spread((unsigned)cell_of((float)(z - oz), gv.g0z, gv.inv_h, gv.Gz)) << 2);
                    key = (c << 11) | (unsigned)i;   // code < 2^18, index < 2^11
                }
                skey[i] = key;
            }
            __syncthreads();
            for (int k = 2; k <= n2; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = tid; t < n2; t += kGT) {
                        const int o = t ^ j;
                        if (o > t) {
                            const unsigned a = skey[t], c = skey[o];
                            if ((a > c) == ((t & k) == 0)) { skey[t] = c; skey[o] = a; }
                        }
                    }
                    __syncthreads();
                }
            }
        }
        for (int t = tid; t < ns; t += kGT) {
            const int i = sorted ? (int)(skey[t] & 0x7ffu) : t;
            double x, y, z;
            posed(i, x, y, z);
            px[t] = x; py[t] = y; pz[t] = z;
            cj[t] = -1;
            if (in_smem) {
                sanc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                sord[t] = i;
            }
        }
    }
    __syncthreads();
    const int rounds = (ns + kGT - 1) / kGT;

    auto exact_d2 = [&](double x, double y, double z, int j) __attribute__((always_inline)) {
        const double dx = __dsub_rn(x, __ldg(qx + j)), dy = __dsub_rn(y, __ldg(qy + j)), dz = __dsub_rn(z, __ldg(qz + j));
        return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    };

    // debug hook (scripts/grid_stats.py): [0] queries, [1] cache hits, [2] block scans, [3] candidates evaluated,
    // [4] exact fallbacks, [5] scans of the whole grid, [7] passes, [12] wait at barrier A, [13] fit + barrier B,
    // [14] longest ICP loop of a CTA (cycles).  dbg[15] selects what is recorded: 1 = counters (their atomics
    // distort every timing), 2 = timers
    unsigned long long *dbg = reinterpret_cast<unsigned long long *>(p.dbg_clock);
    unsigned long long *dbgc = dbg && dbg[15] == 1 ? dbg : nullptr;
    if (dbg && dbg[15] != 2) dbg = nullptr;   // from here on dbg = timers only
    auto count = [&](int k, unsigned long long v) __attribute__((always_inline)) {
        if (dbgc) atomicAdd(dbgc + k, v);
    };
    auto block_candidates = [&](const Block &bk) __attribute__((always_inline)) {
        unsigned long long c = 0;
        for (int cz = bk.lz; cz <= bk.hz; ++cz)
            for (int cy = bk.ly; cy <= bk.hy; ++cy) {
                const int row = (cz * gv.Gy + cy) * gv.Gx, a = row + bk.lx;
                c += gv.ends[row + bk.hx] - (a > 0 ? gv.ends[a - 1] : 0);
            }
        return c;
    };
    // One correspondence pass, one thread per source point: move, cache test, block walk, certificate, moments.
    auto pass_walk = [&](bool apply, auto sh_tag) __attribute__((always_inline)) {
        constexpr bool SH = decltype(sh_tag)::value;
        double acc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = 0.0;
        int cnt = 0;
        const float u = 5.9604645e-8f;
        for (int r = 0; r < rounds; ++r) {
            const int i = r * kGT + tid;
            const bool active = i < ns;
            double x = 0, y = 0, z = 0, bd = INFINITY;
            int bj = -1;
            if (active) {
                x = px[i]; y = py[i]; z = pz[i];
                if (apply) {
                    transform_point(s_U, true, x, y, z);
                    px[i] = x; py[i] = y; pz[i] = z;
                }
            }
            // ---- cache test (per lane)
            const float fx = (float)(x - ox), fy = (float)(y - oy), fz = (float)(z - oz);
            const float amag = fmaxf(aq, fmaxf(fabsf(fx), fmaxf(fabsf(fy), fabsf(fz))));
            const float dl = 4.f * u * amag;
            const float eps = 2e-5f * gv.h + 8.f * u * amag;
            bool scan = false;
            float R = -1.f;
            if (active && nt > 0) {
                scan = true;
                const int j1c = cj[i];
                if (j1c >= 0) {
                    // the previous winner is still the float64 argmin if d(p, winner) + |p - anchor| < rho (icp_small2.cu)
                    const float4 an = in_smem ? sanc[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const double D1 = exact_d2(x, y, z, j1c);
                    const float ex = fx - an.x, ey = fy - an.y, ez = fz - an.z;
                    const float del = sqrtf(fmaf(ez, ez, fmaf(ey, ey, ex * ex)));
                    const float s1 = sqrtf((float)D1);
                    const float lhs = (s1 + del) * 1.0001f + 1e-6f * (amag + del);
                    if (an.w > 0.f && lhs < an.w) {
                        scan = false;
                        bd = D1;
                        bj = j1c;
                        count(1, 1);
                    }
                    R = s1 * 1.01f + 64.f * dl + eps;   // the ball through the previous winner holds the new one
                }
                count(0, 1);
            }
            // ---- the block of cells a scanning lane needs
            Block bk;
            bk.lx = bk.ly = bk.lz = 1 << 20; bk.hx = bk.hy = bk.hz = -1;
            if (scan) {
                const int cx = cell_of(fx, gv.g0x, gv.inv_h, gv.Gx), cy = cell_of(fy, gv.g0y, gv.inv_h, gv.Gy),
                          cz = cell_of(fz, gv.g0z, gv.inv_h, gv.Gz);
                bk.lx = max(cx - 1, 0); bk.hx = min(cx + 1, gv.Gx - 1);
                bk.ly = max(cy - 1, 0); bk.hy = min(cy + 1, gv.Gy - 1);
                bk.lz = max(cz - 1, 0); bk.hz = min(cz + 1, gv.Gz - 1);
                if (R >= 0.f) widen_to_radius(gv, bk, fx, fy, fz, R);
            }
            float m1 = INFINITY, m2 = INFINITY;
            int k1 = -1;
            if (scan) {
                {   // walk the block, widen it until it holds the ball through the best target found
                    int ring = 1;
                    for (;;) {
                        scan_block_f32<SH>(gv, bk, fx, fy, fz, m1, m2, k1);
                        if (dbgc) {
                            count(2, 1);
                            count(3, block_candidates(bk));
                            if (block_is_grid(gv, bk)) count(5, 1);
                        }
                        if (block_is_grid(gv, bk)) break;
                        if (k1 >= 0) {
                            // the block must hold the ball through the best target found so far
                            if (!widen_to_radius(gv, bk, fx, fy, fz, sqrtf(m1) * 1.01f + 64.f * dl + eps)) break;
                        } else {
                            ring *= 2;
                            grow_block(gv, bk, ring);
                        }
                    }
                }
                bool exact = true;
                if (k1 >= 0) {
                    const float fb = face_distance(gv, bk, fx, fy, fz, eps);
                    const float m2e = fminf(m2, fb * fb);
                    const float tau = 16.f * (dl * sqrtf(m2e) * 1.001f + dl * dl + u * m2e);
                    if ((nt == 1 && m2e == INFINITY) || (m2e - m1 > 2.f * tau && m2e < INFINITY)) {
                        exact = false;
                        bj = __float_as_int(ld_tgt<SH>(gv, k1).w);
                        bd = exact_d2(x, y, z, bj);
                        float rho = 0.f;
                        if (m2e == INFINITY) rho = 1e30f;
                        else if (m2e > 128.f * dl * dl && m2e - tau > 0.f) rho = sqrtf(m2e - tau) * 0.9999f;
                        if (in_smem) sanc[i] = make_float4(fx, fy, fz, rho);
                    }
                }
                if (exact) {
                    count(4, 1);
                    scan_block_exact(gv, bk, qx, qy, qz, x, y, z, fx, fy, fz, eps, bd, bj);
                    if (in_smem) sanc[i] = make_float4(fx, fy, fz, 0.f);   // no bound: scan again next time
                }
                cj[i] = bj;
            }
            if (active && bj >= 0 && bd < p.r2) {
                ++cnt;
                const double ax = x - ox, ay = y - oy, az = z - oz;
                const double bx = __ldg(qx + bj) - ox, by = __ldg(qy + bj) - oy, bz = __ldg(qz + bj) - oz;
                acc[0] += bd;
                acc[1] += ax; acc[2] += ay; acc[3] += az;
                acc[4] += bx; acc[5] += by; acc[6] += bz;
                acc[7] += bx * ax; acc[8] += bx * ay; acc[9] += bx * az;
                acc[10] += by * ax; acc[11] += by * ay; acc[12] += by * az;
                acc[13] += bz * ax; acc[14] += bz * ay; acc[15] += bz * az;
            }
        }
        __syncwarp();
        warp_sum16(acc, lane);
        cnt = warp_sum_int(cnt);
        if ((lane & 1) == 0) s_part[warp][lane >> 1] = acc[0];
        if (lane == 0) s_cnt[warp] = cnt;
    };

    auto pass = [&](bool apply) __attribute__((always_inline)) {
        pass_walk(apply, std::false_type{});   // generic loads: the staged copies are reached through the same pointers
    };

    auto totals = [&](int w) __attribute__((always_inline)) -> int {
        if (lane < 16) {
            double t = s_part[0][lane];
#pragma unroll
            for (int k = 1; k < kGW; ++k) t += s_part[k][lane];
            s_tot[w][lane] = t;
        }
        int c = 0;
#pragma unroll
        for (int k = 0; k < kGW; ++k) c += s_cnt[k];
        __syncwarp();
        return c;
    };
    auto cluster_totals = [&](int w) __attribute__((always_inline)) -> int {
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
    auto cluster_publish = [&]() __attribute__((always_inline)) {
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
    auto cluster_fetch = [&]() __attribute__((always_inline)) {
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

    bool have_warm = false;
    auto fit_pose = [&](int c) __attribute__((always_inline)) {
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
    // strict pose fit (warp 0, all lanes; CS == 1 only): the CPU reference's two-pass sums in ascending source
    // index, one accumulator per lane, then its Jacobi SVD on lane 0 (icp_common.cuh, namespace strict)
    auto fit_strict = [&](int c) __attribute__((always_inline)) {
        double Um[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        if (c > 0) {   // warp-uniform
            const double *src1 = lane == 0 ? px : (lane == 1 ? py : (lane == 2 ? pz : (lane == 3 ? qx : (lane == 4 ? qy : qz))));
            double acc = 0.0;
            if (lane < 6) {
                for (int i = 0; i < ns; ++i) {
                    const int j = cj[i];
                    if (j >= 0 && exact_d2(px[i], py[i], pz[i], j) < p.r2) acc = strict::add(acc, src1[lane < 3 ? i : j]);
                }
            }
            const double one_over_n = strict::dvd(1.0, (double)c);
            const double mean = strict::mul(acc, one_over_n);   // lanes 0-2: source mean, 3-5: target mean
            const int rr = lane < 9 ? lane / 3 : 0, cc = lane < 9 ? lane - 3 * rr : 0;
            const double ms_c = __shfl_sync(0xffffffffu, mean, cc), md_r = __shfl_sync(0xffffffffu, mean, 3 + rr);
            const double *pa = cc == 0 ? px : (cc == 1 ? py : pz);
            const double *pb = rr == 0 ? qx : (rr == 1 ? qy : qz);
            double sg = 0.0;
            if (lane < 9) {
                for (int i = 0; i < ns; ++i) {
                    const int j = cj[i];
                    if (j >= 0 && exact_d2(px[i], py[i], pz[i], j) < p.r2)
                        sg = strict::add(sg, strict::mul(strict::sub(pb[j], md_r), strict::sub(pa[i], ms_c)));
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
    auto compose_pose = [&]() __attribute__((always_inline)) {
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

    // it = -1 is open3d's initial correspondence pass (no update applied)
    const long long t_tile = dbg ? clock64() : 0;
    int iters = 0;
#pragma unroll 1
    for (int it = -1; it < p.max_iter; ++it) {
        const bool apply = it >= 0;
        if (apply && rank == 0 && warp == kGW - 1) compose_pose();   // uses s_U of this iteration; next write is after barrier A
        pass(apply);
        if (dbg && tid == 0) atomicAdd(dbg + 7, 1ull);
        long long t_a = dbg ? clock64() : 0;
        __syncthreads();   // barrier A
        if (dbg && tid == 0) { const long long t = clock64(); atomicAdd(dbg + 12, (unsigned long long)(t - t_a)); t_a = t; }
        cluster_publish();
        if (rank == 0) {
            if (warp == 0) {
                const int c = CS > 1 ? cluster_totals(0) : totals(0);
                if (it + 1 < p.max_iter) {   // speculative: overlaps the convergence test
                    if (CS == 1 && strict_tile) fit_strict(c);
                    else if (lane == 0) fit_pose(c);
                }
            } else if (warp == 1) {
                const int c = CS > 1 ? cluster_totals(1) : totals(1);
                if (lane == 0) {
                    const double fit = c > 0 ? (double)c / (double)ns_tile : 0.0;
                    const double rmse = c > 0 ? sqrt(s_tot[1][0] / (double)c) : 0.0;
                    s_stop = (apply && fabs(s_prev[0] - fit) < p.rel_fit && fabs(s_prev[1] - rmse) < p.rel_rmse) ? 1 : 0;
                    s_prev[0] = fit;
                    s_prev[1] = rmse;
                }
            }
        }
        cluster_fetch();
        __syncthreads();   // barrier B
        if (dbg && tid == 0) atomicAdd(dbg + 13, (unsigned long long)(clock64() - t_a));
        if (apply) {
            iters = it + 1;
            if (s_stop) break;
        }
    }
    __syncthreads();
    if (dbg && tid == 0) atomicMax(dbg + 14, (unsigned long long)(clock64() - t_tile));   // longest ICP loop of any CTA

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
    if constexpr (CS > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();
        if (rank != 0 && tid < 16) s_T[tid] = cluster.map_shared_rank(&s_T[0], 0)[tid];
        cluster.sync();
        __syncthreads();
    }
    if (rank == 0 && tid < 16) p.out_T[16 * (size_t)b + tid] = s_T[tid];
    const bool aff = s_T[12] == 0.0 && s_T[13] == 0.0 && s_T[14] == 0.0 && s_T[15] == 1.0;
    for (int t = tid; t < ns; t += kGT) {
        const int i = in_smem ? sord[t] : t;   // slot -> point of the slice
        const int j = cj[t];
        // the winner belongs to the final position of the point (px is not moved after the last pass)
        const bool inl = j >= 0 && exact_d2(px[t], py[t], pz[t], j) < p.r2;
        const size_t e = 3 * (size_t)(s0 + i);
        double x = ld_coord(p.src, p.pts_dtype, e), y = ld_coord(p.src, p.pts_dtype, e + 1),
               z = ld_coord(p.src, p.pts_dtype, e + 2);
        transform_point(s_T, aff, x, y, z);
        p.out_world[e] = x; p.out_world[e + 1] = y; p.out_world[e + 2] = z;
        p.out_corr[s0 + i] = inl ? __ldg(p.qi + q0 + j) : -1;
    }
}

// CS > 1: one cluster per tile.  CS == 1: persistent CTAs (one per SM) take tiles from a queue, so a sweep whose
// tiles all belong to the small-tile kernel costs one CTA per SM walking the tile table, not a grid of n_tiles
// CTAs with 100+ KB of shared memory each.
template <int CS>
__global__ void __launch_bounds__(kGT, 1)
icp_grid_kernel(const IcpParams p) {
    if constexpr (CS > 1) {
        grid_tile<CS>(p, (int)(blockIdx.x / CS));
    } else {
        __shared__ int s_next;
        for (;;) {
            __syncthreads();   // the previous tile is completely done (shared memory is reused)
            if (threadIdx.x == 0) {
                const int k = atomicAdd(p.queue2, 1);   // position in the list tile_scan_kernel made of this kernel's tiles
                s_next = k < p.status_int[5] ? p.glist[k] : p.n_tiles;
            }
            __syncthreads();
            const int b = s_next;
            if (b >= p.n_tiles) return;
            grid_tile<1>(p, b);
        }
    }
}

// per source point: anchor, position, winner, slot -> point index
constexpr size_t kGridBytesPerPoint = sizeof(float4) + 3 * sizeof(double) + 2 * sizeof(int);
size_t icp_grid_smem_bytes(int pcap, int grid_bytes) {
    return (size_t)pcap * kGridBytesPerPoint + (size_t)kGridSortMax * sizeof(unsigned) + (size_t)grid_bytes;
}

int launch_icp_grid(const IcpParams &P, int n_tiles, cudaStream_t stream) {
    if (P.grid_cs > 1) grid_build_kernel<<<n_tiles, kGT, 0, stream>>>(P);   // CS == 1 builds its grid in the ICP kernel
    const size_t smem = icp_grid_smem_bytes(P.grid_pcap, P.grid_smem_bytes);
    int dev = 0;
    const int sms = current_device_sms(&dev);
    // opt-in ceilings: 227 KB per CTA minus the static arrays of each variant
    if (P.grid_cs > 1) {
        if (once_per_device(2, dev))
            AURDF_CUDA_CHECK(cudaFuncSetAttribute(icp_grid_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 223 * 1024));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)n_tiles * 8);
        cfg.blockDim = dim3(kGT);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 8;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        AURDF_CUDA_CHECK(cudaLaunchKernelEx(&cfg, icp_grid_kernel<8>, P));
    } else {
        if (once_per_device(3, dev))
            AURDF_CUDA_CHECK(cudaFuncSetAttribute(icp_grid_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024));
        icp_grid_kernel<1><<<n_tiles < sms ? n_tiles : sms, kGT, smem, stream>>>(P);
    }
    return AURDF_OK;
}

}  // namespace aurdf
