// Microbenchmarks (B200), part 2: FP64 tensor-core mma latency / throughput, generic vs shared loads,
// mixed-pipe issue rate of a single warp.  nvcc -arch=sm_100a -O3 lat2.cu -o lat2 && ./lat2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 1024

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int CH>   // CH independent accumulator chains
__global__ void k_dmma(double a, double b, long long *out, double *sink) {
    double d[8][2];
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k][0] = d[k][1] = k;
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int k = 0; k < CH; ++k) dmma(d[k][0], d[k][1], a, b);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += d[k][0] + d[k][1];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>   // 0: LDS chain (typed shared), 1: generic LD chain on shared, 2: LDL-ish (local array) chain, 3: LDG L2 chain
__global__ void k_ld(int *g, long long *out, int *sink) {
    __shared__ int sm[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sm[i] = (i + 1) & 255;
    __syncthreads();
    int idx = threadIdx.x & 1;
    const int *gp = MODE == 1 ? (const int *)sm : g;   // generic pointer (runtime select defeats address-space inference)
    if (MODE == 1 && g == nullptr) gp = sm;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
        if (MODE == 0) idx = sm[idx];
        if (MODE == 1) idx = gp[idx];
        if (MODE == 3) idx = __ldcg(g + idx);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    sink[threadIdx.x] = idx;
}

// single warp, independent FFMA2 + VIMNMX streams interleaved: can the two pipes overlap for one warp?
template <int MODE>
__global__ void k_mix(float fa, float fb, uint32_t ua, long long *out, float *sink) {
    float2 g[6]; uint32_t u[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { g[k] = make_float2(fa + k, fb - k); u[k] = ua + 7 * k + threadIdx.x; }
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            if (MODE != 1) g[k] = __ffma2_rn(g[k], g[k], make_float2(fb, fa));
            if (MODE != 0) u[k] = min(u[k], ua + i) ^ 1u;      // VIMNMX + LOP3
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    float s = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s += g[k].x + g[k].y + u[k];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    long long *out; double *sink; int *g;
    cudaMalloc(&out, 4096); cudaMalloc(&sink, 1 << 22); cudaMalloc(&g, 1024 * 4);
    int hg[1024]; for (int i = 0; i < 1024; ++i) hg[i] = (i + 1) & 255;
    cudaMemcpy(g, hg, sizeof hg, cudaMemcpyHostToDevice);
    long long h;
#define REP(name, launch, per) launch; cudaDeviceSynchronize(); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost); printf("%-52s %.2f cycles\n", name, (double)h / (per));
    REP("DMMA m8n8k4 dependent chain, 1 warp: per mma", (k_dmma<1><<<1, 32>>>(1.0, 0.5, out, sink)), N)
    REP("DMMA 2 chains, 1 warp: per mma", (k_dmma<2><<<1, 32>>>(1.0, 0.5, out, sink)), 2.0 * N)
    REP("DMMA 4 chains, 1 warp: per mma", (k_dmma<4><<<1, 32>>>(1.0, 0.5, out, sink)), 4.0 * N)
    REP("DMMA 8 chains, 1 warp: per mma", (k_dmma<8><<<1, 32>>>(1.0, 0.5, out, sink)), 8.0 * N)
    REP("DMMA 8 chains, 4 warps (1/SMSP): per mma per warp", (k_dmma<8><<<1, 128>>>(1.0, 0.5, out, sink)), 8.0 * N)
    REP("DMMA 8 chains, 16 warps (4/SMSP): per mma per SMSP", (k_dmma<8><<<1, 512>>>(1.0, 0.5, out, sink)), 8.0 * N * 4)
    REP("LDS dependent chain", (k_ld<0><<<1, 32>>>(g, out, (int *)sink)), N)
    REP("generic LD on shared, dependent chain", (k_ld<1><<<1, 32>>>(nullptr, out, (int *)sink)), N)
    REP("LDG.CG (L2 hit) dependent chain", (k_ld<3><<<1, 32>>>(g, out, (int *)sink)), N)
    REP("1 warp: 6 FFMA2 per iter", (k_mix<0><<<1, 32>>>(1.0001f, 0.9999f, 5u, out, (float *)sink)), N)
    REP("1 warp: 6 (VIMNMX+LOP3) per iter", (k_mix<1><<<1, 32>>>(1.0001f, 0.9999f, 5u, out, (float *)sink)), N)
    REP("1 warp: 6 FFMA2 + 6 (VIMNMX+LOP3) per iter", (k_mix<2><<<1, 32>>>(1.0001f, 0.9999f, 5u, out, (float *)sink)), N)
    REP("2 warps/SMSP: 6 FFMA2 + 6 (VIMNMX+LOP3) per iter", (k_mix<2><<<1, 256>>>(1.0001f, 0.9999f, 5u, out, (float *)sink)), N)
    REP("4 warps/SMSP: 6 FFMA2 + 6 (VIMNMX+LOP3) per iter", (k_mix<2><<<1, 512>>>(1.0001f, 0.9999f, 5u, out, (float *)sink)), N)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
