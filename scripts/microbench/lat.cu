// Microbenchmarks (B200): dependent-chain latency and per-SM throughput of the instructions the
// per-tile ICP kernel is built from.  nvcc -arch=sm_100a -O3 lat.cu -o lat && ./lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N 2048

template <int OP>
__global__ void chain(double a, double b, float fa, float fb, uint32_t ua, uint32_t ub, long long *out, double *sink) {
    double x = a; float f = fa; uint32_t u = ua; float2 f2 = make_float2(fa, fb);
    __shared__ double sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = threadIdx.x & 1;   // indices 0/1 chain
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = fma(x, b, a);
        if (OP == 1) x = x + b;
        if (OP == 2) x = x * b;
        if (OP == 3) f = fmaf(f, fb, fa);
        if (OP == 4) u = min(u ^ ub, ub + i);          // VIMNMX + LOP3 (2 dependent ALU ops)
        if (OP == 5) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
        if (OP == 6) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
        if (OP == 7) x = 1.0 / x;
        if (OP == 8) x = sqrt(x);
        if (OP == 9) x = sm[(int)x];                   // LDS + F2I/I2F-free chain? (double->int conv included)
        if (OP == 10) u = __vimin3_u32(u, ub, ua + i) + 1;
        if (OP == 11) f2 = __ffma2_rn(f2, f2, make_float2(fb, fa));
        if (OP == 12) x = __shfl_xor_sync(0xffffffffu, x, 1);
        if (OP == 13) __syncthreads();
        if (OP == 14) f = fminf(f * fb, fa);            // FMUL + FMNMX
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = x + f + u + f2.x + f2.y;
}

// throughput: W warps per block, each with 8 independent chains of one op
template <int OP>
__global__ void tput(float fa, float fb, uint32_t ua, uint32_t ub, double da, double db, long long *out, float *sink) {
    float f[8]; uint32_t u[8]; float2 g[8]; double d[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { f[k] = fa + k; u[k] = ua + k * 77u + threadIdx.x; g[k] = make_float2(fa + k, fb - k); d[k] = da + k; }
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (OP == 0) f[k] = fmaf(f[k], fb, fa);
            if (OP == 1) u[k] = min(u[k], ub + i);                       // VIMNMX
            if (OP == 2) u[k] = __vimin3_u32(u[k], ub + i, ua);          // VIMNMX3
            if (OP == 3) u[k] = (u[k] & 0xfffffc00u) | (ub + i);         // LOP3
            if (OP == 4) f[k] = fminf(f[k], fb + i);                     // FMNMX
            if (OP == 5) g[k] = __ffma2_rn(g[k], g[k], make_float2(fb, fa));
            if (OP == 6) g[k] = __fadd2_rn(g[k], make_float2(fb, fa));
            if (OP == 7) d[k] = fma(d[k], db, da);
            if (OP == 8) u[k] = u[k] + ub;                               // IADD
            if (OP == 9) u[k] = u[k] * ub + ua;                          // IMAD
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    float s = 0; for (int k = 0; k < 8; ++k) s += f[k] + u[k] + g[k].x + g[k].y + (float)d[k];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    long long *out; double *sink; float *fs;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&sink, 8 * 1024 * 1024); cudaMalloc(&fs, 4 * 1024 * 1024);
    long long h[8];
    const char *names[] = {"DFMA chain", "DADD chain", "DMUL chain", "FFMA chain", "VIMNMX+LOP3 chain (2 ops)", "MUFU.RCP64H (+fixup) chain", "MUFU.RSQ64H chain",
                           "double division chain", "double sqrt chain", "LDS + F2I chain", "VIMNMX3+IADD chain (2 ops)", "FFMA2 chain", "SHFL f64 chain (2 SHFL)",
                           "__syncthreads", "FMUL+FMNMX chain (2 ops)"};
#define RUNC(OP, T) chain<OP><<<1, T>>>(1.0000001, 0.9999999, 1.0001f, 0.9999f, 12345u, 777u, out, sink); cudaDeviceSynchronize(); \
    cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost); printf("%-32s threads %4d: %.2f cycles/iter\n", names[OP], T, (double)h[0] / N);
    RUNC(0, 32) RUNC(1, 32) RUNC(2, 32) RUNC(3, 32) RUNC(4, 32) RUNC(5, 32) RUNC(6, 32) RUNC(7, 32) RUNC(8, 32) RUNC(9, 32) RUNC(10, 32) RUNC(11, 32) RUNC(12, 32)
    RUNC(13, 128) RUNC(13, 256) RUNC(14, 32)
    const char *tn[] = {"FFMA", "VIMNMX", "VIMNMX3", "LOP3", "FMNMX", "FFMA2", "FADD2", "DFMA", "IADD", "IMAD"};
#define RUNT(OP, T) tput<OP><<<1, T>>>(1.0001f, 0.9999f, 12345u, 777u, 1.0000001, 0.9999999, out, fs); cudaDeviceSynchronize(); \
    cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost); printf("tput %-8s threads %4d: %.3f cycles per warp-instr per SMSP (warps/SMSP=%d)\n", tn[OP], T, (double)h[0] / (N * 8.0 * (T / 128.0 > 1 ? T / 128.0 : 1)), T / 128 > 1 ? T / 128 : 1);
    RUNT(0, 128) RUNT(0, 512) RUNT(1, 128) RUNT(1, 512) RUNT(2, 128) RUNT(2, 512) RUNT(3, 128) RUNT(3, 512) RUNT(4, 128) RUNT(4, 512)
    RUNT(5, 128) RUNT(5, 512) RUNT(6, 128) RUNT(6, 512) RUNT(7, 128) RUNT(7, 512) RUNT(8, 128) RUNT(8, 512) RUNT(9, 128) RUNT(9, 512)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
