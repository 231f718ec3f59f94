// Phase-B inner loop in isolation: SB samples per thread, coefficients from shared memory.
#include <cstdio>
#include <cuda_runtime.h>
template <int SB, int M>
__global__ void __launch_bounds__(256, 1) k(double *out, long long *cyc, int iters, const double *in) {
    __shared__ double sAn[M * 2], sAo[M * 2], sC[M];
    for (int i = threadIdx.x; i < M * 2; i += blockDim.x) { sAn[i] = in[i]; sAo[i] = in[64 + i]; }
    for (int i = threadIdx.x; i < M; i += blockDim.x) sC[i] = in[128 + i];
    __syncthreads();
    double c2[SB], f[SB], fm[SB], en[SB][2], eo[SB][2], v[SB];
    for (int q = 0; q < SB; ++q) { c2[q] = in[200 + q] + threadIdx.x * 1e-9; f[q] = in[210 + q]; fm[q] = 0; v[q] = 0; en[q][0] = en[q][1] = eo[q][0] = eo[q][1] = 0; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        asm volatile("" ::: "memory");
#pragma unroll 5
        for (int i = 0; i < M; ++i) {
            const double a0 = sAn[i * 2], a1 = sAn[i * 2 + 1], b0 = sAo[i * 2], b1 = sAo[i * 2 + 1], c = sC[i];
#pragma unroll
            for (int q = 0; q < SB; ++q) {
                en[q][0] = fma(f[q], a0, en[q][0]);
                en[q][1] = fma(f[q], a1, en[q][1]);
                eo[q][0] = fma(f[q], b0, eo[q][0]);
                eo[q][1] = fma(f[q], b1, eo[q][1]);
                v[q] = fma(f[q] * c, f[q], v[q]);
                const double fn = fma(c2[q], f[q], -fm[q]);
                fm[q] = f[q];
                f[q] = fn;
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int q = 0; q < SB; ++q) s += en[q][0] + en[q][1] + eo[q][0] + eo[q][1] + v[q] + f[q];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int SB>
void run(int warps, double *out, long long *cyc, double *in) {
    const int iters = 2000, M = 30;
    k<SB, M><<<148, warps * 32>>>(out, cyc, iters, in);
    cudaDeviceSynchronize();
    k<SB, M><<<148, warps * 32>>>(out, cyc, iters, in);
    cudaError_t e = cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double fp64_per_iter = 7.0 * M * SB;   // warp-level FP64 instructions per thread-iteration
    printf("SB %d warps/SM %2d (%s): %.2f cycles per FP64 warp-instr per SMSP\n", SB, warps, cudaGetErrorString(e),
           (double)h / iters / fp64_per_iter / (warps / 4.0));
}
int main() {
    double *out, *in; long long *cyc;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 64); cudaMalloc(&in, 256 * 8);
    double hin[256]; for (int i = 0; i < 256; ++i) hin[i] = 1e-3 * (i % 17) + 0.1; for (int q = 0; q < 8; ++q) { hin[200 + q] = 1.9 + 0.01 * q; hin[210 + q] = 0.3; }
    cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice);
    for (int w : {4, 8}) { run<1>(w, out, cyc, in); run<2>(w, out, cyc, in); run<4>(w, out, cyc, in); }
    return 0;
}
