// DFMA latency / throughput microbenchmark (B200): cycles per DFMA for C independent chains per thread
// and W warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void k(double *out, long long *cyc, int iters) {
    double a[C];
    for (int c = 0; c < C; ++c) a[c] = 1.0 + threadIdx.x * 1e-9 + c * 1e-3;
    const double m = 0.999999999, b = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < C; ++c) a[c] = fma(a[c], m, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < C; ++c) s += a[c];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int C>
void run(int warps_per_sm, int iters, double *out, long long *cyc) {
    int sms = 148;
    k<C><<<sms, warps_per_sm * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    k<C><<<sms, warps_per_sm * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / iters / C;
    // per-SMSP issue interval: warps_per_sm/4 warps share an SMSP
    printf("chains %2d warps/SM %2d: %.2f cyc per DFMA per warp, %.2f cyc per warp-DFMA per SMSP\n", C, warps_per_sm, per,
           per / ((warps_per_sm + 3) / 4));
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
    int it = 20000;
    for (int w : {1, 4, 8, 16, 32}) {
        run<1>(w, it, out, cyc); run<2>(w, it, out, cyc); run<4>(w, it, out, cyc); run<8>(w, it, out, cyc); run<16>(w, it, out, cyc);
    }
    return 0;
}
