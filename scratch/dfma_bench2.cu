// DFMA issue rate vs operand pattern (register-file bandwidth) on B200.
#include <cstdio>
#include <cuda_runtime.h>
// V: 0 = fma(a, m, b) constants; 1 = fma(x[c], y[c], a[c]) all distinct; 2 = fma(x[c], s, a[c]) one shared operand
// 3 = 16 accumulators, 4 multiplicands x 4 shared scalars (outer-product pattern, like T[i][d] += f[q] r[q][d])
template <int V>
__global__ void k(double *out, long long *cyc, int iters, const double *in) {
    constexpr int C = 16;
    double a[C], x[C], y[C];
    for (int c = 0; c < C; ++c) { a[c] = in[c] + threadIdx.x; x[c] = in[16 + c] * 1e-9 + 1.0; y[c] = in[32 + c] * 1e-9 + 1.0; }
    const double m = in[50], b = in[51], s = in[52];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (V == 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) a[c] = fma(a[c], m, b);
        } else if (V == 1) {
#pragma unroll
            for (int c = 0; c < C; ++c) a[c] = fma(x[c], y[c], a[c]);
        } else if (V == 2) {
#pragma unroll
            for (int c = 0; c < C; ++c) a[c] = fma(x[c], s, a[c]);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int d = 0; d < 4; ++d) a[q * 4 + d] = fma(x[q], y[d], a[q * 4 + d]);
        }
    }
    long long t1 = clock64();
    double sum = 0;
    for (int c = 0; c < C; ++c) sum += a[c] + x[c] + y[c];
    if (sum == 123.456) out[0] = sum;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int V>
void run(int warps, int iters, double *out, long long *cyc, double *in) {
    k<V><<<148, warps * 32>>>(out, cyc, iters, in);
    cudaDeviceSynchronize();
    k<V><<<148, warps * 32>>>(out, cyc, iters, in);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("variant %d warps/SM %2d: %.2f cyc per warp-DFMA per SMSP\n", V, warps, (double)h / iters / 16 / ((warps + 3) / 4));
}
int main() {
    double *out, *in; long long *cyc;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 64); cudaMalloc(&in, 64 * 8);
    double hin[64]; for (int i = 0; i < 64; ++i) hin[i] = 1.0 + i * 1e-3; hin[50] = 0.999999999; hin[51] = 1e-9; hin[52] = 1e-9;
    cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice);
    for (int w : {8, 16, 32}) { run<0>(w, 20000, out, cyc, in); run<1>(w, 20000, out, cyc, in); run<2>(w, 20000, out, cyc, in); run<3>(w, 20000, out, cyc, in); }
    return 0;
}
