// FP64 issue-rate microbenchmark (nvcc -arch=sm_100a -O3 -o fp64_microbench fp64_microbench.cu)
// Measures warp-instructions per cycle per SM for dependent-free DFMA / DADD / DMUL streams at several
// occupancies, to find the FP64 pipe ceiling that bounds the step kernel.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double *out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c); }
        if (OP == 1) { a0 = __dadd_rn(a0, c); a1 = __dadd_rn(a1, c); a2 = __dadd_rn(a2, c); a3 = __dadd_rn(a3, c); a4 = __dadd_rn(a4, c); a5 = __dadd_rn(a5, c); a6 = __dadd_rn(a6, c); a7 = __dadd_rn(a7, c); }
        if (OP == 2) { a0 = __dmul_rn(a0, b); a1 = __dmul_rn(a1, b); a2 = __dmul_rn(a2, b); a3 = __dmul_rn(a3, b); a4 = __dmul_rn(a4, b); a5 = __dmul_rn(a5, b); a6 = __dmul_rn(a6, b); a7 = __dmul_rn(a7, b); }
        if (OP == 3) { a0 = a0 > a1 ? a0 : a1 + c; a2 = a2 > a3 ? a2 : a3 + c; a4 = a4 > a5 ? a4 : a5 + c; a6 = a6 > a7 ? a6 : a7 + c; a1 = __dadd_rn(a1, c); a3 = __dadd_rn(a3, c); a5 = __dadd_rn(a5, c); a7 = __dadd_rn(a7, c); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <int OP>
void run(const char *name, int warps_per_sm) {
    int sms = 148, threads = 128, blocks = sms * warps_per_sm / 4, iters = 20000;
    double *out; cudaMalloc(&out, (size_t)blocks * threads * 8);
    k<OP><<<blocks, threads>>>(out, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<OP><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = (double)blocks * (threads / 32) * iters * 8.0;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-6s warps/SM=%2d  %.3f ms  warp-inst/cycle/SM = %.3f  (lanes/clk/SM = %.1f)\n", name, warps_per_sm, ms, winst / cycles / sms, winst * 32 / cycles / sms);
    cudaFree(out);
}
int main() {
    for (int w : {4, 12, 16, 32}) { run<0>("DFMA", w); run<1>("DADD", w); run<2>("DMUL", w); run<3>("DSETP+", w); }
    return 0;
}
