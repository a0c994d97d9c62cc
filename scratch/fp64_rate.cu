#include <cstdio>
#include <cuda_runtime.h>
template <typename T> __global__ void k(T* out, T a, T b, int iters) {
    T x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
        x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
template <typename T> double run(const char* name) {
    T* d; cudaMalloc(&d, sizeof(T) * 148 * 8 * 512);
    int iters = 20000;
    k<T><<<148 * 8, 512>>>(d, (T)1.0000001, (T)1e-9, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<T><<<148 * 8, 512>>>(d, (T)1.0000001, (T)1e-9, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 8 * iters * 148.0 * 8 * 512;
    printf("%s: %.3f ms  %.2f TFLOP/s\n", name, ms, flops / ms * 1e-9);
    cudaFree(d); return ms;
}
int main() { run<float>("fp32 FMA"); run<double>("fp64 FMA"); run<float>("fp32 FMA"); run<double>("fp64 FMA"); return 0; }
