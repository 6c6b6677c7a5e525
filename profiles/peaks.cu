// measured FP64 FMA peak and pinned PCIe copy bandwidths (evidence for DESIGN.md)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double* out, int iters)
{
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
  double* d; cudaMalloc(&d, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep)
  {
    const int iters = 20000;
    cudaEventRecord(e0); dfma<<<148 * 8, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("fp64 fma: %.2f TFLOP/s (%.3f ms)\n", 2.0 * 148 * 8 * 256 * 8.0 * iters / (ms * 1e-3) / 1e12, ms);
  }
  const size_t n = 300u << 20;
  void *h, *h2, *dd, *dd2; cudaMallocHost(&h, n); cudaMallocHost(&h2, n); cudaMalloc(&dd, n); cudaMalloc(&dd2, n);
  cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  for (int rep = 0; rep < 3; ++rep)
  {
    float ms;
    cudaEventRecord(e0, s1); cudaMemcpyAsync(dd, h, n, cudaMemcpyHostToDevice, s1); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("h2d %.1f GB/s  ", n / (ms * 1e-3) / 1e9);
    cudaEventRecord(e0, s1); cudaMemcpyAsync(h, dd, n, cudaMemcpyDeviceToHost, s1); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("d2h %.1f GB/s  ", n / (ms * 1e-3) / 1e9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, s1); cudaMemcpyAsync(dd, h, n, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(h2, dd2, n, cudaMemcpyDeviceToHost, s2);
    cudaStreamSynchronize(s2); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("both directions at once: %.1f GB/s each\n", n / (ms * 1e-3) / 1e9);
  }
  return 0;
}
