// Measurement aid (not on the hot path): FP64 FMA peak of the current device, the denominator of
// the FP64 roofline bench.py prints (`roofline.fp64.peak`).  Register-resident DFMA chains, 8
// independent accumulators per thread, 8 CTAs x 256 threads per SM.
#include "eqlb_internal.cuh"

namespace
{
__global__ void dfma_chain_kernel(double* out, int iters)
{
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      a[i] = fma(a[i], b, c);
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
} // namespace

extern "C" int eqlb_measure_fp64_peak(int iters, int reps, double* tflops)
{
  try
  {
    if (!tflops || iters < 1 || reps < 1)
      throw EqlbError(EQLB_ERR_INPUT, "eqlb_measure_fp64_peak: bad argument");
    int dev = 0, nsm = 148;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    const int grid = nsm * 8, bs = 256;
    double* d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, (size_t)grid * bs * sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int r = 0; r < reps + 1; ++r)  // first launch = warm-up
    {
      CUDA_CHECK(cudaEventRecord(e0));
      dfma_chain_kernel<<<grid, bs>>>(d, iters);
      CUDA_CHECK(cudaEventRecord(e1));
      CUDA_CHECK(cudaEventSynchronize(e1));
      float ms = 0.f;
      CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
      const double tf = 2.0 * grid * bs * 8.0 * iters / (ms * 1e-3) / 1e12;
      if (r > 0 && tf > best)
        best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return EQLB_OK;
  }
  catch (const EqlbError& e)
  {
    eqlb_set_error(e.what());
    return e.code;
  }
  catch (const std::exception& e)
  {
    eqlb_set_error(e.what());
    return EQLB_ERR_CUDA;
  }
}
