// Cell-wise L2 projection into DG_p (fixed-form fast path of the reference's
// `base::local_solver_cholesky` as used by `lsolver/projection.py:17-77`).
// On affine cells the element mass matrix is |detJ| * Mhat and the load vector is
// |detJ| * Phi^T W f_q, so the solve collapses to one (ndg x nq) reference operator
// applied to the nq point values of each cell: pure streaming, HBM bound
// (8*(nq+ndg) bytes per cell and function).
#include "eqlb_internal.cuh"

namespace
{
__global__ void project_kernel(int ncell, int ndg, int nq, const double* __restrict__ P, const double* __restrict__ qv,
                               double* __restrict__ out)
{
  extern __shared__ double sP[];
  for (int i = threadIdx.x; i < ndg * nq; i += blockDim.x)
    sP[i] = P[i];
  __syncthreads();
  const size_t total = (size_t)ncell * ndg;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
  {
    const size_t c = idx / ndg;
    const int i = (int)(idx - c * ndg);
    const double* v = qv + c * nq;
    const double* p = sP + i * nq;
    double s = 0.0;
    for (int q = 0; q < nq; ++q)
      s += p[q] * v[q];
    out[idx] = s;
  }
}
} // namespace

void launch_project(eqlb_handle* h, int nfun, const double* const* dq, double* const* dout)
{
  const int bs = 256;
  const size_t total = (size_t)h->ncell * h->ndg;
  const int grid = (int)std::min<size_t>((total + bs - 1) / bs, (size_t)148 * 16);
  for (int f = 0; f < nfun; ++f)
  {
    project_kernel<<<grid, bs, (size_t)h->ndg * h->nq * sizeof(double), h->stream>>>(h->ncell, h->ndg, h->nq, h->d_proj.p,
                                                                                  dq[f], dout[f]);
    CUDA_CHECK(cudaGetLastError());
    h->launches++;
  }
}
