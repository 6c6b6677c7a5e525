// Cell-wise L2 projection into DG_p (fixed-form fast path of the reference's
// `base::local_solver_cholesky` as used by `lsolver/projection.py:17-77`).
// On affine cells the element mass matrix is |detJ| * Mhat and the load vector is
// |detJ| * Phi^T W f_q, so the solve collapses to one (ndg x nq) reference operator
// applied to the nq point values of each cell: pure streaming, HBM bound
// (8*(nq+ndg) bytes per cell and function).
#include <algorithm>

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
// Cell-wise squared L2 norm of a DRT_k function: eta_T^2 = c^T M_T c with the RT mass matrix
// M_T = (g00 M00 + g01 (M01 + M10) + g11 M11) / |detJ| of the cell (the flux-error indicator
// `dot(err_sig, err_sig) * v * dx` of the reference demos, demo_error_estimation.py:100-101,
// for the semi-explicit flux).  Thread per cell, the three reference matrices in shared
// memory (warp-uniform reads); 32 + 8 nrt + 8 bytes per cell, HBM bound.
template <int NRT>
__global__ void flux_norm_kernel(int ncell, const double* __restrict__ mass, const double* __restrict__ cellJ,
                                 const double* __restrict__ sig, double* __restrict__ out)
{
  // symmetric packing: pair t = (i <= q) holds M[i][q] (diagonal) or 2 M[i][q]; padded to an even count
  constexpr int NP = NRT * (NRT + 1) / 2, NPP = (NP + 1) / 2 * 2;
  __shared__ __align__(16) double sM[3 * NPP];
  for (int idx = threadIdx.x; idx < 3 * NPP; idx += blockDim.x)
  {
    const int m = idx / NPP, t = idx - m * NPP;
    double v = 0.0;
    if (t < NP)
    {
      int i = 0, rem = t;
      while (rem >= NRT - i)
      {
        rem -= NRT - i;
        ++i;
      }
      const int q = i + rem;
      v = (i == q ? 1.0 : 2.0) * mass[(m * NRT + i) * NRT + q];
    }
    sM[idx] = v;
  }
  __syncthreads();
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < (size_t)ncell; c += (size_t)gridDim.x * blockDim.x)
  {
    const double2 j0 = reinterpret_cast<const double2*>(cellJ)[2 * c];
    const double2 j1 = reinterpret_cast<const double2*>(cellJ)[2 * c + 1];
    const double iad = 1.0 / fabs(j0.x * j1.y - j0.y * j1.x);
    const double g0 = (j0.x * j0.x + j1.x * j1.x) * iad, g1 = (j0.x * j0.y + j1.x * j1.y) * iad,
                 g2 = (j0.y * j0.y + j1.y * j1.y) * iad;
    double cf[NRT];
#pragma unroll
    for (int i = 0; i < NRT; ++i)
      cf[i] = sig[c * NRT + i];
    double pr[NPP];
    {
      int t = 0;
#pragma unroll
      for (int i = 0; i < NRT; ++i)
#pragma unroll
        for (int q = i; q < NRT; ++q)
          pr[t++] = cf[i] * cf[q];
      if (NPP > NP)
        pr[NP] = 0.0;
    }
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    const double2* m0 = reinterpret_cast<const double2*>(sM);
    const double2* m1 = reinterpret_cast<const double2*>(sM + NPP);
    const double2* m2 = reinterpret_cast<const double2*>(sM + 2 * NPP);
#pragma unroll
    for (int t = 0; t < NPP / 2; ++t)
    {
      const double2 v0 = m0[t], v1 = m1[t], v2 = m2[t];
      a0 += pr[2 * t] * v0.x + pr[2 * t + 1] * v0.y;
      a1 += pr[2 * t] * v1.x + pr[2 * t + 1] * v1.y;
      a2 += pr[2 * t] * v2.x + pr[2 * t + 1] * v2.y;
    }
    out[c] = g0 * a0 + g1 * a1 + g2 * a2;
  }
}
} // namespace

void launch_flux_norm(eqlb_handle* h, int nfun, const double* const* dsig, double* const* dout)
{
  const int bs = 128;
  const int grid = (int)std::min<size_t>(((size_t)h->ncell + bs - 1) / bs, (size_t)148 * 16);
  const double* mass = h->tv.data + h->tv.o_rt_mass;
  for (int f = 0; f < nfun; ++f)
  {
    switch (h->nrt)
    {
    case 3:
      flux_norm_kernel<3><<<grid, bs, 0, h->stream>>>(h->ncell, mass, h->d_cellJ.p, dsig[f], dout[f]);
      break;
    case 8:
      flux_norm_kernel<8><<<grid, bs, 0, h->stream>>>(h->ncell, mass, h->d_cellJ.p, dsig[f], dout[f]);
      break;
    case 15:
      flux_norm_kernel<15><<<grid, bs, 0, h->stream>>>(h->ncell, mass, h->d_cellJ.p, dsig[f], dout[f]);
      break;
    case 24:
      flux_norm_kernel<24><<<grid, bs, 0, h->stream>>>(h->ncell, mass, h->d_cellJ.p, dsig[f], dout[f]);
      break;
    default:
      throw EqlbError(EQLB_ERR_INPUT, "eqlb_flux_l2norm: unsupported flux degree");
    }
    CUDA_CHECK(cudaGetLastError());
    h->launches++;
  }
}

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

// ---------------------------------------------------------------------------
// EV epilogue: conforming hierarchic-RT vector -> DOF vector of the same function w.r.t. the functionals of
// Basix' "RT" element, Legendre variant (`FluxEqlbEV.py:95`).  Facet part: k x k matrix A per facet (both sets
// of functionals are normal moments in the global facet orientation).  Interior part: the Basix interior
// moments of a cell are a fixed linear map B of ALL cell-local hierarchic DOFs; the cell-local facet DOFs come
// from the global ones through the reflection matrix R (c_loc = R c_glob, R an involution).
// Pure streaming: 8 (k + ...) B per facet / cell, HBM bound.
// ---------------------------------------------------------------------------
namespace
{
__global__ void basix_facet_kernel(int nfct, int k, const double* __restrict__ A, const double* __restrict__ in,
                                   double* __restrict__ out)
{
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nfct)
    return;
  double c[4], r[4];
  for (int j = 0; j < k; ++j)
    c[j] = in[(size_t)f * k + j];
  for (int i = 0; i < k; ++i)
  {
    double s = 0.0;
    for (int j = 0; j < k; ++j)
      s += A[i * k + j] * c[j];
    r[i] = s;
  }
  for (int i = 0; i < k; ++i)
    out[(size_t)f * k + i] = r[i];
}

__global__ void basix_interior_kernel(int ncell, int nfct, int k, int nrt, const double* __restrict__ B,
                                      const double* __restrict__ trafo, const int32_t* __restrict__ cell_fct,
                                      const uint8_t* __restrict__ fct_perms, const double* __restrict__ in,
                                      double* __restrict__ out)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell)
    return;
  const int ncd = k * k - k;
  double loc[24];
  for (int f = 0; f < 3; ++f)
  {
    const int32_t gf = cell_fct[3 * (size_t)c + f];
    double g[4];
    for (int j = 0; j < k; ++j)
      g[j] = in[(size_t)gf * k + j];
    if (fct_perms[3 * (size_t)c + f])
    {
      // c_loc = R c_glob with R[i][j] = trafo[j][i] (trafo holds R^T, se/KernelData.cpp:55-64)
      for (int i = 0; i < k; ++i)
      {
        double s = 0.0;
        for (int j = 0; j < k; ++j)
          s += trafo[j * k + i] * g[j];
        loc[f * k + i] = s;
      }
    }
    else
      for (int j = 0; j < k; ++j)
        loc[f * k + j] = g[j];
  }
  for (int i = 0; i < ncd; ++i)
    loc[3 * k + i] = in[(size_t)nfct * k + (size_t)c * ncd + i];
  for (int i = 0; i < ncd; ++i)
  {
    double s = 0.0;
    for (int j = 0; j < nrt; ++j)
      s += B[i * nrt + j] * loc[j];
    out[(size_t)nfct * k + (size_t)c * ncd + i] = s;
  }
}
} // namespace

void launch_ev_to_basix(eqlb_handle* h, int nfun, const double* const* din, double* const* dout)
{
  if (!h->d_basix.p)
    throw EqlbError(EQLB_ERR_STATE, "eqlb_ev_to_basix_rt: the tables carry no hierarchic -> Basix maps");
  if (h->k > 4)
    throw EqlbError(EQLB_ERR_INPUT, "eqlb_ev_to_basix_rt: degree > 4 not supported");
  const int bs = 256, k = h->k;
  const double* A = h->d_basix.p;
  const double* B = h->d_basix.p + k * k;
  const double* trafo = h->tv.data + h->tv.o_trafo;
  for (int f = 0; f < nfun; ++f)
  {
    basix_facet_kernel<<<(h->nfct + bs - 1) / bs, bs, 0, h->stream>>>(h->nfct, k, A, din[f], dout[f]);
    CUDA_CHECK(cudaGetLastError());
    h->launches++;
    if (k > 1)
    {
      basix_interior_kernel<<<(h->ncell + bs - 1) / bs, bs, 0, h->stream>>>(h->ncell, h->nfct, k, h->nrt, B, trafo, h->d_cell_fct.p,
                                                                              h->d_fct_perms.p, din[f], dout[f]);
      CUDA_CHECK(cudaGetLastError());
      h->launches++;
    }
  }
}
