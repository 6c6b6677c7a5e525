// Fused input stage and error estimator: the callers right before and right after the
// patch loops, so that a host caller moves the primal solution in and cell-wise indicators out instead of the
// projected fluxes (24 ndg B per cell in) and the flux vector (8 nrt B per cell out).
//   primal_project_kernel  `lsolver/projection.py:17-77` for G = Pi(-grad u_h), F = Pi f_h: on affine cells
//                          grad u_h is piecewise P_{k-1}, so its DG_{k-1} projection is its nodal interpolant:
//                          G_i = -K^T sum_j u_j grad_ref phi_j(x_i) (reference table [ndg][npk][2]);
//                          F = P f_cell with the exact reference projection matrix [ndg][npk].
//   estimate_*_kernel      `demo/poisson/demo_error_estimation.py:52-122`, `demo/elasticity/demo_error_estimation.py
//                          :49-135`: thread per cell, quadrature loop over the tabulated bases (degree 2k+1 rule:
//                          exact for every integrand here).
// All of these stream each cell once: HBM bound, a few hundred flops per cell.
#include "eqlb_internal.cuh"

namespace
{
constexpr int MAXPK = 15, MAXRT = 24, MAXDG = 10;

__global__ void primal_project_kernel(int ncell, int npk, int ndg, const double* __restrict__ tab_g, const double* __restrict__ tab_p,
                                      const double* __restrict__ cellJ, const int32_t* __restrict__ dofmap,
                                      const double* __restrict__ uh, const double* __restrict__ fh, double* __restrict__ G,
                                      double* __restrict__ F)
{
  extern __shared__ double s_tab[];
  double* s_g = s_tab;                  // [ndg][npk][2]
  double* s_p = s_tab + ndg * npk * 2;  // [ndg][npk]
  for (int i = threadIdx.x; i < ndg * npk * 2; i += blockDim.x)
    s_g[i] = tab_g[i];
  for (int i = threadIdx.x; i < ndg * npk; i += blockDim.x)
    s_p[i] = tab_p[i];
  __syncthreads();
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += gridDim.x * blockDim.x)
  {
    const int32_t* dm = dofmap + (size_t)c * npk;
    if (uh)
    {
      const double J00 = cellJ[4 * (size_t)c], J01 = cellJ[4 * (size_t)c + 1], J10 = cellJ[4 * (size_t)c + 2],
                   J11 = cellJ[4 * (size_t)c + 3];
      const double idet = 1.0 / (J00 * J11 - J01 * J10);
      const double K00 = J11 * idet, K01 = -J01 * idet, K10 = -J10 * idet, K11 = J00 * idet;
      double u[MAXPK];
      for (int j = 0; j < npk; ++j)
        u[j] = uh[dm[j]];
      for (int i = 0; i < ndg; ++i)
      {
        double gx = 0.0, gy = 0.0;
        for (int j = 0; j < npk; ++j)
        {
          gx += u[j] * s_g[(i * npk + j) * 2];
          gy += u[j] * s_g[(i * npk + j) * 2 + 1];
        }
        // grad u = K^T grad_ref u;  G = -grad u_h
        G[((size_t)c * ndg + i) * 2] = -(K00 * gx + K10 * gy);
        G[((size_t)c * ndg + i) * 2 + 1] = -(K01 * gx + K11 * gy);
      }
    }
    if (fh)
    {
      double f[MAXPK];
      for (int j = 0; j < npk; ++j)
        f[j] = fh[dm[j]];
      for (int i = 0; i < ndg; ++i)
      {
        double s = 0.0;
        for (int j = 0; j < npk; ++j)
          s += s_p[i * npk + j] * f[j];
        F[(size_t)c * ndg + i] = s;
      }
    }
  }
}

struct EstTables
{
  int nq, nrt, npk, ndg, k;
  const double *qw, *rt_q, *rt_div_q, *pk_q, *pk_gq, *pk_hq, *dg_q;
};

__device__ __forceinline__ void stage(double* dst, const double* src, int n)
{
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    dst[i] = src[i];
}

// cell-local hierarchic RT dofs of a cell from a DRT vector or a conforming vector (reflected facets: c_loc = R c_glob)
__device__ __forceinline__ void cell_rt_dofs(double* cl, int c, int k, int nrt, int nfct, bool is_ev, const double* __restrict__ sigma,
                                             const int32_t* __restrict__ cell_fct, const uint8_t* __restrict__ perms,
                                             const double* __restrict__ trafo)
{
  if (!is_ev)
  {
    for (int i = 0; i < nrt; ++i)
      cl[i] = sigma[(size_t)c * nrt + i];
    return;
  }
  const int ncd = k * k - k;
  for (int f = 0; f < 3; ++f)
  {
    const int32_t gf = cell_fct[3 * (size_t)c + f];
    double g[4];
    for (int j = 0; j < k; ++j)
      g[j] = sigma[(size_t)gf * k + j];
    if (perms[3 * (size_t)c + f])
      for (int i = 0; i < k; ++i)
      {
        double s = 0.0;
        for (int j = 0; j < k; ++j)
          s += trafo[j * k + i] * g[j];
        cl[f * k + i] = s;
      }
    else
      for (int j = 0; j < k; ++j)
        cl[f * k + j] = g[j];
  }
  for (int i = 0; i < ncd; ++i)
    cl[3 * k + i] = sigma[(size_t)nfct * k + (size_t)c * ncd + i];
}

__device__ __forceinline__ double cell_diameter(const double* __restrict__ x, const int32_t* __restrict__ cn)
{
  double h2 = 0.0;
  for (int a = 0; a < 3; ++a)
  {
    const int b = (a + 1) % 3;
    const double dx = x[3 * (size_t)cn[a]] - x[3 * (size_t)cn[b]], dy = x[3 * (size_t)cn[a] + 1] - x[3 * (size_t)cn[b] + 1];
    h2 = fmax(h2, dx * dx + dy * dy);
  }
  return sqrt(h2);  // dolfinx::mesh::h: largest vertex distance
}

__global__ void estimate_poisson_kernel(int ncell, int nfct, EstTables T, const double* __restrict__ trafo,
                                        const double* __restrict__ cellJ, const double* __restrict__ x,
                                        const int32_t* __restrict__ cell_node, const int32_t* __restrict__ cell_fct,
                                        const uint8_t* __restrict__ perms, const int32_t* __restrict__ dofmap,
                                        const double* __restrict__ sigma, const double* __restrict__ uh, const double* __restrict__ fh,
                                        double* __restrict__ eta_sig2, double* __restrict__ eta_osc2, int is_ev)
{
  extern __shared__ double s_tab[];
  double* s_qw = s_tab;
  double* s_rt = s_qw + T.nq;
  double* s_div = s_rt + T.nq * T.nrt * 2;
  double* s_pk = s_div + T.nq * T.nrt;
  double* s_pg = s_pk + T.nq * T.npk;
  double* s_ph = s_pg + T.nq * T.npk * 2;
  stage(s_qw, T.qw, T.nq);
  stage(s_rt, T.rt_q, T.nq * T.nrt * 2);
  stage(s_div, T.rt_div_q, T.nq * T.nrt);
  stage(s_pk, T.pk_q, T.nq * T.npk);
  stage(s_pg, T.pk_gq, T.nq * T.npk * 2);
  stage(s_ph, T.pk_hq, T.nq * T.npk * 3);
  __syncthreads();
  const double pi = 3.14159265358979323846;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += gridDim.x * blockDim.x)
  {
    const double J00 = cellJ[4 * (size_t)c], J01 = cellJ[4 * (size_t)c + 1], J10 = cellJ[4 * (size_t)c + 2],
                 J11 = cellJ[4 * (size_t)c + 3];
    const double det = J00 * J11 - J01 * J10, idet = 1.0 / det, adet = fabs(det);
    const double K00 = J11 * idet, K01 = -J01 * idet, K10 = -J10 * idet, K11 = J00 * idet;
    // Laplacian = sum_ab (K K^T)... : d_x = K00 d_X + K10 d_Y, d_y = K01 d_X + K11 d_Y
    const double a_xx = K00 * K00 + K01 * K01, a_xy = 2.0 * (K00 * K10 + K01 * K11), a_yy = K10 * K10 + K11 * K11;
    double cl[MAXRT], u[MAXPK], f[MAXPK];
    cell_rt_dofs(cl, c, T.k, T.nrt, nfct, is_ev != 0, sigma, cell_fct, perms, trafo);
    const int32_t* dm = dofmap + (size_t)c * T.npk;
    for (int j = 0; j < T.npk; ++j)
    {
      u[j] = uh ? uh[dm[j]] : 0.0;
      f[j] = fh ? fh[dm[j]] : 0.0;
    }
    double e_sig = 0.0, e_osc = 0.0;
    for (int q = 0; q < T.nq; ++q)
    {
      double s0 = 0.0, s1 = 0.0, dv = 0.0;
      for (int i = 0; i < T.nrt; ++i)
      {
        s0 += cl[i] * s_rt[(q * T.nrt + i) * 2];
        s1 += cl[i] * s_rt[(q * T.nrt + i) * 2 + 1];
        dv += cl[i] * s_div[q * T.nrt + i];
      }
      // contravariant Piola: sigma = J s / det, div sigma = div_ref / det
      const double sx = (J00 * s0 + J01 * s1) * idet, sy = (J10 * s0 + J11 * s1) * idet;
      double fq = 0.0, gX = 0.0, gY = 0.0, lap = 0.0;
      for (int j = 0; j < T.npk; ++j)
      {
        fq += f[j] * s_pk[q * T.npk + j];
        gX += u[j] * s_pg[(q * T.npk + j) * 2];
        gY += u[j] * s_pg[(q * T.npk + j) * 2 + 1];
        lap += u[j] * (a_xx * s_ph[(q * T.npk + j) * 3] + a_xy * s_ph[(q * T.npk + j) * 3 + 1] + a_yy * s_ph[(q * T.npk + j) * 3 + 2]);
      }
      const double ux = K00 * gX + K10 * gY, uy = K01 * gX + K11 * gY;
      const double w = s_qw[q] * adet;
      double ex, ey, res;
      if (is_ev)
      {
        ex = ux + sx;  // err_sig = grad u_h + sigma_eqlb, sigma = sigma_eqlb
        ey = uy + sy;
        res = fq - dv * idet;
      }
      else
      {
        ex = sx;  // err_sig = sigma_eqlb, sigma = sigma_eqlb - grad u_h
        ey = sy;
        res = fq - dv * idet + lap;
      }
      e_sig += w * (ex * ex + ey * ey);
      e_osc += w * res * res;
    }
    const double hT = cell_diameter(x, cell_node + 3 * (size_t)c);
    eta_sig2[c] = e_sig;
    eta_osc2[c] = (hT / pi) * (hT / pi) * e_osc;
  }
}

__global__ void estimate_elasticity_kernel(int ncell, EstTables T, const double* __restrict__ cellJ, const double* __restrict__ x,
                                           const int32_t* __restrict__ cell_node, const int32_t* __restrict__ dofmap,
                                           const double* __restrict__ ds0, const double* __restrict__ ds1,
                                           const double* __restrict__ sh0, const double* __restrict__ sh1,
                                           const double* __restrict__ f0, const double* __restrict__ f1,
                                           const double* __restrict__ korn, double pi_1, double* __restrict__ eta0,
                                           double* __restrict__ eta1, double* __restrict__ eta2)
{
  extern __shared__ double s_tab[];
  double* s_qw = s_tab;
  double* s_rt = s_qw + T.nq;
  double* s_div = s_rt + T.nq * T.nrt * 2;
  double* s_pk = s_div + T.nq * T.nrt;
  double* s_dg = s_pk + T.nq * T.npk;  // [3][nq][ndg]: value, d/dX, d/dY
  stage(s_qw, T.qw, T.nq);
  stage(s_rt, T.rt_q, T.nq * T.nrt * 2);
  stage(s_div, T.rt_div_q, T.nq * T.nrt);
  stage(s_pk, T.pk_q, T.nq * T.npk);
  stage(s_dg, T.dg_q, 3 * T.nq * T.ndg);
  __syncthreads();
  const double pi = 3.14159265358979323846;
  const double ctr = pi_1 / (2.0 + 2.0 * pi_1);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += gridDim.x * blockDim.x)
  {
    const double J00 = cellJ[4 * (size_t)c], J01 = cellJ[4 * (size_t)c + 1], J10 = cellJ[4 * (size_t)c + 2],
                 J11 = cellJ[4 * (size_t)c + 3];
    const double det = J00 * J11 - J01 * J10, idet = 1.0 / det, adet = fabs(det);
    const double K00 = J11 * idet, K01 = -J01 * idet, K10 = -J10 * idet, K11 = J00 * idet;
    const double ck = korn ? korn[c] : 1.0;
    const int32_t* dm = dofmap + (size_t)c * T.npk;
    double e0 = 0.0, e1 = 0.0, e2 = 0.0;
    for (int q = 0; q < T.nq; ++q)
    {
      double r[2][2], dvr[2], fq[2], dsh[2];
      for (int row = 0; row < 2; ++row)
      {
        const double* ds = (row ? ds1 : ds0) + (size_t)c * T.nrt;
        double s0 = 0.0, s1 = 0.0, dv = 0.0;
        for (int i = 0; i < T.nrt; ++i)
        {
          s0 += ds[i] * s_rt[(q * T.nrt + i) * 2];
          s1 += ds[i] * s_rt[(q * T.nrt + i) * 2 + 1];
          dv += ds[i] * s_div[q * T.nrt + i];
        }
        r[row][0] = (J00 * s0 + J01 * s1) * idet;
        r[row][1] = (J10 * s0 + J11 * s1) * idet;
        dvr[row] = dv * idet;
        // divergence of the projected stress row (DG_p^2, blocked bs = 2)
        const double* sh = (row ? sh1 : sh0);
        double d = 0.0;
        if (sh)
          for (int i = 0; i < T.ndg; ++i)
          {
            const double vx = sh[((size_t)c * T.ndg + i) * 2], vy = sh[((size_t)c * T.ndg + i) * 2 + 1];
            const double dX = s_dg[(1 * T.nq + q) * T.ndg + i], dY = s_dg[(2 * T.nq + q) * T.ndg + i];
            d += vx * (K00 * dX + K10 * dY) + vy * (K01 * dX + K11 * dY);
          }
        dsh[row] = d;
        const double* fr = row ? f1 : f0;
        double fv = 0.0;
        if (fr)
          for (int j = 0; j < T.npk; ++j)
            fv += fr[dm[j]] * s_pk[q * T.npk + j];
        fq[row] = fv;
      }
      const double w = s_qw[q] * adet;
      const double tr = r[0][0] + r[1][1];
      // inner(dsig, a(dsig)), a(s) = (s - ctr tr(s) I) / 2
      e0 += w * 0.5 * (r[0][0] * r[0][0] + r[0][1] * r[0][1] + r[1][0] * r[1][0] + r[1][1] * r[1][1] - ctr * tr * tr);
      const double ws = 0.5 * ck * (r[0][1] - r[1][0]);
      e1 += w * ws * ws;
      const double o0 = fq[0] + dsh[0] + dvr[0], o1 = fq[1] + dsh[1] + dvr[1];
      e2 += w * (o0 * o0 + o1 * o1);
    }
    const double hT = cell_diameter(x, cell_node + 3 * (size_t)c);
    eta0[c] = e0;
    eta1[c] = e1;
    eta2[c] = (ck * hT / pi) * (ck * hT / pi) * e2;
  }
}

EstTables est_tables(eqlb_handle* h)
{
  if (!h->d_primal.p)
    throw EqlbError(EQLB_ERR_STATE, "the tables carry no primal-space data (pk_*, rt_div_q)");
  EstTables T{};
  T.nq = h->nq;
  T.nrt = h->nrt;
  T.npk = h->npk;
  T.ndg = h->ndg;
  T.k = h->k;
  const double* p = h->d_primal.p;
  T.qw = p + h->o_pr[0];
  T.rt_q = p + h->o_pr[1];
  T.rt_div_q = p + h->o_pr[2];
  T.pk_q = p + h->o_pr[3];
  T.pk_gq = p + h->o_pr[4];
  T.pk_hq = p + h->o_pr[5];
  T.dg_q = p + h->o_pr[6];
  return T;
}
} // namespace

void launch_primal_project(eqlb_handle* h, int nfun, const double* const* uh, const double* const* fh, double* const* G,
                           double* const* F)
{
  if (!h->d_primal.p || !h->d_pk_dofmap.p)
    throw EqlbError(EQLB_ERR_STATE, "eqlb_project_primal: call eqlb_set_primal_space first (and pass tables with pk_* data)");
  if (h->npk > MAXPK || h->ndg > MAXDG)
    throw EqlbError(EQLB_ERR_INPUT, "eqlb_project_primal: degree not supported");
  const int bs = 128;
  const int grid = (int)std::min<size_t>(((size_t)h->ncell + bs - 1) / bs, (size_t)148 * 16);
  const size_t smem = (size_t)h->ndg * h->npk * 3 * sizeof(double);
  const double* p = h->d_primal.p;
  for (int f = 0; f < nfun; ++f)
  {
    primal_project_kernel<<<grid, bs, smem, h->stream>>>(h->ncell, h->npk, h->ndg, p + h->o_pr[7], p + h->o_pr[8], h->d_cellJ.p,
                                                         h->d_pk_dofmap.p, uh ? uh[f] : nullptr, fh ? fh[f] : nullptr,
                                                         G ? G[f] : nullptr, F ? F[f] : nullptr);
    CUDA_CHECK(cudaGetLastError());
    h->launches++;
  }
}

void launch_estimate_poisson(eqlb_handle* h, int nfun, const double* const* sigma, const double* const* uh, const double* const* fh,
                             double* const* e_sig, double* const* e_osc, int is_ev)
{
  if (!h->d_pk_dofmap.p)
    throw EqlbError(EQLB_ERR_STATE, "eqlb_estimate_poisson: call eqlb_set_primal_space first");
  const EstTables T = est_tables(h);
  if (T.npk > MAXPK || T.nrt > MAXRT)
    throw EqlbError(EQLB_ERR_INPUT, "eqlb_estimate_poisson: degree not supported");
  const int bs = 128;
  const int grid = (int)std::min<size_t>(((size_t)h->ncell + bs - 1) / bs, (size_t)148 * 16);
  const size_t smem = (size_t)(T.nq + T.nq * T.nrt * 3 + T.nq * T.npk * 6) * sizeof(double);
  for (int f = 0; f < nfun; ++f)
  {
    estimate_poisson_kernel<<<grid, bs, smem, h->stream>>>(h->ncell, h->nfct, T, h->tv.data + h->tv.o_trafo, h->d_cellJ.p, h->d_x.p,
                                                           h->d_cell_node.p, h->d_cell_fct.p, h->d_fct_perms.p, h->d_pk_dofmap.p,
                                                           sigma[f], uh ? uh[f] : nullptr, fh ? fh[f] : nullptr, e_sig[f], e_osc[f],
                                                           is_ev);
    CUDA_CHECK(cudaGetLastError());
    h->launches++;
  }
}

void launch_estimate_elasticity(eqlb_handle* h, const double* const* ds, const double* const* sh, const double* const* fh,
                                const double* korn, double pi_1, double* const* eta)
{
  if (!h->d_pk_dofmap.p)
    throw EqlbError(EQLB_ERR_STATE, "eqlb_estimate_elasticity: call eqlb_set_primal_space first");
  const EstTables T = est_tables(h);
  if (T.npk > MAXPK || T.nrt > MAXRT)
    throw EqlbError(EQLB_ERR_INPUT, "eqlb_estimate_elasticity: degree not supported");
  const int bs = 128;
  const int grid = (int)std::min<size_t>(((size_t)h->ncell + bs - 1) / bs, (size_t)148 * 16);
  const size_t smem = (size_t)(T.nq + T.nq * T.nrt * 3 + T.nq * T.npk + 3 * T.nq * T.ndg) * sizeof(double);
  estimate_elasticity_kernel<<<grid, bs, smem, h->stream>>>(h->ncell, T, h->d_cellJ.p, h->d_x.p, h->d_cell_node.p, h->d_pk_dofmap.p,
                                                            ds[0], ds[1], sh ? sh[0] : nullptr, sh ? sh[1] : nullptr,
                                                            fh ? fh[0] : nullptr, fh ? fh[1] : nullptr, korn, pi_1, eta[0], eta[1],
                                                            eta[2]);
  CUDA_CHECK(cudaGetLastError());
  h->launches++;
}
