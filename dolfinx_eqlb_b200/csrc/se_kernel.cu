// Semi-explicit equilibration, fused patch kernel (step 1 + step 2 + scatter).
//
// Reference semantics: se/solve_patch_semiexplt.hpp:212-1163 (+ assembly.hpp,
// fluxmin_kernel.hpp, PatchData.hpp, BoundaryData.cpp:686-745).  B200 design
// (DESIGN.md "SE kernel"):
//  * one thread per patch, patches of one colour per launch (no two patches of a
//    colour share a cell -> plain, deterministic += into the global DRT vector);
//  * no quadrature loops: on affine cells every integral of the reference is a
//    contraction of exact reference-cell tables with the cell Jacobian:
//      facet moments   m[j]  = sum_i W[f][v][j][i] (N_f . G_i),  N_f = adj(J)^T n_ref
//      cell moments    cm[t] = sum_i detJ f_i C[v][t][i] - (adj(J)^T D[v][t][i]) . G_i
//      RT mass matrix  M_c   = (g00 M00 + g01 (M01+M10) + g11 M11)/|detJ|, g = J^T J
//    tables live in shared memory (a few KB), inputs are gathered straight from
//    HBM/L2 (every cell is touched by its 3 vertex patches);
//  * the patch system (SPD, hz <= 1+(k-1) n_f + nadd n_c) is factorised by an
//    in-thread Cholesky.
#include <cstdio>
#include <cstdlib>

#include <algorithm>

#include "eqlb_internal.cuh"

namespace
{

__device__ __constant__ double c_nref[3][2] = {{-1.0, -1.0}, {-1.0, 0.0}, {0.0, 1.0}};

template <int K>
struct SeDims
{
  static constexpr int ndiv = K * (K + 1) / 2 - 1;
  static constexpr int nadd = (K - 1) * (K - 2) / 2;
  static constexpr int nrt = K * (K + 2);
  static constexpr int nact = 2 * K + nadd;         // rows of the cell block (facet + add functions)
  static constexpr int ncol = 2 * K + nadd + ndiv;  // columns (active coefficients)
  static constexpr int nz = 2 * K + nadd - 1;       // H(div=0) functions per cell
};

// packed lower-triangular index
__device__ __forceinline__ int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// EV = false: semi-explicit equilibration (corrector sigma_eq, DRT output)
// EV = true : constrained minimisation of FluxEqlbEV solved by the null-space method:
//             explicit conforming particular solution with div = Pi(hat f + grad(hat).G)
//             (+ the mean-value shift the reference's Lagrange multiplier produces),
//             then || sigma_p + Z u - hat G || -> min over the same patch-wise H(div=0)
//             basis Z; output into the conforming hierarchic RT vector.
template <int K, int NDG, int NCMAX, bool EV, bool STRESS>
__global__ void __launch_bounds__(128)
patch_kernel(PatchView pv, int first, int count, TableView tv, const double* __restrict__ cellJ,
             const int32_t* __restrict__ dgmap, int nrhs, RhsPtrs ptrs, const double* __restrict__ bflux,
             size_t bflux_stride, int use_atomics, const int32_t* __restrict__ cell_fct, int nfct, int mode)
{
  using D = SeDims<K>;
  constexpr int k = K, ndiv = D::ndiv, nadd = D::nadd, nrt = D::nrt, nact = D::nact, ncol = D::ncol, nz = D::nz;
  constexpr int HZ = 1 + (K - 1) * (NCMAX + 1) + nadd * NCMAX;
  constexpr int NT = 1 + ndiv;

  extern __shared__ double s_tab[];
  for (int i = threadIdx.x; i < tv.ndoubles; i += blockDim.x)
    s_tab[i] = tv.data[i];
  __syncthreads();
  const double* __restrict__ t_mass = s_tab + tv.o_rt_mass;    // [3][nrt][nrt]
  const double* __restrict__ t_fmom = s_tab + tv.o_fct_mom;    // [3][3][k][NDG]
  const double* __restrict__ t_cmf = s_tab + tv.o_cell_mom_f;  // [3][NT][NDG]
  const double* __restrict__ t_cmg = s_tab + tv.o_cell_mom_g;  // [3][NT][NDG][2]
  const double* __restrict__ t_bc = s_tab + tv.o_bc_mat;       // [3][3][k][k]
  const double* __restrict__ t_trafo = s_tab + tv.o_trafo;     // [k][k]
  const double* __restrict__ t_dgm = s_tab + tv.o_dg_mono;     // [NT][NDG]
  const double* __restrict__ t_hdr = s_tab + tv.o_hat_dg_rt;   // [3][NDG][nrt][2]
  const double* __restrict__ t_mono = s_tab + tv.o_mono_int;   // [NT]

  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= count)
    return;
  const size_t ip = (size_t)first + tid;
  const int nc = pv.ncells[ip];

  // ---- per-cell geometry / orientation (independent of the RHS) ----
  int32_t cell[NCMAX];
  uint16_t info[NCMAX];
  double gm[NCMAX][3];   // J^T J / |detJ|
  double adj[NCMAX][4];  // adj(J) = detJ * K
  double detJ[NCMAX];
  double pm[NCMAX], pp[NCMAX];  // prefactor_dof(a, 0/1)
#pragma unroll 1
  for (int a = 0; a < nc; ++a)
  {
    const int32_t c = pv.cell[(size_t)a * pv.stride + ip];
    cell[a] = c;
    const uint16_t inf = pv.info[(size_t)a * pv.stride + ip];
    info[a] = inf;
    const double2 j0 = reinterpret_cast<const double2*>(cellJ)[2 * (size_t)c];
    const double2 j1 = reinterpret_cast<const double2*>(cellJ)[2 * (size_t)c + 1];
    const double J00 = j0.x, J01 = j0.y, J10 = j1.x, J11 = j1.y;
    const double det = J00 * J11 - J01 * J10;
    const double iad = 1.0 / fabs(det);
    detJ[a] = det;
    adj[a][0] = J11;
    adj[a][1] = -J01;
    adj[a][2] = -J10;
    adj[a][3] = J00;
    gm[a][0] = (J00 * J00 + J10 * J10) * iad;
    gm[a][1] = (J00 * J01 + J10 * J11) * iad;
    gm[a][2] = (J01 * J01 + J11 * J11) * iad;
    const double sgn = det > 0.0 ? 1.0 : -1.0;
    const int fm = (inf >> 2) & 3, fp = (inf >> 4) & 3;
    pm[a] = (fm == 1) ? sgn : -sgn;  // reference normal outward only on facet 1
    pp[a] = (fp == 1) ? sgn : -sgn;
  }

  double mm[NCMAX][K], mp[NCMAX][K];  // own-side facet moments on E_{a-1} / E_a
  double cm[NCMAX][NT];               // cell moments
  double cf[NCMAX][ncol];             // sigma-tilde coefficients: [E_{a-1} (k)][E_a (k)][add][div]
  double cfin[STRESS ? 2 : 1][NCMAX][ncol];  // final patch coefficients of the two stress rows
  const double* __restrict__ t_p1 = s_tab + tv.o_rt_p1;  // [nrt][2][3]
  double A[HZ * (HZ + 1) / 2];
  double L[HZ];

  for (int r = 0; r < nrhs; ++r)
  {
    const double* __restrict__ G = ptrs.G[r];
    const double* __restrict__ Fv = ptrs.F[r];
    const uint8_t ri = pv.rhsinfo[(size_t)r * pv.stride + ip];
    const int ptype = ri & 3;
    const bool reversion = (ri & 4) != 0;
    const bool bc_e0 = (ri & 8) != 0, bc_en = (ri & 16) != 0;
    const bool internal = (ptype == EQLB_PATCH_INTERNAL);
    const int nf = internal ? nc : nc + 1;
    const int hz = 1 + (k - 1) * nf + nadd * nc;
    const int offset_En = nc * (k - 1);
    // mode 1 (grouped patches, se/reconstruction.hpp:170-234): no equilibration of this
    // patch, the stress coefficients come from the accumulated global vectors
    // (impose_weak_symmetry<modified_patch = true>, se/solve_patch_weaksym.hpp:100-131)
    if (mode == 0)
    {
    // ---- moments of every cell ----
#pragma unroll 1
    for (int a = 0; a < nc; ++a)
    {
      const int32_t c = cell[a];
      const int v = info[a] & 3, fm = (info[a] >> 2) & 3, fp = (info[a] >> 4) & 3;
      const double* ad = adj[a];
      // N_f = adj(J)^T n_ref[f]
      const double nmx = c_nref[fm][0] * ad[0] + c_nref[fm][1] * ad[2];
      const double nmy = c_nref[fm][0] * ad[1] + c_nref[fm][1] * ad[3];
      const double npx = c_nref[fp][0] * ad[0] + c_nref[fp][1] * ad[2];
      const double npy = c_nref[fp][0] * ad[1] + c_nref[fp][1] * ad[3];
#pragma unroll
      for (int j = 0; j < k; ++j)
      {
        mm[a][j] = 0.0;
        mp[a][j] = 0.0;
      }
#pragma unroll
      for (int t = 0; t < NT; ++t)
        cm[a][t] = 0.0;
      const double* wm = t_fmom + ((fm * 3 + v) * k) * NDG;
      const double* wp = t_fmom + ((fp * 3 + v) * k) * NDG;
      const double* cf_ = t_cmf + (v * NT) * NDG;
      const double* cg_ = t_cmg + (v * NT) * NDG * 2;
      // EV: reference gradient of the hat function of local vertex v
      const double ghx = (v == 0) ? -1.0 : (v == 1 ? 1.0 : 0.0);
      const double ghy = (v == 0) ? -1.0 : (v == 2 ? 1.0 : 0.0);
#pragma unroll
      for (int i = 0; i < NDG; ++i)
      {
        const size_t dof = dgmap ? (size_t)dgmap[(size_t)c * NDG + i] : (size_t)c * NDG + i;
        const double2 g = reinterpret_cast<const double2*>(G)[dof];
        const double fi = Fv[dof];
        // adj^T applied to reference gradients: d/dx_phys * detJ
        const double a0 = ad[0] * g.x + ad[1] * g.y;  // multiplies d/dxhat
        const double a1 = ad[2] * g.x + ad[3] * g.y;  // multiplies d/dyhat
        const double fd = detJ[a] * fi;
        if (EV)
        {
          // detJ * (hat f + grad(hat).G) tested with x^l y^m
          const double gg = a0 * ghx + a1 * ghy;
#pragma unroll
          for (int t = 0; t < NT; ++t)
            cm[a][t] += fd * cf_[t * NDG + i] + gg * t_dgm[t * NDG + i];
        }
        else
        {
          const double gnm = nmx * g.x + nmy * g.y;
          const double gnp = npx * g.x + npy * g.y;
#pragma unroll
          for (int j = 0; j < k; ++j)
          {
            mm[a][j] += wm[j * NDG + i] * gnm;
            mp[a][j] += wp[j * NDG + i] * gnp;
          }
#pragma unroll
          for (int t = 0; t < NT; ++t)
            cm[a][t] += fd * cf_[t * NDG + i] - a0 * cg_[(t * NDG + i) * 2] - a1 * cg_[(t * NDG + i) * 2 + 1];
        }
      }
    }
    if (EV && (ptype == EQLB_PATCH_INTERNAL || ptype == EQLB_PATCH_ESSNT_DUAL))
    {
      // mean-value shift: the KKT system of the reference carries a Lagrange multiplier
      // on the DG block (ev/assembly.hpp:283-298) which shifts the divergence data by a
      // patch constant so that it is compatible with the prescribed boundary fluxes
      double tot = 0.0, area2 = 0.0;
      for (int a = 0; a < nc; ++a)
      {
        tot += (detJ[a] > 0.0 ? cm[a][0] : -cm[a][0]);
        area2 += fabs(detJ[a]);
      }
      if (ptype == EQLB_PATCH_ESSNT_DUAL)
      {
        for (int e = 0; e < 2; ++e)
        {
          const int a = e ? nc - 1 : 0;
          const int v = info[a] & 3;
          const int fb = e ? (info[a] >> 4) & 3 : (info[a] >> 2) & 3;
          const double* bsrc = bflux + (size_t)r * bflux_stride + (size_t)cell[a] * nrt + fb * k;
          int nzero = 0;
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < k; ++i)
          {
            const double bi = bsrc[i];
            if (fabs(bi) < 1e-7)
              ++nzero;
            s += t_bc[((fb * 3 + v) * k) * k + i] * bi;
          }
          if (nzero < k)
            tot -= (e ? pp[a] : pm[a]) * s;
        }
      }
      const double lam = tot / (0.5 * area2);
#pragma unroll 1
      for (int a = 0; a < nc; ++a)
#pragma unroll
        for (int t = 0; t < NT; ++t)
          cm[a][t] -= lam * detJ[a] * t_mono[t];
    }

    // ---- step 1: explicit sweep ----
    double c_prev = 0.0, c_t1_e0 = 0.0;
#pragma unroll 1
    for (int a = 0; a < nc; ++a)
    {
#pragma unroll
      for (int i = 0; i < ncol; ++i)
        cf[a][i] = 0.0;
      const bool first_c = (a == 0), last_c = (a == nc - 1);
      const bool on_bnd = !internal && (first_c || last_c);
      bool has_bc = false;
      if (on_bnd)
      {
        if (ptype == EQLB_PATCH_ESSNT_DUAL)
          has_bc = true;
        else if (ptype == EQLB_PATCH_MIXED)
          has_bc = first_c ? bc_e0 : bc_en;  // (nc >= 2, so first/last are distinct cells)
      }
      const int v = info[a] & 3, fm = (info[a] >> 2) & 3, fp = (info[a] >> 4) & 3;
      const bool rev1 = (info[a] & 128) != 0;
      const double sgn = detJ[a] > 0.0 ? 1.0 : -1.0;

      double c_m = -c_prev;
      double surf = 0.0;

      // patch boundary condition (RT interpolant of hat * boundary flux)
      if (has_bc)
      {
        const int fb = first_c ? fm : fp;
        const double* bsrc = bflux + (size_t)r * bflux_stride + (size_t)cell[a] * nrt + fb * k;
        double b[K], bv[K];
        int nzero = 0;
#pragma unroll
        for (int i = 0; i < k; ++i)
        {
          b[i] = bsrc[i];
          if (fabs(b[i]) < 1e-7)
            ++nzero;
        }
#pragma unroll
        for (int j = 0; j < k; ++j)
        {
          double s = 0.0;
          if (nzero < k)
          {
#pragma unroll
            for (int i = 0; i < k; ++i)
              s += t_bc[((fb * 3 + v) * k + j) * k + i] * b[i];
          }
          bv[j] = s;
        }
        if (first_c)
          c_m += pm[a] * bv[0];
#pragma unroll
        for (int j = 1; j < k; ++j)
          cf[a][(first_c ? 0 : k) + j] += bv[j];
        if (reversion)
          c_t1_e0 -= pp[a] * bv[0];
      }

      // facet E_{a-1}: zero-order jump contribution
      if (!first_c)
      {
        // tau * m^+_{a-1}[0] - m^-_a[0], tau = -pp_{a-1} pm_a
        surf = -mm[a][0] - pp[a - 1] * pm[a] * mp[a - 1][0];
      }
      else if (!internal)
      {
        if (has_bc || ptype == EQLB_PATCH_MIXED)
        {
          const double sg = (ptype == EQLB_PATCH_MIXED && !has_bc) ? 1.0 : -1.0;
#pragma unroll
          for (int j = 1; j < k; ++j)
            cf[a][j] += sg * mm[a][j];
          if (has_bc)
            surf = -mm[a][0];
        }
      }
      c_m += pm[a] * surf;
      c_t1_e0 -= pm[a] * surf;

      // cell integral
      const double vol = sgn * cm[a][0];
      const double c_p = -c_m + vol;
      c_t1_e0 += vol;

      // boundary contribution to c_t1_e0 on the last facet (mixed patch whose E_0
      // lies on the Dirichlet boundary)
      if (on_bnd && last_c && reversion)
        c_t1_e0 -= pp[a] * (has_bc ? -1.0 : 1.0) * mp[a][0];

      // higher-order moments on E_a
      if (k > 1)
      {
        double h[K];
        if (on_bnd && last_c)
        {
          const double pf = has_bc ? -1.0 : 1.0;
#pragma unroll
          for (int j = 1; j < k; ++j)
            h[j] = pf * mp[a][j];
        }
        else
        {
          const int an = (a == nc - 1) ? 0 : a + 1;
          const double tau = -pp[a] * pm[an];
          double mt[K];
          if (rev1)
          {
            // moments of the T_{a+1}-side trace w.r.t. the facet parameter of T_a:
            // s' = 1 - s  ->  binomial transform
#pragma unroll
            for (int j = 0; j < k; ++j)
            {
              double s = 0.0;
              double binom = 1.0;
#pragma unroll
              for (int i = 0; i <= j; ++i)
              {
                s += ((i & 1) ? -binom : binom) * mm[an][i];
                binom = binom * (double)(j - i) / (double)(i + 1);
              }
              mt[j] = s;
            }
          }
          else
          {
#pragma unroll
            for (int j = 0; j < k; ++j)
              mt[j] = mm[an][j];
          }
          const bool corr = rev1 && !last_c;
          const double j0 = tau * mt[0] - mp[a][0];
#pragma unroll
          for (int j = 1; j < k; ++j)
          {
            h[j] = tau * mt[j] - mp[a][j];
            if (corr)
              h[j] += -j0 + pm[an] * c_p;
          }
        }
#pragma unroll
        for (int j = 1; j < k; ++j)
          cf[a][k + j] += h[j];
#pragma unroll
        for (int t = 0; t < ndiv; ++t)
          cf[a][2 * k + nadd + t] += cm[a][1 + t];
      }
      cf[a][0] += pm[a] * c_m;
      cf[a][k] += pp[a] * c_p;
      c_prev = c_p;
    }
    if (reversion)
    {
#pragma unroll 1
      for (int a = 0; a < nc; ++a)
      {
        cf[a][0] += pm[a] * c_t1_e0;
        cf[a][k] -= pp[a] * c_t1_e0;
        if ((info[a] & 128) && a != nc - 1)
#pragma unroll
          for (int j = 1; j < k; ++j)
            cf[a][k + j] -= pm[a + 1] * c_t1_e0;
      }
    }

    // ---- step 2: assemble the patch system ----
    const bool req_bc = (ptype == EQLB_PATCH_ESSNT_DUAL || ptype == EQLB_PATCH_MIXED);
    auto marked = [&](int pd) -> bool
    {
      if (!req_bc)
        return false;
      if (pd == 0)
        return true;
      if (pd >= 1 + (k - 1) * nf)
        return false;
      const bool onE0 = pd < k;
      const bool onEn = pd > offset_En && pd < offset_En + k;
      if (ptype == EQLB_PATCH_ESSNT_DUAL)
        return onE0 || onEn;
      return reversion ? onEn : onE0;
    };
    for (int i = 0; i < hz * (hz + 1) / 2; ++i)
      A[i] = 0.0;
    for (int i = 0; i < hz; ++i)
      L[i] = 0.0;

#pragma unroll 1
    for (int a = 0; a < nc; ++a)
    {
      const int fm = (info[a] >> 2) & 3, fp = (info[a] >> 4) & 3;
      const bool rev0 = (info[a] & 64) != 0;
      // reference dof of active row/col q
      auto rdof = [&](int q) -> int
      {
        if (q < k)
          return fm * k + q;
        if (q < 2 * k)
          return fp * k + (q - k);
        if (q < 2 * k + nadd)
          return 3 * k + ndiv + (q - 2 * k);
        return 3 * k + (q - 2 * k - nadd);
      };
      // cell block MB[nact][ncol] of the physical RT mass matrix
      double MB[nact][ncol];
      const double g0 = gm[a][0], g1 = gm[a][1], g2 = gm[a][2];
#pragma unroll
      for (int q = 0; q < nact; ++q)
      {
        const int rq = rdof(q);
#pragma unroll
        for (int s = 0; s < ncol; ++s)
        {
          const int rs = rdof(s);
          const int o = rq * nrt + rs;
          MB[q][s] = g0 * t_mass[o] + g1 * t_mass[nrt * nrt + o] + g2 * t_mass[2 * nrt * nrt + o];
        }
      }
      // y = MB * cf  (untransformed rows)
      double y[nact];
#pragma unroll
      for (int q = 0; q < nact; ++q)
      {
        double s = 0.0;
#pragma unroll
        for (int c2 = 0; c2 < ncol; ++c2)
          s += MB[q][c2] * cf[a][c2];
        y[q] = s;
      }
      if (EV)
      {
        // y -= (hat G, phi_q)_T = sgn * sum_m (J^T G_m) . H[v][m][rdof(q)]
        const int v = info[a] & 3;
        const double sgn = detJ[a] > 0.0 ? 1.0 : -1.0;
        const double* ad = adj[a];
#pragma unroll
        for (int mI = 0; mI < NDG; ++mI)
        {
          const size_t dof = dgmap ? (size_t)dgmap[(size_t)cell[a] * NDG + mI] : (size_t)cell[a] * NDG + mI;
          const double2 g = reinterpret_cast<const double2*>(G)[dof];
          // J = [[adj3, -adj1], [-adj2, adj0]]
          const double jg0 = sgn * (ad[3] * g.x - ad[2] * g.y);
          const double jg1 = sgn * (-ad[1] * g.x + ad[0] * g.y);
          const double* hh = t_hdr + ((size_t)(v * NDG + mI) * nrt) * 2;
#pragma unroll
          for (int q = 0; q < nact; ++q)
          {
            const int rq = rdof(q);
            y[q] -= jg0 * hh[rq * 2] + jg1 * hh[rq * 2 + 1];
          }
        }
      }
      // reversed E_{a-1}: transform the test/trial functions of that facet
      if (rev0)
      {
        double tmp[K];
        // rows
#pragma unroll
        for (int s = 0; s < nact; ++s)
        {
#pragma unroll
          for (int i = 0; i < k; ++i)
          {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < k; ++j)
              acc += t_trafo[i * k + j] * MB[j][s];
            tmp[i] = acc;
          }
#pragma unroll
          for (int i = 0; i < k; ++i)
            MB[i][s] = tmp[i];
        }
        // columns
#pragma unroll
        for (int q = 0; q < nact; ++q)
        {
#pragma unroll
          for (int i = 0; i < k; ++i)
          {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < k; ++j)
              acc += t_trafo[i * k + j] * MB[q][j];
            tmp[i] = acc;
          }
#pragma unroll
          for (int i = 0; i < k; ++i)
            MB[q][i] = tmp[i];
        }
#pragma unroll
        for (int i = 0; i < k; ++i)
        {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < k; ++j)
            acc += t_trafo[i * k + j] * y[j];
          tmp[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < k; ++i)
          y[i] = tmp[i];
      }
      // signs: p_ea = -pp_a ; p_eam1 = rev0 ? -pp_{a-1} : pm_a
      const int am = (a == 0) ? nc - 1 : a - 1;
      const double p_ea = -pp[a];
      const double p_em = rev0 ? -pp[am] : pm[a];
      // sub-basis function i (0..nz-1) = sum of up to two signed active rows
      //   i <  k-1      : p_em * row(i+1)
      //   i == k-1      : p_em * row(0) + p_ea * row(k)          (d0)
      //   k <= i < 2k-1 : p_ea * row(i+1)
      //   i >= 2k-1     : row(i+1)                                (additional)
      auto sgn_of = [&](int q) -> double { return q < k ? p_em : (q < 2 * k ? p_ea : 1.0); };
      // patch-local dof of sub-basis function i
      auto pdof = [&](int i) -> int
      {
        if (i < k - 1)
          return a * (k - 1) + i + 1;
        if (i == k - 1)
          return 0;
        if (i < 2 * k - 1)
          return ((internal && a == nc - 1) ? 0 : (a + 1) * (k - 1)) + (i - k + 1);
        return nf * (k - 1) + 1 + a * nadd + (i - (2 * k - 1));
      };
      // Te(i, j) and load(i)
#pragma unroll
      for (int i = 0; i < nz; ++i)
      {
        const int qi = i + 1;
        double li = (i == k - 1) ? -(p_em * y[0] + p_ea * y[k]) : -sgn_of(qi) * y[qi];
        const int di = pdof(i);
        const bool bi = marked(di);
        if (bi)
        {
          L[di] = 0.0;
          A[tri(di, di)] = 1.0;
          continue;
        }
        L[di] += li;
#pragma unroll
        for (int j = 0; j < nz; ++j)
        {
          const int dj = pdof(j);
          if (dj > di)
            continue;  // lower triangle only (A symmetric)
          if (marked(dj))
            continue;
          const int qj = j + 1;
          double val;
          if (i == k - 1 && j == k - 1)
            val = MB[0][0] + MB[k][k] + 2.0 * p_em * p_ea * MB[0][k];
          else if (i == k - 1)
            val = sgn_of(qj) * (p_em * MB[0][qj] + p_ea * MB[k][qj]);
          else if (j == k - 1)
            val = sgn_of(qi) * (p_em * MB[qi][0] + p_ea * MB[qi][k]);
          else
            val = sgn_of(qi) * sgn_of(qj) * MB[qi][qj];
          // each unordered pair of distinct patch dofs appears twice per cell (i,j)
          // and (j,i): keep the one with dj <= di; pairs with di == dj and i != j
          // cannot occur (distinct functions of a cell have distinct patch dofs)
          A[tri(di, dj)] += val;
        }
      }
    }

    // ---- Cholesky solve ----
    if (k == 1)
    {
      L[0] = L[0] / A[0];
    }
    else
    {
      for (int j = 0; j < hz; ++j)
      {
        double d = A[tri(j, j)];
        for (int q = 0; q < j; ++q)
          d -= A[tri(j, q)] * A[tri(j, q)];
        d = sqrt(d);
        A[tri(j, j)] = d;
        const double id = 1.0 / d;
        for (int i = j + 1; i < hz; ++i)
        {
          double s = A[tri(i, j)];
          for (int q = 0; q < j; ++q)
            s -= A[tri(i, q)] * A[tri(j, q)];
          A[tri(i, j)] = s * id;
        }
      }
      for (int i = 0; i < hz; ++i)
      {
        double s = L[i];
        for (int q = 0; q < i; ++q)
          s -= A[tri(i, q)] * L[q];
        L[i] = s / A[tri(i, i)];
      }
      for (int i = hz - 1; i >= 0; --i)
      {
        double s = L[i];
        for (int q = i + 1; q < hz; ++q)
          s -= A[tri(q, i)] * L[q];
        L[i] = s / A[tri(i, i)];
      }
    }

    // ---- map back to cell coefficients and accumulate ----
    double* __restrict__ sig = ptrs.S[r];
#pragma unroll 1
    for (int a = 0; a < nc; ++a)
    {
      const int fm = (info[a] >> 2) & 3, fp = (info[a] >> 4) & 3;
      const bool rev0 = (info[a] & 64) != 0;
      const int am = (a == 0) ? nc - 1 : a - 1;
      const double p_ea = -pp[a];
      const double p_em = rev0 ? -pp[am] : pm[a];
      // slot -> patch dof (slot 0 and slot k are d0)
      double um[K], up[K];
#pragma unroll
      for (int j = 0; j < k; ++j)
      {
        const int pd_m = (j == 0) ? 0 : a * (k - 1) + j;
        const int pd_p = (j == 0) ? 0 : ((internal && a == nc - 1) ? 0 : (a + 1) * (k - 1)) + j;
        um[j] = p_em * L[pd_m];
        up[j] = p_ea * L[pd_p];
      }
      if (rev0)
      {
        double tmp[K];
#pragma unroll
        for (int i = 0; i < k; ++i)
        {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < k; ++j)
            acc += t_trafo[j * k + i] * um[j];
          tmp[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < k; ++i)
          um[i] = tmp[i];
      }
      if (EV)
      {
        // conforming hierarchic RT vector: [fct*k + j] facet dofs in the global (low ->
        // high vertex) orientation, then [nfct*k + cell*(k*k-k) + i] cell dofs.
        // T_a owns facet E_a; T_1 of a boundary patch also owns E_0.
        constexpr int ncd = k * k - k;
        const int32_t c = cell[a];
        for (int side = (a == 0 && !internal) ? 0 : 1; side < 2; ++side)
        {
          const int fl = side ? fp : fm;
          const bool refl = (info[a] & (side ? 512 : 256)) != 0;
          const int32_t fg = cell_fct[3 * (size_t)c + fl];
          double cl[K], cg[K];
#pragma unroll
          for (int j = 0; j < k; ++j)
            cl[j] = side ? (cf[a][k + j] + up[j]) : (cf[a][j] + um[j]);
#pragma unroll
          for (int i = 0; i < k; ++i)
          {
            double acc = cl[i];
            if (refl)
            {
              acc = 0.0;
#pragma unroll
              for (int j = 0; j < k; ++j)
                acc += t_trafo[j * k + i] * cl[j];
            }
            cg[i] = acc;
          }
#pragma unroll
          for (int i = 0; i < k; ++i)
          {
            if (use_atomics)
              atomicAdd(sig + (size_t)fg * k + i, cg[i]);
            else
              sig[(size_t)fg * k + i] += cg[i];
          }
        }
        double* dstc = sig + (size_t)nfct * k + (size_t)c * ncd;
#pragma unroll
        for (int t = 0; t < ndiv; ++t)
        {
          if (use_atomics)
            atomicAdd(dstc + t, cf[a][2 * k + nadd + t]);
          else
            dstc[t] += cf[a][2 * k + nadd + t];
        }
#pragma unroll
        for (int i = 0; i < nadd; ++i)
        {
          const double val = cf[a][2 * k + i] + L[nf * (k - 1) + 1 + a * nadd + i];
          if (use_atomics)
            atomicAdd(dstc + ndiv + i, val);
          else
            dstc[ndiv + i] += val;
        }
        continue;
      }
      if (STRESS && r < 2)
      {
#pragma unroll
        for (int j = 0; j < k; ++j)
        {
          cfin[r][a][j] = cf[a][j] + um[j];
          cfin[r][a][k + j] = cf[a][k + j] + up[j];
        }
#pragma unroll
        for (int i = 0; i < nadd; ++i)
          cfin[r][a][2 * k + i] = cf[a][2 * k + i] + L[nf * (k - 1) + 1 + a * nadd + i];
#pragma unroll
        for (int t = 0; t < ndiv; ++t)
          cfin[r][a][2 * k + nadd + t] = cf[a][2 * k + nadd + t];
      }
      double* dst = sig + (size_t)cell[a] * nrt;
      if (use_atomics)
      {
#pragma unroll
        for (int j = 0; j < k; ++j)
        {
          atomicAdd(dst + fm * k + j, cf[a][j] + um[j]);
          atomicAdd(dst + fp * k + j, cf[a][k + j] + up[j]);
        }
#pragma unroll
        for (int i = 0; i < nadd; ++i)
          atomicAdd(dst + 3 * k + ndiv + i, cf[a][2 * k + i] + L[nf * (k - 1) + 1 + a * nadd + i]);
#pragma unroll
        for (int t = 0; t < ndiv; ++t)
          atomicAdd(dst + 3 * k + t, cf[a][2 * k + nadd + t]);
      }
      else
      {
#pragma unroll
        for (int j = 0; j < k; ++j)
        {
          dst[fm * k + j] += cf[a][j] + um[j];
          dst[fp * k + j] += cf[a][k + j] + up[j];
        }
#pragma unroll
        for (int i = 0; i < nadd; ++i)
          dst[3 * k + ndiv + i] += cf[a][2 * k + i] + L[nf * (k - 1) + 1 + a * nadd + i];
#pragma unroll
        for (int t = 0; t < ndiv; ++t)
          dst[3 * k + t] += cf[a][2 * k + nadd + t];
      }
    }
    }  // mode == 0

    // ---- weak symmetry of the stress rows 0/1 (se/solve_patch_weaksym.hpp:59-233) ----
    if constexpr (STRESS && !EV)
    {
      if (r == 1)
      {
        constexpr int NCS = NCMAX + 3;  // P1 constraint dofs (n_f + 1 <= NCMAX + 2) + mean-value multiplier
        const int ncs = nf + 1;
        int rtype[2];
        bool rrev[2];
        bool need_bc = false, lagr = true;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
        {
          const uint8_t q = pv.rhsinfo[(size_t)rr * pv.stride + ip];
          rtype[rr] = q & 3;
          rrev[rr] = (q & 4) != 0;
          if (rtype[rr] == EQLB_PATCH_ESSNT_DUAL || rtype[rr] == EQLB_PATCH_MIXED)
            need_bc = true;
          if (rtype[rr] == EQLB_PATCH_ESSNT_PRIMAL || rtype[rr] == EQLB_PATCH_MIXED)
            lagr = false;  // se/PatchData.hpp:189-205
        }
        auto marked_row = [&](int rr, int pd) -> bool
        {
          const int tp = rtype[rr];
          if (!(tp == EQLB_PATCH_ESSNT_DUAL || tp == EQLB_PATCH_MIXED))
            return false;
          if (pd == 0)
            return true;
          if (pd >= 1 + (k - 1) * nf)
            return false;
          const bool onE0 = pd < k;
          const bool onEn = pd > offset_En && pd < offset_En + k;
          if (tp == EQLB_PATCH_ESSNT_DUAL)
            return onE0 || onEn;
          return rrev[rr] ? onEn : onE0;
        };
        double Bm[2][HZ][NCS - 1];
        double Cm[NCS][NCS];
        double Lc[NCS];
        double Fk[2][HZ * (HZ + 1) / 2];
        for (int i = 0; i < hz; ++i)
          for (int c2 = 0; c2 < ncs; ++c2)
          {
            Bm[0][i][c2] = 0.0;
            Bm[1][i][c2] = 0.0;
          }
        for (int i = 0; i <= ncs; ++i)
        {
          Lc[i] = 0.0;
          for (int c2 = 0; c2 <= ncs; ++c2)
            Cm[i][c2] = 0.0;
        }
        for (int i = 0; i < hz * (hz + 1) / 2; ++i)
          A[i] = 0.0;

#pragma unroll 1
        for (int a = 0; a < nc; ++a)
        {
          const int v = info[a] & 3, fm = (info[a] >> 2) & 3, fp = (info[a] >> 4) & 3;
          const bool rev0 = (info[a] & 64) != 0;
          if (mode == 1)
          {
            // accumulated global stress on the cell; the dofs of the third facet (which
            // other patches of the group have filled) enter L_c below
            const double sgn3 = detJ[a] > 0.0 ? 1.0 : -1.0;
            const double* ad3 = adj[a];
            const int f3 = 3 - fm - fp;
            const int vl3[3] = {v, 3 - fp - v, 3 - fm - v};
            const int cd3[3] = {0, (!internal && a == nc - 1) ? nf : a + 1, (a == 0) ? (internal ? nc : nf - 1) : a};
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
            {
              const double* src = ptrs.S[rr] + (size_t)cell[a] * nrt;
#pragma unroll
              for (int j = 0; j < k; ++j)
              {
                cfin[rr][a][j] = src[fm * k + j];
                cfin[rr][a][k + j] = src[fp * k + j];
              }
#pragma unroll
              for (int i = 0; i < nadd; ++i)
                cfin[rr][a][2 * k + i] = src[3 * k + ndiv + i];
#pragma unroll
              for (int t = 0; t < ndiv; ++t)
                cfin[rr][a][2 * k + nadd + t] = src[3 * k + t];
#pragma unroll
              for (int j = 0; j < k; ++j)
              {
                const double c3 = src[f3 * k + j];
                const double* tp = t_p1 + (f3 * k + j) * 6;
#pragma unroll
                for (int jj = 0; jj < 3; ++jj)
                {
                  const double px = sgn3 * (ad3[3] * tp[vl3[jj]] - ad3[1] * tp[3 + vl3[jj]]);
                  const double py = sgn3 * (-ad3[2] * tp[vl3[jj]] + ad3[0] * tp[3 + vl3[jj]]);
                  Lc[cd3[jj]] -= (rr == 0) ? c3 * py : -c3 * px;
                }
              }
            }
          }
          auto rdof = [&](int q) -> int
          {
            if (q < k)
              return fm * k + q;
            if (q < 2 * k)
              return fp * k + (q - k);
            if (q < 2 * k + nadd)
              return 3 * k + ndiv + (q - 2 * k);
            return 3 * k + (q - 2 * k - nadd);
          };
          // constraint slots: patch node, outer node of E_a, outer node of E_{a-1}
          const int vl[3] = {v, 3 - fp - v, 3 - fm - v};
          const int cd[3] = {0, (!internal && a == nc - 1) ? nf : a + 1, (a == 0) ? (internal ? nc : nf - 1) : a};
          const double sgn = detJ[a] > 0.0 ? 1.0 : -1.0;
          const double* ad = adj[a];
          const double J00 = sgn * ad[3], J01 = -sgn * ad[1], J10 = -sgn * ad[2], J11 = sgn * ad[0];
          // P[q][d][j] = int_T phi_q^d lambda_j  (physical)
          double P[ncol][2][3];
#pragma unroll
          for (int q = 0; q < ncol; ++q)
          {
            const double* tp = t_p1 + rdof(q) * 6;
#pragma unroll
            for (int j = 0; j < 3; ++j)
            {
              P[q][0][j] = J00 * tp[vl[j]] + J01 * tp[3 + vl[j]];
              P[q][1][j] = J10 * tp[vl[j]] + J11 * tp[3 + vl[j]];
            }
          }
          // L_c = -(sigma_01 - sigma_10, psi), mean-value row
          const double ce = fabs(detJ[a]) * (1.0 / 6.0);
#pragma unroll
          for (int j = 0; j < 3; ++j)
          {
            double sj = 0.0;
#pragma unroll
            for (int q = 0; q < ncol; ++q)
              sj += cfin[0][a][q] * P[q][1][j] - cfin[1][a][q] * P[q][0][j];
            Lc[cd[j]] -= sj;
            if (lagr)
            {
              Cm[cd[j]][ncs] += ce;
              Cm[ncs][cd[j]] += ce;
            }
          }
          // mass block of the H(div=0) functions (no BC masks: A_rec)
          double MB[nact][nact];
          const double g0 = gm[a][0], g1 = gm[a][1], g2 = gm[a][2];
#pragma unroll
          for (int q = 0; q < nact; ++q)
#pragma unroll
            for (int s2 = 0; s2 < nact; ++s2)
            {
              const int o = rdof(q) * nrt + rdof(s2);
              MB[q][s2] = g0 * t_mass[o] + g1 * t_mass[nrt * nrt + o] + g2 * t_mass[2 * nrt * nrt + o];
            }
          if (rev0)
          {
            double tmp[K];
#pragma unroll
            for (int s2 = 0; s2 < nact; ++s2)
            {
#pragma unroll
              for (int i = 0; i < k; ++i)
              {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < k; ++j)
                  acc += t_trafo[i * k + j] * MB[j][s2];
                tmp[i] = acc;
              }
#pragma unroll
              for (int i = 0; i < k; ++i)
                MB[i][s2] = tmp[i];
            }
#pragma unroll
            for (int q = 0; q < nact; ++q)
            {
#pragma unroll
              for (int i = 0; i < k; ++i)
              {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < k; ++j)
                  acc += t_trafo[i * k + j] * MB[q][j];
                tmp[i] = acc;
              }
#pragma unroll
              for (int i = 0; i < k; ++i)
                MB[q][i] = tmp[i];
            }
#pragma unroll
            for (int d = 0; d < 2; ++d)
#pragma unroll
              for (int j = 0; j < 3; ++j)
              {
#pragma unroll
                for (int i = 0; i < k; ++i)
                {
                  double acc = 0.0;
#pragma unroll
                  for (int q = 0; q < k; ++q)
                    acc += t_trafo[i * k + q] * P[q][d][j];
                  tmp[i] = acc;
                }
#pragma unroll
                for (int i = 0; i < k; ++i)
                  P[i][d][j] = tmp[i];
              }
          }
          const int am = (a == 0) ? nc - 1 : a - 1;
          const double p_ea = -pp[a];
          const double p_em = rev0 ? -pp[am] : pm[a];
          auto sgn_of = [&](int q) -> double { return q < k ? p_em : (q < 2 * k ? p_ea : 1.0); };
          auto pdof = [&](int i) -> int
          {
            if (i < k - 1)
              return a * (k - 1) + i + 1;
            if (i == k - 1)
              return 0;
            if (i < 2 * k - 1)
              return ((internal && a == nc - 1) ? 0 : (a + 1) * (k - 1)) + (i - k + 1);
            return nf * (k - 1) + 1 + a * nadd + (i - (2 * k - 1));
          };
#pragma unroll
          for (int i = 0; i < nz; ++i)
          {
            const int qi = i + 1;
            const int di = pdof(i);
            // B_1 = (psi_y, lambda), B_2 = -(psi_x, lambda)   (se/stressmin_kernel.hpp:214-222)
#pragma unroll
            for (int j = 0; j < 3; ++j)
            {
              double py, px;
              if (i == k - 1)
              {
                py = p_em * P[0][1][j] + p_ea * P[k][1][j];
                px = p_em * P[0][0][j] + p_ea * P[k][0][j];
              }
              else
              {
                py = sgn_of(qi) * P[qi][1][j];
                px = sgn_of(qi) * P[qi][0][j];
              }
              Bm[0][di][cd[j]] += py;
              Bm[1][di][cd[j]] -= px;
            }
#pragma unroll
            for (int j = 0; j < nz; ++j)
            {
              const int dj = pdof(j);
              if (dj > di)
                continue;
              const int qj = j + 1;
              double val;
              if (i == k - 1 && j == k - 1)
                val = MB[0][0] + MB[k][k] + 2.0 * p_em * p_ea * MB[0][k];
              else if (i == k - 1)
                val = sgn_of(qj) * (p_em * MB[0][qj] + p_ea * MB[k][qj]);
              else if (j == k - 1)
                val = sgn_of(qi) * (p_em * MB[qi][0] + p_ea * MB[qi][k]);
              else
                val = sgn_of(qi) * sgn_of(qj) * MB[qi][qj];
              A[tri(di, dj)] += val;
            }
          }
        }

        // Schur complement C -= B_k^T A_k^-1 B_k  (se/PatchData.hpp:604-628)
        double xcol[HZ];
        for (int kk = 0; kk < 2; ++kk)
        {
          double* F = Fk[(need_bc && kk == 1) ? 1 : 0];
          if (kk == 0 || need_bc)
          {
            for (int i = 0; i < hz; ++i)
              for (int j = 0; j <= i; ++j)
              {
                const bool mi = marked_row(kk, i), mj = marked_row(kk, j);
                F[tri(i, j)] = (mi || mj) ? ((i == j) ? 1.0 : 0.0) : A[tri(i, j)];
              }
            for (int j = 0; j < hz; ++j)
            {
              double d = F[tri(j, j)];
              for (int q = 0; q < j; ++q)
                d -= F[tri(j, q)] * F[tri(j, q)];
              d = sqrt(d);
              F[tri(j, j)] = d;
              const double id = 1.0 / d;
              for (int i = j + 1; i < hz; ++i)
              {
                double s2 = F[tri(i, j)];
                for (int q = 0; q < j; ++q)
                  s2 -= F[tri(i, q)] * F[tri(j, q)];
                F[tri(i, j)] = s2 * id;
              }
            }
          }
          for (int c2 = 0; c2 < ncs; ++c2)
          {
            for (int i = 0; i < hz; ++i)
              xcol[i] = marked_row(kk, i) ? 0.0 : Bm[kk][i][c2];
            for (int i = 0; i < hz; ++i)
            {
              double s2 = xcol[i];
              for (int q = 0; q < i; ++q)
                s2 -= F[tri(i, q)] * xcol[q];
              xcol[i] = s2 / F[tri(i, i)];
            }
            for (int i = hz - 1; i >= 0; --i)
            {
              double s2 = xcol[i];
              for (int q = i + 1; q < hz; ++q)
                s2 -= F[tri(q, i)] * xcol[q];
              xcol[i] = s2 / F[tri(i, i)];
            }
            for (int rr = 0; rr < ncs; ++rr)
            {
              double s2 = 0.0;
              for (int i = 0; i < hz; ++i)
                s2 += (marked_row(kk, i) ? 0.0 : Bm[kk][i][rr]) * xcol[i];
              Cm[rr][c2] -= s2;
            }
          }
        }
        // partial-pivot LU of the Schur complement (se/PatchData.hpp:630-637)
        const int dim_c = lagr ? ncs + 1 : ncs;
        for (int p2 = 0; p2 < dim_c; ++p2)
        {
          int piv = p2;
          double mx = fabs(Cm[p2][p2]);
          for (int i = p2 + 1; i < dim_c; ++i)
            if (fabs(Cm[i][p2]) > mx)
            {
              mx = fabs(Cm[i][p2]);
              piv = i;
            }
          if (piv != p2)
          {
            for (int j = 0; j < dim_c; ++j)
            {
              const double tmpv = Cm[p2][j];
              Cm[p2][j] = Cm[piv][j];
              Cm[piv][j] = tmpv;
            }
            const double tl = Lc[p2];
            Lc[p2] = Lc[piv];
            Lc[piv] = tl;
          }
          const double dinv = 1.0 / Cm[p2][p2];
          for (int i = p2 + 1; i < dim_c; ++i)
          {
            const double l = Cm[i][p2] * dinv;
            for (int j = p2 + 1; j < dim_c; ++j)
              Cm[i][j] -= l * Cm[p2][j];
            Lc[i] -= l * Lc[p2];
          }
        }
        for (int i = dim_c - 1; i >= 0; --i)
        {
          double s2 = Lc[i];
          for (int j = i + 1; j < dim_c; ++j)
            s2 -= Cm[i][j] * Lc[j];
          Lc[i] = s2 / Cm[i][i];
        }
        // u_k = A_k^-1 (-B_k u_c) and scatter into both rows (:170-232)
        for (int kk = 0; kk < 2; ++kk)
        {
          const double* F = Fk[(need_bc && kk == 1) ? 1 : 0];
          for (int i = 0; i < hz; ++i)
          {
            double s2 = 0.0;
            if (!marked_row(kk, i))
              for (int c2 = 0; c2 < ncs; ++c2)
                s2 -= Bm[kk][i][c2] * Lc[c2];
            xcol[i] = s2;
          }
          for (int i = 0; i < hz; ++i)
          {
            double s2 = xcol[i];
            for (int q = 0; q < i; ++q)
              s2 -= F[tri(i, q)] * xcol[q];
            xcol[i] = s2 / F[tri(i, i)];
          }
          for (int i = hz - 1; i >= 0; --i)
          {
            double s2 = xcol[i];
            for (int q = i + 1; q < hz; ++q)
              s2 -= F[tri(q, i)] * xcol[q];
            xcol[i] = s2 / F[tri(i, i)];
          }
          double* __restrict__ sg = ptrs.S[kk];
#pragma unroll 1
          for (int a = 0; a < nc; ++a)
          {
            const int fm = (info[a] >> 2) & 3, fp = (info[a] >> 4) & 3;
            const bool rev0 = (info[a] & 64) != 0;
            const int am = (a == 0) ? nc - 1 : a - 1;
            const double p_ea = -pp[a];
            const double p_em = rev0 ? -pp[am] : pm[a];
            double um[K], up[K];
#pragma unroll
            for (int j = 0; j < k; ++j)
            {
              const int pd_m = (j == 0) ? 0 : a * (k - 1) + j;
              const int pd_p = (j == 0) ? 0 : ((internal && a == nc - 1) ? 0 : (a + 1) * (k - 1)) + j;
              um[j] = p_em * xcol[pd_m];
              up[j] = p_ea * xcol[pd_p];
            }
            if (rev0)
            {
              double tmp[K];
#pragma unroll
              for (int i = 0; i < k; ++i)
              {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < k; ++j)
                  acc += t_trafo[j * k + i] * um[j];
                tmp[i] = acc;
              }
#pragma unroll
              for (int i = 0; i < k; ++i)
                um[i] = tmp[i];
            }
            double* dst = sg + (size_t)cell[a] * nrt;
#pragma unroll
            for (int j = 0; j < k; ++j)
            {
              if (use_atomics)
              {
                atomicAdd(dst + fm * k + j, um[j]);
                atomicAdd(dst + fp * k + j, up[j]);
              }
              else
              {
                dst[fm * k + j] += um[j];
                dst[fp * k + j] += up[j];
              }
            }
#pragma unroll
            for (int i = 0; i < nadd; ++i)
            {
              const double val = xcol[nf * (k - 1) + 1 + a * nadd + i];
              if (use_atomics)
                atomicAdd(dst + 3 * k + ndiv + i, val);
              else
                dst[3 * k + ndiv + i] += val;
            }
          }
        }
      }
    }
  }
}

// Instantiations by the largest patch of the mesh: 8 / 16 cells, and 32 cells for degrees <= 3 (the
// per-thread frame of the dense patch system grows with NCMAX^2: 18 KB at k=2, 107 KB at k=3; degree 4
// would need 340 KB per thread and stays at 16).
using patch_kernel_t = void (*)(PatchView, int, int, TableView, const double*, const int32_t*, int, RhsPtrs, const double*, size_t,
                                int, const int32_t*, int, int);
template <int K, int NDG, bool EV, bool STRESS>
patch_kernel_t select_kernel(int ncmax)
{
  if (ncmax <= 8)
    return patch_kernel<K, NDG, 8, EV, STRESS>;
  if (ncmax <= 16)
    return patch_kernel<K, NDG, 16, EV, STRESS>;
  if constexpr (K <= EQLB_KMAX_WIDE)
    return patch_kernel<K, NDG, EQLB_NCMAX, EV, STRESS>;
  throw EqlbError(EQLB_ERR_INPUT, "eqlb_b200: patches with more than 16 cells are supported up to flux degree "
                                      + std::to_string(EQLB_KMAX_WIDE) + " (largest patch: " + std::to_string(ncmax) + " cells)");
}

template <int K, int NDG, bool EV>
void launch_patch_t(eqlb_handle* h, const double* const* dG, const double* const* dF, double* const* dSigma)
{
  RhsPtrs ptrs;
  for (int r = 0; r < h->nrhs; ++r)
  {
    ptrs.G[r] = dG[r];
    ptrs.F[r] = dF[r];
    ptrs.S[r] = dSigma[r];
  }
  const PatchView pv = h->patch_view();
  const int bs = 128;
  const size_t smem = (size_t)h->tv.ndoubles * sizeof(double);
  const bool atomics = (h->flags & EQLB_FLAG_ATOMIC) != 0;
  const bool stress = !EV && (h->flags & EQLB_FLAG_STRESS) != 0 && K >= 2;
  constexpr bool CAN_STRESS = !EV && K >= 2 && K <= 4;
  auto kern = select_kernel<K, NDG, EV, false>(h->ncmax);
  if (stress)
  {
    if constexpr (CAN_STRESS)
      kern = select_kernel<K, NDG, EV, true>(h->ncmax);
    else
      throw EqlbError(EQLB_ERR_INPUT, "stress equilibration: flux degree not supported by the CUDA kernel");
  }
  CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int32_t* dgmap = h->dg_identity ? nullptr : h->d_dg_dofmap.p;
  const size_t bstride = (size_t)h->ncell * h->nrt;
  if constexpr (CAN_STRESS)
  {
    // grouped boundary patches first, group by group in the reference's order
    // (se/reconstruction.hpp:170-234): equilibrate all members without weak symmetry
    // (they overlap -> atomics), then weak symmetry on the first member from the
    // accumulated global stress
    if (stress)
      for (size_t g = 0; g + 1 < h->h_group_off.size(); ++g)
      {
        const int first = h->h_group_off[g], count = h->h_group_off[g + 1] - first;
        auto kplain = select_kernel<K, NDG, EV, false>(h->ncmax);
        CUDA_CHECK(cudaFuncSetAttribute(kplain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kplain<<<1, bs, smem, h->stream>>>(pv, first, count, h->tv, h->d_cellJ.p, dgmap, h->nrhs, ptrs, h->d_bflux.p, bstride,
                                           1, h->d_cell_fct.p, h->nfct, 0);
        CUDA_CHECK(cudaGetLastError());
        kern<<<1, bs, smem, h->stream>>>(pv, first, 1, h->tv, h->d_cellJ.p, dgmap, h->nrhs, ptrs, h->d_bflux.p, bstride, 1,
                                         h->d_cell_fct.p, h->nfct, 1);
        CUDA_CHECK(cudaGetLastError());
        h->launches += 2;
      }
  }
  const int ngrouped = h->h_group_off.empty() ? 0 : h->h_group_off.back();
  if (atomics)
  {
    const int count = h->nactive - ngrouped;
    if (count == 0)
      return;
    kern<<<(count + bs - 1) / bs, bs, smem, h->stream>>>(pv, ngrouped, count, h->tv, h->d_cellJ.p, dgmap, h->nrhs, ptrs,
                                                        h->d_bflux.p, bstride, 1, h->d_cell_fct.p, h->nfct, 0);
    CUDA_CHECK(cudaGetLastError());
    h->launches++;
  }
  else
  {
    bool launched_before = false;  // a kernel of this call precedes the next launch in the stream
    for (int c = std::max(0, h->win_lo); c < std::min(h->nseg, h->win_hi); ++c)
    {
      int first = h->h_colour_off[c];
      int count = h->h_colour_off[c + 1] - first;
      static const bool stress_fast = getenv("EQLB_STRESS_GENERIC") == nullptr;
      const bool k2ok = (K == 2 && NDG == 3 && h->d_k2tab.p);
      const bool k1ok = (K == 1 && NDG == 1 && h->d_k1tab.p);
      if (!(h->flags & EQLB_FLAG_GENERIC) && h->dg_identity && !h->h_seg_subs[c].empty()  // (the lane-per-cell kernels read G, f in the DOLFINx layout)
          && (stress ? (k2ok && stress_fast && !EV) : (k1ok || k2ok || (kw_supported(K, NDG) && h->d_kwtab.p))))
      {
        // specialised kernels for the eligible head of the segment (one launch per lane class), generic
        // kernel for the rest
        static const bool force_kw = getenv("EQLB_KW") != nullptr;
        // Within one colour every DOF is touched by exactly one patch, so the accumulation can
        // use fire-and-forget RED.ADD.F64 and stay bitwise deterministic (the colours are
        // serialised on the stream); measured 14-25 % faster than load-add-store.
        static const int use_red = getenv("EQLB_RED") ? atoi(getenv("EQLB_RED")) : 1;
        // programmatic dependent launch behind a kernel of this very call (never behind foreign work
        // that may still be producing G or f)
        static const bool pdl_on = getenv("EQLB_PDL") ? atoi(getenv("EQLB_PDL")) != 0 : true;
        for (const auto& sub : h->h_seg_subs[c])
        {
          const bool pdl = pdl_on && launched_before && use_red;
          if (K == 1)
            launch_k1(h, EV, ptrs, sub.first, sub.count, use_red, sub.lanes, sub.recoff, pdl);
          else if (K == 2 && (!force_kw || stress))
            launch_k2(h, EV, ptrs, sub.first, sub.count, use_red, sub.lanes, sub.lanes, sub.recoff, stress, pdl);
          else
            launch_kw(h, EV, ptrs, sub.first, sub.count, use_red, sub.lanes, sub.recoff);
          launched_before = true;
        }
        const int nfast = h->h_colour_fast[c];
        first += nfast;
        count -= nfast;
      }
      if (count == 0)
        continue;
      kern<<<(count + bs - 1) / bs, bs, smem, h->stream>>>(pv, first, count, h->tv, h->d_cellJ.p, dgmap, h->nrhs, ptrs,
                                                          h->d_bflux.p, bstride, 0, h->d_cell_fct.p, h->nfct, 0);
      CUDA_CHECK(cudaGetLastError());
      h->launches++;
      launched_before = true;
    }
  }
}

// This file is compiled in EQLB_SE_PARTS translation units (-DEQLB_SE_PART=n, see the Makefile): each part
// instantiates the kernels of some (degree_flux, degree_dg) combinations, so that the parts build in parallel.
#ifndef EQLB_SE_PART
#error "se_kernel.cu: compile with -DEQLB_SE_PART=1..4"
#endif
#define SE_CASE(part, KK, ND)                                                                                          \
  case KK * 100 + ND:                                                                                                  \
    if constexpr (EQLB_SE_PART == part)                                                                                \
    {                                                                                                                  \
      launch_patch_t<KK, ND, EV>(h, dG, dF, dSigma);                                                                   \
      return true;                                                                                                     \
    }                                                                                                                  \
    return false;

template <bool EV>
bool dispatch(eqlb_handle* h, const double* const* dG, const double* const* dF, double* const* dSigma)
{
  switch (h->k * 100 + h->ndg)
  {
    SE_CASE(1, 1, 1)
    SE_CASE(1, 2, 1)
    SE_CASE(1, 2, 3)
    SE_CASE(2, 3, 1)
    SE_CASE(2, 3, 3)
    SE_CASE(2, 3, 6)
    SE_CASE(3, 4, 1)
    SE_CASE(3, 4, 3)
    SE_CASE(4, 4, 6)
    SE_CASE(4, 4, 10)
  default:
    return false;
  }
}

} // namespace

#define SE_CAT2(a, b) a##b
#define SE_CAT(a, b) SE_CAT2(a, b)
bool SE_CAT(launch_generic_part, EQLB_SE_PART)(eqlb_handle* h, bool ev, const double* const* dG, const double* const* dF,
                                                double* const* dSigma)
{
  return ev ? dispatch<true>(h, dG, dF, dSigma) : dispatch<false>(h, dG, dF, dSigma);
}

#if EQLB_SE_PART == 1
bool launch_generic_part2(eqlb_handle*, bool, const double* const*, const double* const*, double* const*);
bool launch_generic_part3(eqlb_handle*, bool, const double* const*, const double* const*, double* const*);
bool launch_generic_part4(eqlb_handle*, bool, const double* const*, const double* const*, double* const*);

static void launch_any(eqlb_handle* h, bool ev, const double* const* dG, const double* const* dF, double* const* dSigma)
{
  if (!(launch_generic_part1(h, ev, dG, dF, dSigma) || launch_generic_part2(h, ev, dG, dF, dSigma)
        || launch_generic_part3(h, ev, dG, dF, dSigma) || launch_generic_part4(h, ev, dG, dF, dSigma)))
    throw EqlbError(EQLB_ERR_INPUT, "patch kernel: unsupported (degree_flux, degree_dg) combination");
}

void launch_se(eqlb_handle* h, const double* const* dG, const double* const* dF, double* const* dSigma, double* dKorn)
{
  launch_any(h, false, dG, dF, dSigma);
  if (dKorn)
    launch_korn(h, dKorn);
}

void launch_ev(eqlb_handle* h, const double* const* dG, const double* const* dF, double* const* dSigma)
{
  launch_any(h, true, dG, dF, dSigma);
}
#endif
