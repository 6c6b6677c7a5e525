// Warp-cooperative patch kernel for general flux degree K (K = 2, 3; DG_{K-1} data):
// S lanes per patch, one lane per patch cell.  Generalises patch_k2w_kernel
// (patch_k2_kernel.cu) to facet blocks of size B = K-1:
//   * per cell (lane): Jacobian data, facet / cell moments, the explicit step-1 sweep as a
//     segmented scan, the (2K+nadd) x (2K+nadd) block of the RT mass matrix and the load
//     vector - all from exact reference matrices gathered per local facet pair;
//   * the additional interior functions of a cell (nadd = (K-1)(K-2)/2) are condensed
//     statically inside the lane;
//   * the patch system is block-tridiagonal along the fan (B x B blocks per facet) with a
//     (B+1) x (B+1) border (facet E_0 and the circulation dof d0); the elimination runs
//     along the lanes with warp shuffles, the border is reduced with butterfly sums.
// Same semantics as patch_kernel (se_kernel.cu) - reference
// se/solve_patch_semiexplt.hpp:212-1163 (SE) and the null-space form of
// ev/solve_patch.hpp:58-239 (EV); patches that may need `reversion_required` and the
// stress path stay on the generic kernel.
#include <type_traits>

#include "eqlb_internal.cuh"

namespace
{

template <int K>
struct KW
{
  static constexpr int B = K - 1;
  static constexpr int nadd = (K - 1) * (K - 2) / 2;
  static constexpr int ndiv = K * (K + 1) / 2 - 1;
  static constexpr int NT = 1 + ndiv;
  static constexpr int NDG = K * (K + 1) / 2;
  static constexpr int nrt = K * (K + 2);
  static constexpr int nact = 2 * K + nadd;
  static constexpr int ncol = nact + ndiv;
  static constexpr int nz = nact - 1;
  // gathered table block per local facet pair (doubles)
  static constexpr int O_MASS = 0;                          // [3][nact][ncol]
  static constexpr int O_H = O_MASS + 3 * nact * ncol;      // [NDG][nact][2]
  static constexpr int O_CMF = O_H + NDG * nact * 2;        // [NT][NDG]
  static constexpr int O_CMG = O_CMF + NT * NDG;            // [NT][NDG][2]
  static constexpr int O_FM = O_CMG + NT * NDG * 2;         // [2][K][NDG]
  static constexpr int O_BC = O_FM + 2 * K * NDG;           // [2][K][K]
  static constexpr int RAW = O_BC + 2 * K * K;
  static constexpr int BLOCK = ((RAW + 13) / 16) * 16 + 2;  // == 2 (mod 16): combos 4 banks apart
  static constexpr int O_DGM = 6 * BLOCK;                   // [NT][NDG]
  static constexpr int O_MONO = O_DGM + NT * NDG;           // [NT]
  static constexpr int O_R = O_MONO + NT;                   // [K][K] reversed-facet transform
  static constexpr int TAB = O_R + K * K;
};

__host__ __device__ __forceinline__ int kw_combo(int fm, int fp) { return fm * 2 + (fp > fm ? fp - 1 : fp); }

template <int S>
__device__ __forceinline__ double kw_seg_sum(double v)
{
#pragma unroll
  for (int o = S / 2; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o, S);
  return v;
}

// inverse of a small SPD matrix (Gauss-Jordan without pivoting)
template <int N>
__device__ __forceinline__ void inv_spd(const double (&A)[N][N], double (&Ai)[N][N])
{
  double M[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j)
    {
      M[i][j] = A[i][j];
      Ai[i][j] = (i == j) ? 1.0 : 0.0;
    }
#pragma unroll
  for (int c = 0; c < N; ++c)
  {
    const double ip = eqlb_rcp(M[c][c]);
#pragma unroll
    for (int j = 0; j < N; ++j)
    {
      M[c][j] *= ip;
      Ai[c][j] *= ip;
    }
#pragma unroll
    for (int r = 0; r < N; ++r)
      if (r != c)
      {
        const double f = M[r][c];
#pragma unroll
        for (int j = 0; j < N; ++j)
        {
          M[r][j] -= f * M[c][j];
          Ai[r][j] -= f * Ai[c][j];
        }
      }
  }
}

// MINB: resident CTAs per SM the register allocation is held to (degree 3: 2 -> 254 registers, 3 -> 168 with
// ~10 spilled doubles, 4 -> 128; EQLB_KW3_MINB selects, measurements in profiles/r2_kernel_experiments.md)
template <int K, bool EV, int S, int MINB>
__global__ void __launch_bounds__(128, MINB)
patch_kw_kernel(PatchView pv, int first, int count, const double* __restrict__ tab, const double* __restrict__ cellJ, int nrhs,
                RhsPtrs ptrs, const double* __restrict__ bflux, size_t bflux_stride, int use_atomics,
                const int4* __restrict__ rec, int nfct, int nwt)
{
  using D = KW<K>;
  constexpr int B = D::B, nadd = D::nadd, ndiv = D::ndiv, NT = D::NT, NDG = D::NDG, nrt = D::nrt, nact = D::nact,
                ncol = D::ncol, nz = D::nz;
  constexpr int NB = 2 * B + 1;  // facet/d0 functions of a cell: [lo (B)][d0][hi (B)]
  extern __shared__ double s_mem[];
  for (int i = threadIdx.x; i < D::TAB; i += blockDim.x)
    s_mem[i] = tab[i];
  __syncthreads();
  const double* s_dgm = s_mem + D::O_DGM;
  const double* s_mono = s_mem + D::O_MONO;
  const double* s_R = s_mem + D::O_R;

  constexpr unsigned FULL = 0xffffffffu;
  constexpr int PPW = 32 / S;
  const int lane = threadIdx.x & 31;
  const int j = lane % S;
  // persistent CTAs: grid-stride loop over warp tiles (32 lane records), tables staged once
  // per CTA, record of the next tile in flight
  const int wstride = gridDim.x * (blockDim.x >> 5);
  int wt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int4 rc_next = (wt < nwt) ? rec[(size_t)wt * 32 + lane] : make_int4(0, 0, 0, 0);
  for (; wt < nwt; wt += wstride)
  {
  const int4 rc = rc_next;
  if (wt + wstride < nwt)
    rc_next = rec[(size_t)(wt + wstride) * 32 + lane];
  const int p = wt * PPW + lane / S;
  const bool valid = p < count;
  const size_t ip = (size_t)first + (valid ? p : 0);
  const int nc = valid ? (rc.y >> 16) : 0;
  const bool active = j < nc;
  const int32_t c = rc.x;
  const int info = active ? (rc.y & 0xffff) : 0;
  const int v = info & 3, fm = (info >> 2) & 3, fp = active ? (info >> 4) & 3 : 1;
  const bool rev0 = (info & 64) != 0, rev1 = (info & 128) != 0;
  const bool first_c = (j == 0), last_c = (j == nc - 1);
  const double* blk = s_mem + kw_combo(fm, fp) * D::BLOCK;

  // geometry (RHS independent)
  double adj[4] = {0.0, 0.0, 0.0, 0.0}, gmm[3] = {0.0, 0.0, 0.0}, det = 1.0;
  if (active)
  {
    const double2 j0 = reinterpret_cast<const double2*>(cellJ)[2 * (size_t)c];
    const double2 j1 = reinterpret_cast<const double2*>(cellJ)[2 * (size_t)c + 1];
    det = j0.x * j1.y - j0.y * j1.x;
    const double iad = eqlb_rcp(fabs(det));
    adj[0] = j1.y;
    adj[1] = -j0.y;
    adj[2] = -j1.x;
    adj[3] = j0.x;
    gmm[0] = (j0.x * j0.x + j1.x * j1.x) * iad;
    gmm[1] = (j0.x * j0.y + j1.x * j1.y) * iad;
    gmm[2] = (j0.y * j0.y + j1.y * j1.y) * iad;
  }
  const double sgn = det > 0.0 ? 1.0 : -1.0;
  const double pm = active ? ((fm == 1) ? sgn : -sgn) : 0.0;
  const double pp = active ? ((fp == 1) ? sgn : -sgn) : 0.0;
  double pp_prev = __shfl_up_sync(FULL, pp, 1, S);
  const double pp_last = __shfl_sync(FULL, pp, max(nc - 1, 0), S);
  double n_pm = __shfl_down_sync(FULL, pm, 1, S);
  const double f_pm = __shfl_sync(FULL, pm, 0, S);

  for (int r = 0; r < nrhs; ++r)
  {
    const double* __restrict__ G = ptrs.G[r];
    const double* __restrict__ Fv = ptrs.F[r];
    double* __restrict__ sig = ptrs.S[r];
    const uint8_t ri = valid ? pv.rhsinfo[(size_t)r * pv.stride + ip] : 0;
    const int ptype = ri & 3;
    const bool bc_e0 = (ri & 8) != 0, bc_en = (ri & 16) != 0;
    const bool internal = (ptype == EQLB_PATCH_INTERNAL);
    const bool req_bc = (ptype == EQLB_PATCH_ESSNT_DUAL || ptype == EQLB_PATCH_MIXED);
    const bool mark_z = req_bc, mark_f0 = req_bc, mark_fn = (ptype == EQLB_PATCH_ESSNT_DUAL);
    const int nch = internal ? nc - 1 : nc;
    const double ppv = first_c ? (internal ? pp_last : 0.0) : pp_prev;
    const double npm = last_c ? f_pm : n_pm;

    // ---- data of the cell: moments ----
    double cm[NT], mm[K], mp[K], Gc[EV ? 2 * NDG : 1];
#pragma unroll
    for (int t = 0; t < NT; ++t)
      cm[t] = 0.0;
#pragma unroll
    for (int q = 0; q < K; ++q)
      mm[q] = mp[q] = 0.0;
    if (active)
    {
      const double* cmf = blk + D::O_CMF;
      const double* cmg = blk + D::O_CMG;
      const double* fmo = blk + D::O_FM;
      const double ghx = (v == 0) ? -1.0 : (v == 1 ? 1.0 : 0.0);
      const double ghy = (v == 0) ? -1.0 : (v == 2 ? 1.0 : 0.0);
      const double nmx = ((fm == 2) ? 0.0 : -1.0) * adj[0] + ((fm == 0) ? -1.0 : (fm == 2 ? 1.0 : 0.0)) * adj[2];
      const double nmy = ((fm == 2) ? 0.0 : -1.0) * adj[1] + ((fm == 0) ? -1.0 : (fm == 2 ? 1.0 : 0.0)) * adj[3];
      const double npx = ((fp == 2) ? 0.0 : -1.0) * adj[0] + ((fp == 0) ? -1.0 : (fp == 2 ? 1.0 : 0.0)) * adj[2];
      const double npy = ((fp == 2) ? 0.0 : -1.0) * adj[1] + ((fp == 0) ? -1.0 : (fp == 2 ? 1.0 : 0.0)) * adj[3];
      const double2* gp = reinterpret_cast<const double2*>(G) + (size_t)NDG * c;
      const double* fpn = Fv + (size_t)NDG * c;
#pragma unroll
      for (int i = 0; i < NDG; ++i)
      {
        const double2 g = gp[i];
        const double a0 = adj[0] * g.x + adj[1] * g.y;
        const double a1 = adj[2] * g.x + adj[3] * g.y;
        const double fd = det * fpn[i];
        if (EV)
        {
          Gc[2 * i] = g.x;
          Gc[2 * i + 1] = g.y;
          const double gg = a0 * ghx + a1 * ghy;
#pragma unroll
          for (int t = 0; t < NT; ++t)
            cm[t] += fd * cmf[t * NDG + i] + gg * s_dgm[t * NDG + i];
        }
        else
        {
          const double gnm = nmx * g.x + nmy * g.y;
          const double gnp = npx * g.x + npy * g.y;
#pragma unroll
          for (int q = 0; q < K; ++q)
          {
            mm[q] += fmo[q * NDG + i] * gnm;
            mp[q] += fmo[(K + q) * NDG + i] * gnp;
          }
#pragma unroll
          for (int t = 0; t < NT; ++t)
            cm[t] += fd * cmf[t * NDG + i] - a0 * cmg[(t * NDG + i) * 2] - a1 * cmg[(t * NDG + i) * 2 + 1];
        }
      }
    }

    const bool on_bnd = active && !internal && (first_c || last_c);
    bool has_bc = false;
    if (on_bnd)
    {
      if (ptype == EQLB_PATCH_ESSNT_DUAL)
        has_bc = true;
      else if (ptype == EQLB_PATCH_MIXED)
        has_bc = first_c ? bc_e0 : bc_en;
    }
    double bv[K];
#pragma unroll
    for (int q = 0; q < K; ++q)
      bv[q] = 0.0;
    if (has_bc)
    {
      const double* bsrc = bflux + (size_t)r * bflux_stride + (size_t)c * nrt + (first_c ? fm : fp) * K;
      const double* bc = blk + D::O_BC + (first_c ? 0 : K * K);
      double b[K];
      int nzero = 0;
#pragma unroll
      for (int i = 0; i < K; ++i)
      {
        b[i] = bsrc[i];
        if (fabs(b[i]) < 1e-7)
          ++nzero;
      }
      if (nzero < K)
      {
#pragma unroll
        for (int q = 0; q < K; ++q)
#pragma unroll
          for (int i = 0; i < K; ++i)
            bv[q] += bc[q * K + i] * b[i];
      }
    }

    // ---- EV: mean-value shift (ev/assembly.hpp:283-298) ----
    if (EV)
    {
      double tot = active ? sgn * cm[0] : 0.0;
      if (has_bc && ptype == EQLB_PATCH_ESSNT_DUAL)
        tot -= (first_c ? pm : pp) * bv[0];
      tot = kw_seg_sum<S>(tot);
      const double area2 = kw_seg_sum<S>(active ? fabs(det) : 0.0);
      if (valid && (internal || ptype == EQLB_PATCH_ESSNT_DUAL))
      {
        const double lam = tot * eqlb_rcp(0.5 * area2);
#pragma unroll
        for (int t = 0; t < NT; ++t)
          cm[t] -= lam * det * s_mono[t];
      }
    }

    // ---- step 1 (explicit sweep) ----
    double cfv[ncol];
#pragma unroll
    for (int q = 0; q < ncol; ++q)
      cfv[q] = 0.0;
#pragma unroll
    for (int t = 0; t < ndiv; ++t)
      cfv[nact + t] = cm[1 + t];
    double n_m[K];
    {
      const double mp0_up = __shfl_up_sync(FULL, mp[0], 1, S);
#pragma unroll
      for (int q = 0; q < K; ++q)
      {
        const double dn = __shfl_down_sync(FULL, mm[q], 1, S);
        const double fs = __shfl_sync(FULL, mm[q], 0, S);
        n_m[q] = last_c ? fs : dn;
      }
      double surf = 0.0;
      if (!EV && active)
      {
        if (!first_c)
          surf = -mm[0] - ppv * pm * mp0_up;
        else if (!internal && (has_bc || ptype == EQLB_PATCH_MIXED))
        {
          const double sg = (ptype == EQLB_PATCH_MIXED && !has_bc) ? 1.0 : -1.0;
#pragma unroll
          for (int q = 1; q < K; ++q)
            cfv[q] += sg * mm[q];
          if (has_bc)
            surf = -mm[0];
        }
      }
      const double vol = active ? sgn * cm[0] : 0.0;
      const double t_add = pm * surf + ((has_bc && first_c) ? pm * bv[0] : 0.0);
      double c_p = vol - t_add;
#pragma unroll
      for (int o = 1; o < S; o <<= 1)
      {
        const double up = __shfl_up_sync(FULL, c_p, o, S);
        if (j >= o)
          c_p += up;
      }
      const double c_m = vol - c_p;
      if (has_bc)
      {
#pragma unroll
        for (int q = 1; q < K; ++q)
          cfv[(first_c ? 0 : K) + q] += bv[q];
      }
      if (active)
      {
        if (!EV)
        {
          if (on_bnd && last_c)
          {
            const double pf = has_bc ? -1.0 : 1.0;
#pragma unroll
            for (int q = 1; q < K; ++q)
              cfv[K + q] += pf * mp[q];
          }
          else
          {
            const double tau = -pp * npm;
            const bool corr = rev1 && !last_c;
            const double j0 = tau * n_m[0] - mp[0];
#pragma unroll
            for (int q = 1; q < K; ++q)
            {
              double mt = n_m[q];
              if (rev1)
              {
                // moments of the neighbour's trace in this cell's facet parameter: s' = 1 - s
                mt = 0.0;
                double binom = 1.0;
#pragma unroll
                for (int i = 0; i <= q; ++i)
                {
                  mt += ((i & 1) ? -binom : binom) * n_m[i];
                  binom = binom * (double)(q - i) / (double)(i + 1);
                }
              }
              double h = tau * mt - mp[q];
              if (corr)
                h += -j0 + npm * c_p;
              cfv[K + q] += h;
            }
          }
        }
        else if (rev1 && !last_c)
        {
#pragma unroll
          for (int q = 1; q < K; ++q)
            cfv[K + q] += npm * c_p;
        }
        cfv[0] += pm * c_m;
        cfv[K] += pp * c_p;
      }
    }

    // ---- cell tensors: y = M_c cfv (- (hat G, phi)), MQ = square block of M_c ----
    double y[nact], MQ[nact][nact];
    {
      const double* tm = blk + D::O_MASS;
#pragma unroll
      for (int q = 0; q < nact; ++q)
      {
        double acc = 0.0;
#pragma unroll
        for (int s2 = 0; s2 < ncol; ++s2)
        {
          // the square block is symmetric (the three reference matrices are): reuse the transposed entry
          const double mv = (s2 < q && s2 < nact) ? MQ[s2][q]
                                                  : gmm[0] * tm[q * ncol + s2] + gmm[1] * tm[(nact + q) * ncol + s2]
                                                        + gmm[2] * tm[(2 * nact + q) * ncol + s2];
          acc += mv * cfv[s2];
          if (s2 < nact)
            MQ[q][s2] = mv;
        }
        y[q] = acc;
      }
    }
    if (EV)
    {
      const double* hh = blk + D::O_H;
#pragma unroll
      for (int mI = 0; mI < NDG; ++mI)
      {
        const double gx = Gc[2 * mI], gy = Gc[2 * mI + 1];
        const double jg0 = sgn * (adj[3] * gx - adj[2] * gy);
        const double jg1 = sgn * (-adj[1] * gx + adj[0] * gy);
#pragma unroll
        for (int q = 0; q < nact; ++q)
          y[q] -= jg0 * hh[(mI * nact + q) * 2] + jg1 * hh[(mI * nact + q) * 2 + 1];
      }
    }
    if (rev0)
    {
      // reversed E_{a-1}: functions of that facet are transformed with R (rows and columns)
      double tmp[K];
#pragma unroll
      for (int s2 = 0; s2 < nact; ++s2)
      {
#pragma unroll
        for (int i = 0; i < K; ++i)
        {
          double acc = 0.0;
#pragma unroll
          for (int q = 0; q < K; ++q)
            acc += s_R[i * K + q] * MQ[q][s2];
          tmp[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < K; ++i)
          MQ[i][s2] = tmp[i];
      }
#pragma unroll
      for (int q = 0; q < nact; ++q)
      {
#pragma unroll
        for (int i = 0; i < K; ++i)
        {
          double acc = 0.0;
#pragma unroll
          for (int q2 = 0; q2 < K; ++q2)
            acc += s_R[i * K + q2] * MQ[q][q2];
          tmp[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < K; ++i)
          MQ[q][i] = tmp[i];
      }
#pragma unroll
      for (int i = 0; i < K; ++i)
      {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < K; ++q)
          acc += s_R[i * K + q] * y[q];
        tmp[i] = acc;
      }
#pragma unroll
      for (int i = 0; i < K; ++i)
        y[i] = tmp[i];
    }
    const double p_ea = -pp;
    const double p_em = rev0 ? -ppv : pm;
    // H(div=0) functions of the cell, index i: [lo: i < B][d0: i == B][hi: B < i <= 2B][add: i > 2B]
    // function i != B is  s_i * row(i+1), d0 = p_em row(0) + p_ea row(K)
    double Te[nz][nz], lv[nz];
#pragma unroll
    for (int i = 0; i < nz; ++i)
    {
      const double si = (i < B) ? p_em : ((i <= 2 * B) ? p_ea : 1.0);
      lv[i] = (i == B) ? -(p_em * y[0] + p_ea * y[K]) : -si * y[i + 1];
#pragma unroll
      for (int i2 = 0; i2 < nz; ++i2)
      {
        const double sj = (i2 < B) ? p_em : ((i2 <= 2 * B) ? p_ea : 1.0);
        double val;
        if (i == B && i2 == B)
          val = MQ[0][0] + MQ[K][K] + 2.0 * p_em * p_ea * MQ[0][K];
        else if (i == B)
          val = sj * (p_em * MQ[0][i2 + 1] + p_ea * MQ[K][i2 + 1]);
        else if (i2 == B)
          val = si * (p_em * MQ[i + 1][0] + p_ea * MQ[i + 1][K]);
        else
          val = si * sj * MQ[i + 1][i2 + 1];
        Te[i][i2] = val;
      }
    }
    // static condensation of the additional cell functions (indices NB .. nz-1)
    double Aai[nadd > 0 ? nadd : 1][nadd > 0 ? nadd : 1], Arow[nadd > 0 ? nadd : 1][NB], la[nadd > 0 ? nadd : 1];
    if constexpr (nadd > 0)
    {
      double Aaa[nadd][nadd];
#pragma unroll
      for (int a = 0; a < nadd; ++a)
      {
        la[a] = lv[NB + a];
#pragma unroll
        for (int b2 = 0; b2 < nadd; ++b2)
          Aaa[a][b2] = Te[NB + a][NB + b2];
#pragma unroll
        for (int i = 0; i < NB; ++i)
          Arow[a][i] = Te[NB + a][i];
      }
      if (!active)
      {
#pragma unroll
        for (int a = 0; a < nadd; ++a)
          Aaa[a][a] = 1.0;
      }
      inv_spd<nadd>(Aaa, Aai);
#pragma unroll
      for (int i = 0; i < NB; ++i)
      {
        // z = Aaa^-1 Arow[:, i]
        double z[nadd];
#pragma unroll
        for (int a = 0; a < nadd; ++a)
        {
          z[a] = 0.0;
#pragma unroll
          for (int b2 = 0; b2 < nadd; ++b2)
            z[a] += Aai[a][b2] * Arow[b2][i];
        }
#pragma unroll
        for (int i2 = 0; i2 < NB; ++i2)
        {
          double acc = 0.0;
#pragma unroll
          for (int a = 0; a < nadd; ++a)
            acc += Arow[a][i2] * z[a];
          Te[i2][i] -= acc;
        }
        double accl = 0.0;
#pragma unroll
        for (int a = 0; a < nadd; ++a)
          accl += z[a] * la[a];
        lv[i] -= accl;
      }
    }
    // constrained dofs / inactive lanes
    const bool hi_is_F = active && internal && last_c;
    const bool m_lo = first_c ? mark_f0 : false;
    const bool m_hi = (!internal && last_c) ? mark_fn : false;
#pragma unroll
    for (int i = 0; i < NB; ++i)
    {
      const bool mi = !active || (i < B ? m_lo : (i == B ? mark_z : m_hi));
#pragma unroll
      for (int i2 = 0; i2 < NB; ++i2)
      {
        const bool mj = !active || (i2 < B ? m_lo : (i2 == B ? mark_z : m_hi));
        if (mi || mj)
          Te[i][i2] = 0.0;
      }
      if (mi)
        lv[i] = 0.0;
    }

    // ---- patch system: lane b owns chain facet E_b (1 <= b <= nch); border = (E_0, d0) ----
    const bool hi_chain = active && !hi_is_F;
    const bool owns = (j >= 1 && j <= nch);
    double Dm[B][B], Em[B][B], Gm[B][B], wv[B], lb[B];
#pragma unroll
    for (int a = 0; a < B; ++a)
    {
      const double upl = __shfl_up_sync(FULL, hi_chain ? lv[B + 1 + a] : 0.0, 1, S);
      const double upw = __shfl_up_sync(FULL, hi_chain ? Te[B + 1 + a][B] : 0.0, 1, S);
      wv[a] = owns ? Te[a][B] + upw : 0.0;
      lb[a] = owns ? lv[a] + upl : 0.0;
#pragma unroll
      for (int b2 = 0; b2 < B; ++b2)
      {
        const double uphh = __shfl_up_sync(FULL, hi_chain ? Te[B + 1 + a][B + 1 + b2] : 0.0, 1, S);
        // coupling (E_0 comp b2, E_1 comp a) assembled by cell 0: Te[lo b2][hi a]
        const double uplh = __shfl_up_sync(FULL, Te[b2][B + 1 + a], 1, S);
        Dm[a][b2] = owns ? Te[a][b2] + uphh : ((a == b2) ? 1.0 : 0.0);
        Em[a][b2] = (owns && hi_chain && j < nch) ? Te[a][B + 1 + b2] : 0.0;
        Gm[a][b2] = owns ? ((j == 1 ? uplh : 0.0) + (hi_is_F ? Te[a][B + 1 + b2] : 0.0)) : 0.0;
      }
    }
    if (owns && !internal && j == nc && mark_fn)
    {
#pragma unroll
      for (int a = 0; a < B; ++a)
      {
        wv[a] = lb[a] = 0.0;
#pragma unroll
        for (int b2 = 0; b2 < B; ++b2)
        {
          Dm[a][b2] = (a == b2) ? 1.0 : 0.0;
          Gm[a][b2] = 0.0;
        }
      }
    }
    // border parts of this lane: [F (B)][Z]
    double bS[B + 1][B + 1], bL[B + 1];
#pragma unroll
    for (int a = 0; a < B; ++a)
    {
      bL[a] = (first_c ? lv[a] : 0.0) + (hi_is_F ? lv[B + 1 + a] : 0.0);
      bS[a][B] = (first_c ? Te[a][B] : 0.0) + (hi_is_F ? Te[B + 1 + a][B] : 0.0);
      bS[B][a] = bS[a][B];
#pragma unroll
      for (int b2 = 0; b2 < B; ++b2)
        bS[a][b2] = (first_c ? Te[a][b2] : 0.0) + (hi_is_F ? Te[B + 1 + a][B + 1 + b2] : 0.0);
    }
    bS[B][B] = Te[B][B];
    bL[B] = lv[B];

    double Di[B][B];
#pragma unroll
    for (int a = 0; a < B; ++a)
#pragma unroll
      for (int b2 = 0; b2 < B; ++b2)
        Di[a][b2] = (a == b2) ? 1.0 : 0.0;
    double oD[B][B], oG[B][B], oW[B], oL[B];
#pragma unroll
    for (int a = 0; a < B; ++a)
    {
      oW[a] = oL[a] = 0.0;
#pragma unroll
      for (int b2 = 0; b2 < B; ++b2)
        oD[a][b2] = oG[a][b2] = 0.0;
    }
#pragma unroll
    for (int b = 1; b < S; ++b)
    {
      double rD[B][B], rG[B][B], rW[B], rL[B];
#pragma unroll
      for (int a = 0; a < B; ++a)
      {
        rW[a] = __shfl_up_sync(FULL, oW[a], 1, S);
        rL[a] = __shfl_up_sync(FULL, oL[a], 1, S);
#pragma unroll
        for (int b2 = 0; b2 < B; ++b2)
        {
          rG[a][b2] = __shfl_up_sync(FULL, oG[a][b2], 1, S);
          if (b2 >= a)
            rD[a][b2] = __shfl_up_sync(FULL, oD[a][b2], 1, S);
        }
      }
      if (j == b)
      {
#pragma unroll
        for (int a = 0; a < B; ++a)
        {
          wv[a] -= rW[a];
          lb[a] -= rL[a];
#pragma unroll
          for (int b2 = 0; b2 < B; ++b2)
          {
            Dm[a][b2] -= (b2 >= a) ? rD[a][b2] : rD[b2][a];
            Gm[a][b2] -= rG[a][b2];
          }
        }
        inv_spd<B>(Dm, Di);
        // X = E^T D^-1 (rows: next facet comps)
        double X[B][B];
#pragma unroll
        for (int a = 0; a < B; ++a)
#pragma unroll
          for (int b2 = 0; b2 < B; ++b2)
          {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < B; ++q)
              acc += Em[q][a] * Di[q][b2];
            X[a][b2] = acc;
          }
#pragma unroll
        for (int a = 0; a < B; ++a)
        {
          double aw = 0.0, al = 0.0;
#pragma unroll
          for (int q = 0; q < B; ++q)
          {
            aw += X[a][q] * wv[q];
            al += X[a][q] * lb[q];
          }
          oW[a] = aw;
          oL[a] = al;
#pragma unroll
          for (int b2 = 0; b2 < B; ++b2)
          {
            double ad = 0.0, ag = 0.0;
#pragma unroll
            for (int q = 0; q < B; ++q)
            {
              ad += X[a][q] * Em[q][b2];
              ag += X[a][q] * Gm[q][b2];
            }
            oD[a][b2] = ad;
            oG[a][b2] = ag;
          }
        }
        // border: subtract [G w]^T D^-1 [G w | l]
        double DG[B][B], Dw[B], Dl[B];
#pragma unroll
        for (int a = 0; a < B; ++a)
        {
          double aw = 0.0, al = 0.0;
#pragma unroll
          for (int q = 0; q < B; ++q)
          {
            aw += Di[a][q] * wv[q];
            al += Di[a][q] * lb[q];
          }
          Dw[a] = aw;
          Dl[a] = al;
#pragma unroll
          for (int b2 = 0; b2 < B; ++b2)
          {
            double ag = 0.0;
#pragma unroll
            for (int q = 0; q < B; ++q)
              ag += Di[a][q] * Gm[q][b2];
            DG[a][b2] = ag;
          }
        }
        double szz = 0.0, slz = 0.0;
#pragma unroll
        for (int q = 0; q < B; ++q)
        {
          szz += wv[q] * Dw[q];
          slz += wv[q] * Dl[q];
        }
        bS[B][B] -= szz;
        bL[B] -= slz;
#pragma unroll
        for (int a = 0; a < B; ++a)
        {
          double sfz = 0.0, slf = 0.0;
#pragma unroll
          for (int q = 0; q < B; ++q)
          {
            sfz += Gm[q][a] * Dw[q];
            slf += Gm[q][a] * Dl[q];
          }
          bS[a][B] -= sfz;
          bS[B][a] -= sfz;
          bL[a] -= slf;
#pragma unroll
          for (int b2 = 0; b2 < B; ++b2)
          {
            double sff = 0.0;
#pragma unroll
            for (int q = 0; q < B; ++q)
              sff += Gm[q][a] * DG[q][b2];
            bS[a][b2] -= sff;
          }
        }
      }
    }
    // border system (B+1) x (B+1)
    double Sb[B + 1][B + 1], ub[B + 1];
#pragma unroll
    for (int a = 0; a <= B; ++a)
    {
      ub[a] = kw_seg_sum<S>(bL[a]);
#pragma unroll
      for (int b2 = a; b2 <= B; ++b2)
      {
        Sb[a][b2] = kw_seg_sum<S>(bS[a][b2]);
        Sb[b2][a] = Sb[a][b2];
      }
    }
    if (mark_f0 || !valid)
    {
#pragma unroll
      for (int a = 0; a < B; ++a)
#pragma unroll
        for (int b2 = 0; b2 < B; ++b2)
          Sb[a][b2] = (a == b2) ? 1.0 : 0.0;
    }
    if (mark_z || !valid)
      Sb[B][B] = 1.0;
    {
      double Si[B + 1][B + 1], tmpu[B + 1];
      inv_spd<B + 1>(Sb, Si);
#pragma unroll
      for (int a = 0; a <= B; ++a)
      {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q <= B; ++q)
          acc += Si[a][q] * ub[q];
        tmpu[a] = acc;
      }
#pragma unroll
      for (int a = 0; a <= B; ++a)
        ub[a] = tmpu[a];
    }
    // back substitution along the chain
    double u[B];
#pragma unroll
    for (int a = 0; a < B; ++a)
      u[a] = 0.0;
#pragma unroll
    for (int b = S - 1; b >= 1; --b)
    {
      double un[B];
#pragma unroll
      for (int a = 0; a < B; ++a)
        un[a] = __shfl_down_sync(FULL, u[a], 1, S);
      if (j == b)
      {
        double rhs[B];
#pragma unroll
        for (int a = 0; a < B; ++a)
        {
          double acc = lb[a] - wv[a] * ub[B];
#pragma unroll
          for (int q = 0; q < B; ++q)
            acc -= Em[a][q] * un[q] + Gm[a][q] * ub[q];
          rhs[a] = acc;
        }
#pragma unroll
        for (int a = 0; a < B; ++a)
        {
          double acc = 0.0;
#pragma unroll
          for (int q = 0; q < B; ++q)
            acc += Di[a][q] * rhs[q];
          u[a] = acc;
        }
      }
    }
    double u_next[B];
#pragma unroll
    for (int a = 0; a < B; ++a)
      u_next[a] = __shfl_down_sync(FULL, u[a], 1, S);

    // ---- map back and accumulate ----
    if (active)
    {
      // patch unknowns of the cell functions [lo][d0][hi]
      double uc[NB];
#pragma unroll
      for (int a = 0; a < B; ++a)
      {
        uc[a] = first_c ? ub[a] : u[a];
        uc[B + 1 + a] = hi_is_F ? ub[a] : u_next[a];
      }
      uc[B] = ub[B];
      double um[K], up[K];
      um[0] = p_em * uc[B];
      up[0] = p_ea * uc[B];
#pragma unroll
      for (int q = 1; q < K; ++q)
      {
        um[q] = p_em * uc[q - 1];
        up[q] = p_ea * uc[B + q];
      }
      if (rev0)
      {
        double tmp[K];
#pragma unroll
        for (int i = 0; i < K; ++i)
        {
          double acc = 0.0;
#pragma unroll
          for (int q = 0; q < K; ++q)
            acc += s_R[q * K + i] * um[q];
          tmp[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < K; ++i)
          um[i] = tmp[i];
      }
      double clo[K], chi[K];
#pragma unroll
      for (int q = 0; q < K; ++q)
      {
        clo[q] = cfv[q] + um[q];
        chi[q] = cfv[K + q] + up[q];
      }
      double cadd[nadd > 0 ? nadd : 1];
      if constexpr (nadd > 0)
      {
#pragma unroll
        for (int a = 0; a < nadd; ++a)
        {
          double rhs = 0.0;
#pragma unroll
          for (int b2 = 0; b2 < nadd; ++b2)
          {
            double t2 = la[b2];
#pragma unroll
            for (int i = 0; i < NB; ++i)
              t2 -= Arow[b2][i] * uc[i];
            rhs += Aai[a][b2] * t2;
          }
          cadd[a] = cfv[2 * K + a] + rhs;
        }
      }
      auto accum = [&](double* d, double val)
      {
        if (use_atomics)
          atomicAdd(d, val);
        else
          *d += val;
      };
      if (EV)
      {
        for (int side = (first_c && !internal) ? 0 : 1; side < 2; ++side)
        {
          const bool refl = (info & (side ? 512 : 256)) != 0;
          double* d = sig + (size_t)(side ? rc.w : rc.z) * K;
#pragma unroll
          for (int i = 0; i < K; ++i)
          {
            double acc = side ? chi[i] : clo[i];
            if (refl)
            {
              acc = 0.0;
#pragma unroll
              for (int q = 0; q < K; ++q)
                acc += s_R[q * K + i] * (side ? chi[q] : clo[q]);
            }
            accum(d + i, acc);
          }
        }
        double* dstc = sig + (size_t)nfct * K + (size_t)c * (K * K - K);
#pragma unroll
        for (int t = 0; t < ndiv; ++t)
          accum(dstc + t, cfv[nact + t]);
        if constexpr (nadd > 0)
        {
#pragma unroll
          for (int a = 0; a < nadd; ++a)
            accum(dstc + ndiv + a, cadd[a]);
        }
      }
      else
      {
        double* d = sig + (size_t)c * nrt;
#pragma unroll
        for (int q = 0; q < K; ++q)
        {
          accum(d + fm * K + q, clo[q]);
          accum(d + fp * K + q, chi[q]);
        }
#pragma unroll
        for (int t = 0; t < ndiv; ++t)
          accum(d + 3 * K + t, cfv[nact + t]);
        if constexpr (nadd > 0)
        {
#pragma unroll
          for (int a = 0; a < nadd; ++a)
            accum(d + 3 * K + ndiv + a, cadd[a]);
        }
      }
    }
  }
  }
}

template <int K>
void build_kw_tables_t(eqlb_handle* h, const eqlb_tables* t, DevBuf<double>& dst)
{
  using D = KW<K>;
  std::vector<double> tab(D::TAB, 0.0);
  constexpr int nrt = D::nrt, ndg = D::NDG, nt = D::NT, nact = D::nact, ncol = D::ncol, ndiv = D::ndiv, nadd = D::nadd;
  for (int fm = 0; fm < 3; ++fm)
    for (int fp = 0; fp < 3; ++fp)
    {
      if (fm == fp)
        continue;
      const int v = 3 - fm - fp;
      double* blk = tab.data() + kw_combo(fm, fp) * D::BLOCK;
      auto rdof = [&](int q)
      {
        if (q < K)
          return fm * K + q;
        if (q < 2 * K)
          return fp * K + (q - K);
        if (q < 2 * K + nadd)
          return 3 * K + ndiv + (q - 2 * K);
        return 3 * K + (q - 2 * K - nadd);
      };
      for (int m = 0; m < 3; ++m)
        for (int q = 0; q < nact; ++q)
          for (int s = 0; s < ncol; ++s)
            blk[D::O_MASS + (m * nact + q) * ncol + s] = t->rt_mass[((size_t)m * nrt + rdof(q)) * nrt + rdof(s)];
      for (int m = 0; m < ndg; ++m)
        for (int q = 0; q < nact; ++q)
          for (int d = 0; d < 2; ++d)
            blk[D::O_H + (m * nact + q) * 2 + d] = t->hat_dg_rt[(((size_t)v * ndg + m) * nrt + rdof(q)) * 2 + d];
      for (int tt = 0; tt < nt; ++tt)
        for (int i = 0; i < ndg; ++i)
        {
          blk[D::O_CMF + tt * ndg + i] = t->cell_mom_f[((size_t)v * nt + tt) * ndg + i];
          for (int d = 0; d < 2; ++d)
            blk[D::O_CMG + (tt * ndg + i) * 2 + d] = t->cell_mom_g[(((size_t)v * nt + tt) * ndg + i) * 2 + d];
        }
      for (int side = 0; side < 2; ++side)
      {
        const int f = side ? fp : fm;
        for (int q = 0; q < K; ++q)
        {
          for (int i = 0; i < ndg; ++i)
            blk[D::O_FM + (side * K + q) * ndg + i] = t->fct_mom[(((size_t)f * 3 + v) * K + q) * ndg + i];
          for (int i = 0; i < K; ++i)
            blk[D::O_BC + side * K * K + q * K + i] = t->bc_mat[(((size_t)f * 3 + v) * K + q) * K + i];
        }
      }
    }
  for (int i = 0; i < nt * ndg; ++i)
    tab[D::O_DGM + i] = t->dg_mono[i];
  for (int i = 0; i < nt; ++i)
    tab[D::O_MONO + i] = t->mono_int[i];
  for (int i = 0; i < K * K; ++i)
    tab[D::O_R + i] = t->trafo[i];
  dst.upload(tab.data(), tab.size());
}

#ifndef EQLB_KW3_MINB_DEFAULT
#define EQLB_KW3_MINB_DEFAULT 2
#endif

template <int K, bool EV>
void launch_kw_t(eqlb_handle* h, const RhsPtrs& ptrs, int first, int count, int use_atomics, int S, int64_t recoff)
{
  using D = KW<K>;
  const int bs = 128;
  const size_t smem = (size_t)D::TAB * sizeof(double);
  if ((S != 4 && S != 8 && S != 16) || recoff < 0)
    throw EqlbError(EQLB_ERR_STATE, "warp-cooperative kernel: segment without lane records");
  const int nwt = (count + (32 / S) - 1) / (32 / S);  // warp tiles
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
  static const int minb3 = getenv("EQLB_KW3_MINB") ? atoi(getenv("EQLB_KW3_MINB")) : EQLB_KW3_MINB_DEFAULT;
  const int minb = (K == 2) ? 4 : std::min(4, std::max(2, minb3));
  const int grid = std::max(1, std::min((nwt + 3) / 4, nsm * minb));
  using kern_t = void (*)(PatchView, int, int, const double*, const double*, int, RhsPtrs, const double*, size_t, int, const int4*,
                          int, int);
  kern_t kern = nullptr;
  auto pick = [&](auto mb)
  {
    constexpr int MB = decltype(mb)::value;
    return (S == 4) ? patch_kw_kernel<K, EV, 4, MB> : (S == 8 ? patch_kw_kernel<K, EV, 8, MB> : patch_kw_kernel<K, EV, 16, MB>);
  };
  if constexpr (K == 2)
    kern = pick(std::integral_constant<int, 4>{});
  else
    kern = (minb == 2) ? pick(std::integral_constant<int, 2>{})
                       : (minb == 3 ? pick(std::integral_constant<int, 3>{}) : pick(std::integral_constant<int, 4>{}));
  CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, bs, smem, h->stream>>>(h->patch_view(), first, count, h->d_kwtab.p, h->d_cellJ.p, h->nrhs, ptrs, h->d_bflux.p,
                                      (size_t)h->ncell * h->nrt, use_atomics, h->d_prec.p + recoff, h->nfct, nwt);
  CUDA_CHECK(cudaGetLastError());
  h->launches++;
}

} // namespace

bool kw_supported(int k, int ndg) { return (k == 2 && ndg == 3) || (k == 3 && ndg == 6); }

void build_kw_tables(eqlb_handle* h, const eqlb_tables* t)
{
  if (t->k == 2)
    build_kw_tables_t<2>(h, t, h->d_kwtab);
  else if (t->k == 3)
    build_kw_tables_t<3>(h, t, h->d_kwtab);
}

void launch_kw(eqlb_handle* h, bool ev, const RhsPtrs& ptrs, int first, int count, int use_atomics, int lanes, int64_t recoff)
{
  if (count <= 0)
    return;
  if (h->k == 2)
  {
    if (ev)
      launch_kw_t<2, true>(h, ptrs, first, count, use_atomics, lanes, recoff);
    else
      launch_kw_t<2, false>(h, ptrs, first, count, use_atomics, lanes, recoff);
  }
  else
  {
    if (ev)
      launch_kw_t<3, true>(h, ptrs, first, count, use_atomics, lanes, recoff);
    else
      launch_kw_t<3, false>(h, ptrs, first, count, use_atomics, lanes, recoff);
  }
}
