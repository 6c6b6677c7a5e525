// Lane-per-cell patch kernel for flux degree 1 with DG_0 data (BASELINE.json configs[0]:
// the reference's own demo configuration).  Same mathematics as `patch_kernel` (se_kernel.cu;
// reference se/solve_patch_semiexplt.hpp:212-1163 for SE, the null-space form of
// ev/solve_patch.hpp:58-239 for EV) and the same organisation as the degree-2 kernel
// (patch_k2_kernel.cu): S lanes per patch, one lane per patch cell, coalesced lane records,
// persistent CTAs, RED accumulation.  For k = 1 the patch H(div=0) space is one-dimensional
// (the circulation d0), so step 2 is a scalar division (se/PatchData.hpp:587-590) and the
// explicit sweep is a segmented scan.
// Patches whose RHS may need the `reversion_required` correction (boundary patches of
// multi-RHS problems) stay on the generic kernel.
#include <algorithm>

#include "eqlb_internal.cuh"

namespace
{

// per-combo table block (doubles): mass [3][2][2], H [2][2], cmf, cmg [2], fmom [2 sides], bc [2 sides]
constexpr int K1_O_MASS = 0, K1_O_H = 12, K1_O_CMF = 16, K1_O_CMG = 17, K1_O_FM = 19, K1_O_BC = 21, K1_BLOCK = 34;
constexpr int K1_TAB = 6 * K1_BLOCK + 2;  // + dg_mono, mono_int

__host__ __device__ __forceinline__ int k1_combo(int fm, int fp) { return fm * 2 + (fp > fm ? fp - 1 : fp); }

template <int S>
__device__ __forceinline__ double k1_seg_sum(double v)
{
#pragma unroll
  for (int o = S / 2; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o, S);
  return v;
}

template <bool EV, int S>
__global__ void __launch_bounds__(128, 6)
patch_k1w_kernel(PatchView pv, int first, int count, const double* __restrict__ tab, const double* __restrict__ cellJ, int nrhs,
                 RhsPtrs ptrs, const double* __restrict__ bflux, size_t bflux_stride, int use_atomics,
                 const int4* __restrict__ rec, int nwt)
{
  __shared__ double s_tab[K1_TAB];
  for (int i = threadIdx.x; i < K1_TAB; i += blockDim.x)
    s_tab[i] = tab[i];
  __syncthreads();
  const double s_dgm = s_tab[6 * K1_BLOCK], s_mono = s_tab[6 * K1_BLOCK + 1];
  // programmatic dependent launch (see patch_k2_kernel.cu): the next colour starts while this one
  // drains and waits right before its first update of sigma
  asm volatile("griddepcontrol.launch_dependents;");
  bool dep_pending = true;

  constexpr unsigned FULL = 0xffffffffu;
  constexpr int PPW = 32 / S;
  constexpr int nrt = 3;
  const int lane = threadIdx.x & 31;
  const int j = lane % S;
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
    const bool rev0 = (info & 64) != 0;
    const bool first_c = (j == 0), last_c = (j == nc - 1);
    const double* blk = s_tab + k1_combo(fm, fp) * K1_BLOCK;

    // geometry of the cell (RHS independent)
    double adj[4] = {0.0, 0.0, 0.0, 0.0}, gm[3] = {0.0, 0.0, 0.0}, det = 1.0;
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
      gm[0] = (j0.x * j0.x + j1.x * j1.x) * iad;
      gm[1] = (j0.x * j0.y + j1.x * j1.y) * iad;
      gm[2] = (j0.y * j0.y + j1.y * j1.y) * iad;
    }
    const double sgn = det > 0.0 ? 1.0 : -1.0;
    const double pm = active ? ((fm == 1) ? sgn : -sgn) : 0.0;
    const double pp = active ? ((fp == 1) ? sgn : -sgn) : 0.0;
    double pp_prev = __shfl_up_sync(FULL, pp, 1, S);
    const double pp_last = __shfl_sync(FULL, pp, max(nc - 1, 0), S);
    const double area2 = k1_seg_sum<S>(active ? fabs(det) : 0.0);
    // cell block of the RT mass matrix on the two patch facets of the cell
    double M00, M01, M11;
    {
      const double* tm = blk + K1_O_MASS;
      M00 = gm[0] * tm[0] + gm[1] * tm[4] + gm[2] * tm[8];
      M01 = gm[0] * tm[1] + gm[1] * tm[5] + gm[2] * tm[9];
      M11 = gm[0] * tm[3] + gm[1] * tm[7] + gm[2] * tm[11];
    }

    for (int r = 0; r < nrhs; ++r)
    {
      const double* __restrict__ G = ptrs.G[r];
      const double* __restrict__ Fv = ptrs.F[r];
      double* __restrict__ sig = ptrs.S[r];
      const uint8_t ri = valid ? pv.rhsinfo[(size_t)r * pv.stride + ip] : 0;
      const int ptype = ri & 3;
      const bool bc_e0 = (ri & 8) != 0, bc_en = (ri & 16) != 0;
      const bool internal = (ptype == EQLB_PATCH_INTERNAL);
      const bool mark_z = (ptype == EQLB_PATCH_ESSNT_DUAL || ptype == EQLB_PATCH_MIXED);
      const double ppv = first_c ? (internal ? pp_last : 0.0) : pp_prev;

      // ---- data of the cell: moments ----
      double cm0 = 0.0, mm0 = 0.0, mp0 = 0.0, gx = 0.0, gy = 0.0;
      if (active)
      {
        const double2 g0 = reinterpret_cast<const double2*>(G)[(size_t)c];
        const double f0 = Fv[(size_t)c];
        gx = g0.x;
        gy = g0.y;
        const double a0 = adj[0] * gx + adj[1] * gy, a1 = adj[2] * gx + adj[3] * gy;
        const double fd = det * f0;
        if (EV)
        {
          const double ghx = (v == 0) ? -1.0 : (v == 1 ? 1.0 : 0.0);
          const double ghy = (v == 0) ? -1.0 : (v == 2 ? 1.0 : 0.0);
          cm0 = fd * blk[K1_O_CMF] + (a0 * ghx + a1 * ghy) * s_dgm;
        }
        else
        {
          const double nmx = ((fm == 2) ? 0.0 : -1.0) * adj[0] + ((fm == 0) ? -1.0 : (fm == 2 ? 1.0 : 0.0)) * adj[2];
          const double nmy = ((fm == 2) ? 0.0 : -1.0) * adj[1] + ((fm == 0) ? -1.0 : (fm == 2 ? 1.0 : 0.0)) * adj[3];
          const double npx = ((fp == 2) ? 0.0 : -1.0) * adj[0] + ((fp == 0) ? -1.0 : (fp == 2 ? 1.0 : 0.0)) * adj[2];
          const double npy = ((fp == 2) ? 0.0 : -1.0) * adj[1] + ((fp == 0) ? -1.0 : (fp == 2 ? 1.0 : 0.0)) * adj[3];
          mm0 = blk[K1_O_FM] * (nmx * gx + nmy * gy);
          mp0 = blk[K1_O_FM + 1] * (npx * gx + npy * gy);
          cm0 = fd * blk[K1_O_CMF] - a0 * blk[K1_O_CMG] - a1 * blk[K1_O_CMG + 1];
        }
      }

      // ---- patch boundary value on the first / last facet (base/BoundaryData.cpp:686-745) ----
      const bool on_bnd = active && !internal && (first_c || last_c);
      bool has_bc = false;
      if (on_bnd)
      {
        if (ptype == EQLB_PATCH_ESSNT_DUAL)
          has_bc = true;
        else if (ptype == EQLB_PATCH_MIXED)
          has_bc = first_c ? bc_e0 : bc_en;
      }
      double bv0 = 0.0;
      if (has_bc)
      {
        const double b0 = bflux[(size_t)r * bflux_stride + (size_t)c * nrt + (first_c ? fm : fp)];
        if (!(fabs(b0) < 1e-7))
          bv0 = blk[K1_O_BC + (first_c ? 0 : 1)] * b0;
      }

      // ---- EV: mean-value shift (ev/assembly.hpp:283-298) ----
      if (EV)
      {
        double tot = active ? sgn * cm0 : 0.0;
        if (has_bc && ptype == EQLB_PATCH_ESSNT_DUAL)
          tot -= (first_c ? pm : pp) * bv0;
        tot = k1_seg_sum<S>(tot);
        if (valid && (internal || ptype == EQLB_PATCH_ESSNT_DUAL))
          cm0 -= tot * eqlb_rcp(0.5 * area2) * det * s_mono;
      }

      // ---- step 1 as a segmented scan ----
      const double mp0_up = __shfl_up_sync(FULL, mp0, 1, S);  // (shuffles stay outside of divergent code)
      const double mp0_prev = first_c ? 0.0 : mp0_up;
      double surf = 0.0;
      if (!EV && active)
      {
        if (!first_c)
          surf = -mm0 - ppv * pm * mp0_prev;
        else if (!internal && has_bc)
          surf = -mm0;
      }
      const double vol = active ? sgn * cm0 : 0.0;
      const double t_add = pm * surf + ((has_bc && first_c) ? pm * bv0 : 0.0);
      double c_p = vol - t_add;
#pragma unroll
      for (int o = 1; o < S; o <<= 1)
      {
        const double up = __shfl_up_sync(FULL, c_p, o, S);
        if (j >= o)
          c_p += up;
      }
      const double c_m = vol - c_p;
      const double c_lo = active ? pm * c_m : 0.0, c_hi = active ? pp * c_p : 0.0;

      // ---- step 2: one unknown (d0) ----
      const double p_ea = -pp;
      const double p_em = rev0 ? ppv : pm;  // (-pp_prev) * R, R = -1 on a reversed facet
      double y0 = M00 * c_lo + M01 * c_hi, y1 = M01 * c_lo + M11 * c_hi;
      if (EV)
      {
        const double jg0 = sgn * (adj[3] * gx - adj[2] * gy);
        const double jg1 = sgn * (-adj[1] * gx + adj[0] * gy);
        y0 -= jg0 * blk[K1_O_H] + jg1 * blk[K1_O_H + 1];
        y1 -= jg0 * blk[K1_O_H + 2] + jg1 * blk[K1_O_H + 3];
      }
      const double a_zz = k1_seg_sum<S>(active ? M00 + M11 + 2.0 * p_em * p_ea * M01 : 0.0);
      const double l_z = k1_seg_sum<S>(active ? -(p_em * y0 + p_ea * y1) : 0.0);
      const double u_z = (valid && !mark_z) ? l_z * eqlb_rcp(a_zz) : 0.0;

      // ---- accumulate ----
      if (dep_pending)
      {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        dep_pending = false;
      }
      if (active)
      {
        const double olo = c_lo + p_em * u_z, ohi = c_hi + p_ea * u_z;
        if (EV)
        {
          // conforming vector: T_a owns facet E_a (hi side), T_1 of a boundary patch also E_0;
          // reflected facets: c_g = R^T c_loc = -c_loc
          if (first_c && !internal)
          {
            const double val = (info & 256) ? -olo : olo;
            if (use_atomics)
              atomicAdd(sig + rc.z, val);
            else
              sig[rc.z] += val;
          }
          const double val = (info & 512) ? -ohi : ohi;
          if (use_atomics)
            atomicAdd(sig + rc.w, val);
          else
            sig[rc.w] += val;
        }
        else
        {
          double* d = sig + (size_t)c * nrt;
          if (use_atomics)
          {
            atomicAdd(d + fm, olo);
            atomicAdd(d + fp, ohi);
          }
          else
          {
            d[fm] += olo;
            d[fp] += ohi;
          }
        }
      }
    }
  }
}

template <bool EV>
void launch_k1_t(eqlb_handle* h, const RhsPtrs& ptrs, int first, int count, int use_atomics, int S, int64_t recoff, bool pdl)
{
  if ((S != 4 && S != 8 && S != 16) || recoff < 0)
    throw EqlbError(EQLB_ERR_STATE, "degree-1 kernel: segment without lane records");
  const int bs = 128;
  const int nwt = (count + (32 / S) - 1) / (32 / S);
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
  const int grid = std::max(1, std::min((nwt + 3) / 4, nsm * 6));
  auto kern = (S == 4) ? patch_k1w_kernel<EV, 4> : (S == 8 ? patch_k1w_kernel<EV, 8> : patch_k1w_kernel<EV, 16>);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(bs);
  cfg.stream = h->stream;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl ? 1 : 0;  // only behind one of our own launches of the same call (launch_patch_t)
  const PatchView pv = h->patch_view();
  const double *tabp = h->d_k1tab.p, *cj = h->d_cellJ.p, *bfl = h->d_bflux.p;
  const int4* recp = h->d_prec.p + recoff;
  const size_t bstride = (size_t)h->ncell * h->nrt;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, pv, first, count, tabp, cj, h->nrhs, ptrs, bfl, bstride, use_atomics, recp, nwt));
  CUDA_CHECK(cudaGetLastError());
  h->launches++;
}

} // namespace

// gather the reference tables per local facet pair (fm, fp); v = 3 - fm - fp
void build_k1_tables(eqlb_handle* h, const eqlb_tables* t)
{
  std::vector<double> tab(K1_TAB, 0.0);
  const int nrt = 3;
  for (int fm = 0; fm < 3; ++fm)
    for (int fp = 0; fp < 3; ++fp)
    {
      if (fm == fp)
        continue;
      const int v = 3 - fm - fp;
      double* blk = tab.data() + k1_combo(fm, fp) * K1_BLOCK;
      const int rd[2] = {fm, fp};
      for (int m = 0; m < 3; ++m)
        for (int q = 0; q < 2; ++q)
          for (int s = 0; s < 2; ++s)
            blk[K1_O_MASS + (m * 2 + q) * 2 + s] = t->rt_mass[((size_t)m * nrt + rd[q]) * nrt + rd[s]];
      for (int q = 0; q < 2; ++q)
        for (int d = 0; d < 2; ++d)
          blk[K1_O_H + q * 2 + d] = t->hat_dg_rt[(((size_t)v * 1 + 0) * nrt + rd[q]) * 2 + d];
      blk[K1_O_CMF] = t->cell_mom_f[v];
      blk[K1_O_CMG] = t->cell_mom_g[(size_t)v * 2];
      blk[K1_O_CMG + 1] = t->cell_mom_g[(size_t)v * 2 + 1];
      for (int side = 0; side < 2; ++side)
      {
        const int f = side ? fp : fm;
        blk[K1_O_FM + side] = t->fct_mom[(size_t)f * 3 + v];
        blk[K1_O_BC + side] = t->bc_mat[(size_t)f * 3 + v];
      }
    }
  tab[6 * K1_BLOCK] = t->dg_mono[0];
  tab[6 * K1_BLOCK + 1] = t->mono_int[0];
  h->d_k1tab.upload(tab.data(), tab.size());
}

void launch_k1(eqlb_handle* h, bool ev, const RhsPtrs& ptrs, int first, int count, int use_atomics, int lanes, int64_t recoff,
               bool pdl)
{
  if (count <= 0)
    return;
  if (ev)
    launch_k1_t<true>(h, ptrs, first, count, use_atomics, lanes, recoff, pdl);
  else
    launch_k1_t<false>(h, ptrs, first, count, use_atomics, lanes, recoff, pdl);
}
