// Lane-per-cell patch kernel for flux degree 2 with DG_1 data (the configuration of
// BASELINE.json configs 2, 4, 5).  Same mathematics as `patch_kernel` (se_kernel.cu:
// reference se/solve_patch_semiexplt.hpp:212-1163 for SE, the null-space form of
// ev/solve_patch.hpp:58-239 for EV), organised around what the ncu captures of round 1
// showed (profiles/README.md):
//   * for k = 2 the patch system is "arrow + (cyclic) tridiagonal": one circulation dof
//     d0 coupled to everything, one dof per patch facet coupled to its two neighbours.
//     With facet E_0 (=: F) and d0 (=: Z) as a 2x2 border the remaining chain E_1..E_n is
//     eliminated along the lanes;
//   * the reference tables are pre-gathered per local facet pair (fm, fp) so that their
//     reads are contiguous and warp-uniform on structured meshes.
// Patches whose RHS may need the `reversion_required` correction (boundary patches of
// multi-RHS problems), the weak-symmetry stage of boundary patches and patches with more
// than 16 facets stay on the generic kernel.
// (An earlier per-thread streaming variant - one thread per patch, chain state in shared
// memory, 1.9 ms/step - was superseded by the lane-per-cell kernel and removed.)
#include "eqlb_internal.cuh"

namespace
{

constexpr int K2_NT = 3;  // cell moments: (0,0), (0,1), (1,0)
// per-combo table block (doubles): mass [3][4][6], H [3][4][2], cmf [3][3], cmg [3][3][2],
// fmom [2 sides][2][3], bc [2 sides][2][2]
// (block stride 146: consecutive combos are 4 banks apart -> lanes working on different
//  local facet pairs read their tables without shared-memory bank conflicts, 16-B aligned)
constexpr int K2_O_MASS = 0, K2_O_H = 72, K2_O_CMF = 96, K2_O_CMG = 106, K2_O_FM = 124, K2_O_BC = 136, K2_BLOCK = 146;
constexpr int K2_TAB = 6 * K2_BLOCK + 12;  // + dg_mono [3][3] + mono_int [3]
// stress (weak symmetry) variant: per-combo block of int phi_q^d lambda_j on the reference cell,
// [q: lo0 lo1 hi0 hi1 div0 div1][d][j: patch node, outer node of E_a, outer node of E_{a-1}]
constexpr int K2P_BLOCK = 50;  // 36 used; stride = 2 mod 16 doubles (bank-conflict free across combos)
constexpr int K2_TAB_STRESS = K2_TAB + 6 * K2P_BLOCK;

struct K2Cell
{
  int32_t c;
  int info;
  double adj[4], g[3], det, pm, pp;
  double cm[K2_NT];
  double mm[2], mp[2];
  double G[6];
};

__host__ __device__ __forceinline__ int combo_of(int fm, int fp) { return fm * 2 + (fp > fm ? fp - 1 : fp); }

// raw per-cell input: Jacobian (J00 J01 | J10 J11), projected flux (3 x double2), projected RHS (3)
struct K2Raw
{
  double2 j0, j1, g0, g1, g2;
  double f0, f1, f2;
};

__device__ __forceinline__ K2Raw load_raw(int32_t c, const double* __restrict__ cellJ, const double* __restrict__ G,
                                          const double* __restrict__ Fv)
{
  K2Raw r;
  r.j0 = reinterpret_cast<const double2*>(cellJ)[2 * (size_t)c];
  r.j1 = reinterpret_cast<const double2*>(cellJ)[2 * (size_t)c + 1];
  const double2* gp = reinterpret_cast<const double2*>(G) + 3 * (size_t)c;
  r.g0 = gp[0];
  r.g1 = gp[1];
  r.g2 = gp[2];
  r.f0 = Fv[3 * (size_t)c];
  r.f1 = Fv[3 * (size_t)c + 1];
  r.f2 = Fv[3 * (size_t)c + 2];
  return r;
}

// ---- software pipeline of the cell data (single-RHS launches): while a warp works on tile t, the J / G / f
// records of tile t+1 travel global -> shared memory with cp.async (no registers held, no scoreboard stall
// at the point of use).  ncu on the un-pipelined kernel: 25 % of all warp samples were long-scoreboard
// stalls in `load_cell` / the record load (profiles/r2a_ev_k2w_source_regions.txt).
// Per warp and stage: 5 x double2[32] + 3 x double[32] = 3328 B, lane-contiguous (conflict free).
constexpr int K2_STAGE_DOUBLES = 5 * 64 + 3 * 32;

__device__ __forceinline__ void cp_async16(void* dst, const void* src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

__device__ __forceinline__ void stage_issue(double* st, int lane, int32_t c, const double* __restrict__ cellJ,
                                            const double* __restrict__ G, const double* __restrict__ Fv)
{
  double2* d2 = reinterpret_cast<double2*>(st);
  const double2* js = reinterpret_cast<const double2*>(cellJ) + 2 * (size_t)c;
  const double2* gs = reinterpret_cast<const double2*>(G) + 3 * (size_t)c;
  cp_async16(d2 + lane, js);
  cp_async16(d2 + 32 + lane, js + 1);
  cp_async16(d2 + 64 + lane, gs);
  cp_async16(d2 + 96 + lane, gs + 1);
  cp_async16(d2 + 128 + lane, gs + 2);
  double* d1 = st + 320;
  const double* fs = Fv + 3 * (size_t)c;
  cp_async8(d1 + lane, fs);
  cp_async8(d1 + 32 + lane, fs + 1);
  cp_async8(d1 + 64 + lane, fs + 2);
}

__device__ __forceinline__ K2Raw stage_read(const double* st, int lane)
{
  K2Raw r;
  const double2* d2 = reinterpret_cast<const double2*>(st);
  r.j0 = d2[lane];
  r.j1 = d2[32 + lane];
  r.g0 = d2[64 + lane];
  r.g1 = d2[96 + lane];
  r.g2 = d2[128 + lane];
  const double* d1 = st + 320;
  r.f0 = d1[lane];
  r.f1 = d1[32 + lane];
  r.f2 = d1[64 + lane];
  return r;
}

template <bool EV>
__device__ __forceinline__ void load_cell(K2Cell& C, int32_t c, int info, const K2Raw& raw,
                                          const double* __restrict__ s_blk, const double* __restrict__ s_dgm)
{
  C.c = c;
  C.info = info;
  const double2 j0 = raw.j0, j1 = raw.j1, g0 = raw.g0, g1 = raw.g1, g2 = raw.g2;
  const double f0 = raw.f0, f1 = raw.f1, f2 = raw.f2;
  const double J00 = j0.x, J01 = j0.y, J10 = j1.x, J11 = j1.y;
  const double det = J00 * J11 - J01 * J10;
  const double iad = eqlb_rcp(fabs(det));
  C.det = det;
  C.adj[0] = J11;
  C.adj[1] = -J01;
  C.adj[2] = -J10;
  C.adj[3] = J00;
  C.g[0] = (J00 * J00 + J10 * J10) * iad;
  C.g[1] = (J00 * J01 + J10 * J11) * iad;
  C.g[2] = (J01 * J01 + J11 * J11) * iad;
  const double sgn = det > 0.0 ? 1.0 : -1.0;
  const int v = info & 3, fm = (info >> 2) & 3, fp = (info >> 4) & 3;
  C.pm = (fm == 1) ? sgn : -sgn;
  C.pp = (fp == 1) ? sgn : -sgn;
  const double gx[3] = {g0.x, g1.x, g2.x}, gy[3] = {g0.y, g1.y, g2.y}, ff[3] = {f0, f1, f2};
  const double* blk = s_blk + combo_of(fm, fp) * K2_BLOCK;
  const double* cmf = blk + K2_O_CMF;
  C.cm[0] = C.cm[1] = C.cm[2] = 0.0;
  C.mm[0] = C.mm[1] = C.mp[0] = C.mp[1] = 0.0;
  if (EV)
  {
    const double ghx = (v == 0) ? -1.0 : (v == 1 ? 1.0 : 0.0);
    const double ghy = (v == 0) ? -1.0 : (v == 2 ? 1.0 : 0.0);
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
      C.G[2 * i] = gx[i];
      C.G[2 * i + 1] = gy[i];
      const double a0 = C.adj[0] * gx[i] + C.adj[1] * gy[i];
      const double a1 = C.adj[2] * gx[i] + C.adj[3] * gy[i];
      const double fd = det * ff[i];
      const double gg = a0 * ghx + a1 * ghy;
#pragma unroll
      for (int t = 0; t < K2_NT; ++t)
        C.cm[t] += fd * cmf[t * 3 + i] + gg * s_dgm[t * 3 + i];
    }
  }
  else
  {
    const double* cmg = blk + K2_O_CMG;
    const double* fmo = blk + K2_O_FM;  // [side][j][i]
    const double nmx = ((fm == 2) ? 0.0 : -1.0) * C.adj[0] + ((fm == 0) ? -1.0 : (fm == 2 ? 1.0 : 0.0)) * C.adj[2];
    const double nmy = ((fm == 2) ? 0.0 : -1.0) * C.adj[1] + ((fm == 0) ? -1.0 : (fm == 2 ? 1.0 : 0.0)) * C.adj[3];
    const double npx = ((fp == 2) ? 0.0 : -1.0) * C.adj[0] + ((fp == 0) ? -1.0 : (fp == 2 ? 1.0 : 0.0)) * C.adj[2];
    const double npy = ((fp == 2) ? 0.0 : -1.0) * C.adj[1] + ((fp == 0) ? -1.0 : (fp == 2 ? 1.0 : 0.0)) * C.adj[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
      const double a0 = C.adj[0] * gx[i] + C.adj[1] * gy[i];
      const double a1 = C.adj[2] * gx[i] + C.adj[3] * gy[i];
      const double fd = det * ff[i];
      const double gnm = nmx * gx[i] + nmy * gy[i];
      const double gnp = npx * gx[i] + npy * gy[i];
#pragma unroll
      for (int j = 0; j < 2; ++j)
      {
        C.mm[j] += fmo[j * 3 + i] * gnm;
        C.mp[j] += fmo[6 + j * 3 + i] * gnp;
      }
#pragma unroll
      for (int t = 0; t < K2_NT; ++t)
        C.cm[t] += fd * cmf[t * 3 + i] - a0 * cmg[(t * 3 + i) * 2] - a1 * cmg[(t * 3 + i) * 2 + 1];
    }
  }
}

// patch boundary values (hat-weighted RT interpolant of the boundary flux) on the first /
// last facet; zero-traction threshold of base/BoundaryData.cpp:718
__device__ __forceinline__ void patch_bc(const double* __restrict__ bsrc, const double* __restrict__ bc, double& bv0,
                                         double& bv1)
{
  const double b0 = bsrc[0], b1 = bsrc[1];
  bv0 = 0.0;
  bv1 = 0.0;
  if (!(fabs(b0) < 1e-7 && fabs(b1) < 1e-7))
  {
    bv0 = bc[0] * b0 + bc[1] * b1;
    bv1 = bc[2] * b0 + bc[3] * b1;
  }
}

// ---------------------------------------------------------------------------
// Warp-cooperative variant: S lanes per patch, ONE LANE PER PATCH CELL.
// ncu on the per-thread streaming kernel (profiles/README.md, round-1 second capture)
// showed it latency bound: every thread walks its 8 cells through dependent global loads
// (record -> J/G/f) with 8 warps/SM of occupancy (170 registers + 72 KB shared state per
// CTA), issue slots 17 % busy.  Here all cells of a patch are loaded and turned into cell
// tensors concurrently (32 independent load streams per warp), the explicit sweep becomes
// a segmented inclusive scan, the bordered tridiagonal elimination runs along the lanes
// with shuffles, and no per-thread state is left in shared or local memory.
// S = 4, 8 or 16 >= number of patch facets of every patch of the launch.
// ---------------------------------------------------------------------------
template <int S>
__device__ __forceinline__ double seg_sum(double v)
{
#pragma unroll
  for (int o = S / 2; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o, S);
  return v;
}

// shared-memory scratch of the weak-symmetry stage per patch (doubles): factor [4][S],
// X [2][S+1][S+1]; tile stride = S mod 16 so that the 32/S patches of a warp spread over
// the banks (ncu: 17.9 M bank conflicts per S=4 launch with a stride of 8 mod 16)
template <int S>
struct K2Stress
{
  static constexpr int XROW = S + 1;
  static constexpr int O_X = 4 * S;
  static constexpr int RAW = O_X + 2 * (S + 1) * XROW;
  static constexpr int TILE = RAW + ((S - RAW % 16) + 32) % 16;
};

template <bool EV, int S, int MINB, bool STRESS, bool PIPE = false>
__global__ void __launch_bounds__(128, MINB)
patch_k2w_kernel(PatchView pv, int first, int count, const double* __restrict__ k2tab, const double* __restrict__ cellJ,
                 int nrhs, RhsPtrs ptrs, const double* __restrict__ bflux, size_t bflux_stride, int use_atomics,
                 const int4* __restrict__ rec, int nfct, int nwt)
{
  extern __shared__ double s_mem[];
  double* s_blk = s_mem;
  double* s_dgm = s_mem + 6 * K2_BLOCK;
  double* s_mono = s_dgm + 9;
  for (int i = threadIdx.x; i < (STRESS ? K2_TAB_STRESS : K2_TAB); i += blockDim.x)
    s_mem[i] = k2tab[i];
  __syncthreads();
  // PIPE: two stages of cell data per warp behind the tables (16-byte aligned: K2_TAB is even)
  [[maybe_unused]] double* s_stage = s_mem + K2_TAB + (threadIdx.x >> 5) * 2 * K2_STAGE_DOUBLES;
  [[maybe_unused]] const double* s_p1 = s_mem + K2_TAB;
  [[maybe_unused]] double* s_tile = s_mem + K2_TAB_STRESS + ((threadIdx.x >> 5) * (32 / S) + (threadIdx.x & 31) / S) * K2Stress<S>::TILE;
  [[maybe_unused]] double cf0[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // stress row 0, kept until row 1 is done
  // EV: programmatic dependent launch - the next colour may start (tables, records, cell data, the
  // arithmetic of its first tile) while this one drains; it waits right before its first update of
  // sigma.  Measured 0.740 -> 0.722 ms/step; for the SE instantiation the same in-loop branch costs
  // more than it gains (0.763 -> 0.806 ms), so SE waits right here after the prologue instead.
  [[maybe_unused]] bool dep_pending = true;
  asm volatile("griddepcontrol.launch_dependents;");
  if constexpr (!EV)
    asm volatile("griddepcontrol.wait;" ::: "memory");  // SE: only launch latency and the table staging overlap

  constexpr unsigned FULL = 0xffffffffu;
  constexpr int PPW = 32 / S;
  constexpr int k = 2, nrt = 8;
  const int lane = threadIdx.x & 31;
  const int j = lane % S;  // cell index within the patch / owned chain facet
  // persistent CTAs: every warp walks over warp tiles (32 lane records = PPW patches) with a
  // grid stride, the tables are staged once per CTA and the record of the next tile is in
  // flight while the current one is processed
  const int wstride = gridDim.x * (blockDim.x >> 5);
  int wt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int4 rc_next = (wt < nwt) ? rec[(size_t)wt * 32 + lane] : make_int4(0, 0, 0, 0);
  [[maybe_unused]] int32_t c_next = 0;  // PIPE: cell of this lane in the tile after the current one
  [[maybe_unused]] int stage = 0;
  if constexpr (PIPE)
  {
    // prologue: cell data of the first tile in flight, cell index of the second tile loaded
    if (wt < nwt)
      stage_issue(s_stage, lane, rc_next.x, cellJ, ptrs.G[0], ptrs.F[0]);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (wt + wstride < nwt)
      c_next = rec[(size_t)(wt + wstride) * 32 + lane].x;
  }
  for (; wt < nwt; wt += wstride)
  {
  const int4 rc = rc_next;
  if constexpr (PIPE)
  {
    // tile t+1: issue the copies of its cell data (cell index fetched one tile earlier) and load its full
    // record; tile t+2: fetch the cell index
    if (wt + wstride < nwt)
    {
      stage_issue(s_stage + (stage ^ 1) * K2_STAGE_DOUBLES, lane, c_next, cellJ, ptrs.G[0], ptrs.F[0]);
      rc_next = rec[(size_t)(wt + wstride) * 32 + lane];
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (wt + 2 * wstride < nwt)
      c_next = rec[(size_t)(wt + 2 * wstride) * 32 + lane].x;
  }
  else
  {
    if (wt + wstride < nwt)
      rc_next = rec[(size_t)(wt + wstride) * 32 + lane];
  }
  const int p = wt * PPW + lane / S;
  const bool valid = p < count;
  const size_t ip = (size_t)first + (valid ? p : 0);
  const int nc = valid ? (rc.y >> 16) : 0;
  const bool active = j < nc;
  const int32_t c = rc.x;
  const int info = active ? (rc.y & 0xffff) : 0;
  const int fm = (info >> 2) & 3, fp = active ? (info >> 4) & 3 : 1;
  const bool rev0 = (info & 64) != 0, rev1 = (info & 128) != 0;
  const bool first_c = (j == 0), last_c = (j == nc - 1);
  const double* blk = s_blk + combo_of(fm, fp) * K2_BLOCK;

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

    K2Cell cur;
    K2Raw raw;
    if constexpr (PIPE)
    {
      asm volatile("cp.async.wait_group 1;" ::: "memory");  // everything but the copies of tile t+1 has landed
      raw = stage_read(s_stage + stage * K2_STAGE_DOUBLES, lane);
      stage ^= 1;
    }
    else if (active)
      raw = load_raw(c, cellJ, G, Fv);
    if (active)
      load_cell<EV>(cur, c, info, raw, s_blk, s_dgm);
    else
    {
      cur.det = 1.0;
      cur.pm = cur.pp = 0.0;
      cur.cm[0] = cur.cm[1] = cur.cm[2] = 0.0;
      cur.mm[0] = cur.mm[1] = cur.mp[0] = cur.mp[1] = 0.0;
      cur.g[0] = cur.g[1] = cur.g[2] = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        cur.adj[i] = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i)
        cur.G[i] = 0.0;
    }
    const double sgn = cur.det > 0.0 ? 1.0 : -1.0;

    const bool on_bnd = active && !internal && (first_c || last_c);
    bool has_bc = false;
    if (on_bnd)
    {
      if (ptype == EQLB_PATCH_ESSNT_DUAL)
        has_bc = true;
      else if (ptype == EQLB_PATCH_MIXED)
        has_bc = first_c ? bc_e0 : bc_en;
    }
    double bv0 = 0.0, bv1 = 0.0;
    if (has_bc)
      patch_bc(bflux + (size_t)r * bflux_stride + (size_t)c * nrt + (first_c ? fm : fp) * k, blk + K2_O_BC + (first_c ? 0 : 4),
               bv0, bv1);
    // a 2-cell... every boundary patch has nc >= 2, so first and last cell are different lanes

    // ---- EV: mean-value shift (ev/assembly.hpp:283-298) ----
    if (EV)
    {
      double tot = active ? sgn * cur.cm[0] : 0.0;
      if (has_bc && ptype == EQLB_PATCH_ESSNT_DUAL)
        tot -= (first_c ? cur.pm : cur.pp) * bv0;
      tot = seg_sum<S>(tot);
      const double area2 = seg_sum<S>(active ? fabs(cur.det) : 0.0);
      if (valid && (internal || ptype == EQLB_PATCH_ESSNT_DUAL))
      {
        const double lam = tot * eqlb_rcp(0.5 * area2);
#pragma unroll
        for (int t = 0; t < K2_NT; ++t)
          cur.cm[t] -= lam * cur.det * s_mono[t];
      }
    }

    // ---- neighbours along the fan ----
    double pp_prev = __shfl_up_sync(FULL, cur.pp, 1, S);
    double mp0_prev = __shfl_up_sync(FULL, cur.mp[0], 1, S);
    const double pp_last = __shfl_sync(FULL, cur.pp, max(nc - 1, 0), S);
    if (first_c)
    {
      pp_prev = internal ? pp_last : 0.0;
      mp0_prev = 0.0;
    }
    double n_pm = __shfl_down_sync(FULL, cur.pm, 1, S);
    double n_m0 = __shfl_down_sync(FULL, cur.mm[0], 1, S);
    double n_m1 = __shfl_down_sync(FULL, cur.mm[1], 1, S);
    const double f_pm = __shfl_sync(FULL, cur.pm, 0, S), f_m0 = __shfl_sync(FULL, cur.mm[0], 0, S),
                 f_m1 = __shfl_sync(FULL, cur.mm[1], 0, S);
    if (last_c)
    {
      n_pm = f_pm;
      n_m0 = f_m0;
      n_m1 = f_m1;
    }

    // ---- step 1 as a segmented scan: c+_a = sum_{b<=a} (vol_b - t_b), c-_a = vol_a - c+_a ----
    double cfv[6] = {0.0, 0.0, 0.0, 0.0, cur.cm[1], cur.cm[2]};
    double surf = 0.0;
    if (!EV && active)
    {
      if (!first_c)
        surf = -cur.mm[0] - pp_prev * cur.pm * mp0_prev;
      else if (!internal && (has_bc || ptype == EQLB_PATCH_MIXED))
      {
        cfv[1] += ((ptype == EQLB_PATCH_MIXED && !has_bc) ? 1.0 : -1.0) * cur.mm[1];
        if (has_bc)
          surf = -cur.mm[0];
      }
    }
    const double vol = active ? sgn * cur.cm[0] : 0.0;
    const double t_add = cur.pm * surf + ((has_bc && first_c) ? cur.pm * bv0 : 0.0);
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
      if (first_c)
        cfv[1] += bv1;
      else
        cfv[3] += bv1;
    }
    if (active)
    {
      if (!EV)
      {
        if (on_bnd && last_c)
          cfv[3] += (has_bc ? -1.0 : 1.0) * cur.mp[1];
        else
        {
          const double tau = -cur.pp * n_pm;
          const double mt1 = rev1 ? (n_m0 - n_m1) : n_m1;
          double h = tau * mt1 - cur.mp[1];
          if (rev1 && !last_c)
            h += -(tau * n_m0 - cur.mp[0]) + n_pm * c_p;
          cfv[3] += h;
        }
      }
      else if (rev1 && !last_c)
        cfv[3] += n_pm * c_p;
      cfv[0] += cur.pm * c_m;
      cfv[2] += cur.pp * c_p;
    }

    // ---- cell block of the RT mass matrix and load ----
    double MB[4][6];
    {
      const double2* tm = reinterpret_cast<const double2*>(blk + K2_O_MASS);
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2)
        {
          // the 4 x 4 part is symmetric (so are the three reference matrices): entries below the
          // diagonal are taken from the transposed position
          // (EV only: measured 0.717 -> 0.692 ms/step; the SE instantiation gets slower with it,
          //  0.758 -> 0.801 ms, its register allocation is at the edge of spilling)
          const int c0 = 2 * s2, c1 = 2 * s2 + 1;
          if (EV && c1 < q)
          {
            MB[q][c0] = MB[c0][q];
            MB[q][c1] = MB[c1][q];
            continue;
          }
          const double2 a0 = tm[q * 3 + s2], a1 = tm[12 + q * 3 + s2], a2 = tm[24 + q * 3 + s2];
          MB[q][c0] = (EV && c0 < q) ? MB[c0][q] : cur.g[0] * a0.x + cur.g[1] * a1.x + cur.g[2] * a2.x;
          MB[q][c1] = cur.g[0] * a0.y + cur.g[1] * a1.y + cur.g[2] * a2.y;
        }
    }
    double y[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
    {
      double s2 = 0.0;
#pragma unroll
      for (int c2 = 0; c2 < 6; ++c2)
        s2 += MB[q][c2] * cfv[c2];
      y[q] = s2;
    }
    if (EV)
    {
      const double2* hh = reinterpret_cast<const double2*>(blk + K2_O_H);
#pragma unroll
      for (int mI = 0; mI < 3; ++mI)
      {
        const double gx = cur.G[2 * mI], gy = cur.G[2 * mI + 1];
        const double jg0 = sgn * (cur.adj[3] * gx - cur.adj[2] * gy);
        const double jg1 = sgn * (-cur.adj[1] * gx + cur.adj[0] * gy);
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
          const double2 hv = hh[mI * 4 + q];
          y[q] -= jg0 * hv.x + jg1 * hv.y;
        }
      }
    }
    if (rev0)
    {
#pragma unroll
      for (int s2 = 0; s2 < 4; ++s2)
        MB[0][s2] = -MB[0][s2] - MB[1][s2];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        MB[q][0] = -MB[q][0] - MB[q][1];
      y[0] = -y[0] - y[1];
    }
    const double p_ea = -cur.pp;
    const double p_em = rev0 ? -pp_prev : cur.pm;
    const double pe = p_em * p_ea;
    const bool hi_is_F = active && internal && last_c;
    const bool m_lo = first_c ? mark_f0 : false;
    const bool m_hi = (!internal && last_c) ? mark_fn : false;
    double T00 = MB[1][1], T22 = MB[3][3], T02 = pe * MB[1][3];
    double T11 = MB[0][0] + MB[2][2] + 2.0 * pe * MB[0][2];
    double T01 = MB[1][0] + pe * MB[1][2], T12 = pe * MB[3][0] + MB[3][2];
    double l0 = -p_em * y[1], l1 = -(p_em * y[0] + p_ea * y[2]), l2 = -p_ea * y[3];
    if (m_lo)
      T00 = T01 = T02 = l0 = 0.0;
    if (m_hi)
      T22 = T12 = T02 = l2 = 0.0;
    if (mark_z)
      T11 = T01 = T12 = l1 = 0.0;
    if (!active)
      T00 = T22 = T02 = T11 = T01 = T12 = l0 = l1 = l2 = 0.0;

    // ---- patch system: lane b owns chain facet E_b (1 <= b <= nch); E_0 and d0 are the border ----
    const bool hi_chain = active && !hi_is_F;
    const double up22 = __shfl_up_sync(FULL, hi_chain ? T22 : 0.0, 1, S);
    const double up12 = __shfl_up_sync(FULL, hi_chain ? T12 : 0.0, 1, S);
    const double upl2 = __shfl_up_sync(FULL, hi_chain ? l2 : 0.0, 1, S);
    const double up02 = __shfl_up_sync(FULL, T02, 1, S);  // coupling (E_0, E_1) from cell 0
    const bool owns = (j >= 1 && j <= nch);
    double D = 1.0, e = 0.0, g = 0.0, w = 0.0, l = 0.0;
    if (owns)
    {
      D = T00 + up22;
      w = T01 + up12;
      l = l0 + upl2;
      g = (j == 1 ? up02 : 0.0) + (hi_is_F ? T02 : 0.0);
      e = (hi_chain && j < nch) ? T02 : 0.0;
      if (!internal && j == nc && mark_fn)
      {
        D = 1.0;
        g = w = l = 0.0;
      }
    }
    // border parts held by this lane
    double bFF = (first_c ? T00 : 0.0) + (hi_is_F ? T22 : 0.0);
    double bFZ = (first_c ? T01 : 0.0) + (hi_is_F ? T12 : 0.0);
    double bLF = (first_c ? l0 : 0.0) + (hi_is_F ? l2 : 0.0);
    double bZZ = T11, bLZ = l1;

    double ipv = 1.0;
    double oD = 0.0, oG = 0.0, oW = 0.0, oL = 0.0;
#pragma unroll
    for (int b = 1; b < S; ++b)
    {
      const double rD = __shfl_up_sync(FULL, oD, 1, S), rG = __shfl_up_sync(FULL, oG, 1, S),
                   rW = __shfl_up_sync(FULL, oW, 1, S), rL = __shfl_up_sync(FULL, oL, 1, S);
      if (j == b)
      {
        D -= rD;
        g -= rG;
        w -= rW;
        l -= rL;
        ipv = eqlb_rcp(D);
        const double ei = e * ipv, gi = g * ipv, wi = w * ipv;
        oD = ei * e;
        oG = ei * g;
        oW = ei * w;
        oL = ei * l;
        bFF -= gi * g;
        bFZ -= gi * w;
        bZZ -= wi * w;
        bLF -= gi * l;
        bLZ -= wi * l;
      }
    }
    double S_FF = seg_sum<S>(bFF), S_FZ = seg_sum<S>(bFZ), S_ZZ = seg_sum<S>(bZZ);
    const double l_F = seg_sum<S>(bLF), l_Z = seg_sum<S>(bLZ);
    if (mark_z)
      S_ZZ = 1.0;
    if (mark_f0)
      S_FF = 1.0;
    double u_F = 0.0, u_Z = 0.0, idet = 0.0;
    if (valid)
    {
      idet = eqlb_rcp(S_FF * S_ZZ - S_FZ * S_FZ);
      u_F = (l_F * S_ZZ - S_FZ * l_Z) * idet;
      u_Z = (S_FF * l_Z - S_FZ * l_F) * idet;
    }
    double u = 0.0;
#pragma unroll
    for (int b = S - 1; b >= 1; --b)
    {
      const double un = __shfl_down_sync(FULL, u, 1, S);
      if (j == b)
        u = (l - e * un - g * u_F - w * u_Z) * ipv;
    }
    const double u_next = __shfl_down_sync(FULL, u, 1, S);

    // ---- map back ----
    double co[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // corrector of the cell: [lo0 lo1 hi0 hi1 div0 div1]
    if (active)
    {
      const double u_lo = first_c ? u_F : u;
      const double u_hi = hi_is_F ? u_F : u_next;
      double um0 = p_em * u_Z, um1 = p_em * u_lo;
      if (rev0)
      {
        const double t0 = -um0, t1 = -um0 + um1;
        um0 = t0;
        um1 = t1;
      }
      co[0] = cfv[0] + um0;
      co[1] = cfv[1] + um1;
      co[2] = cfv[2] + p_ea * u_Z;
      co[3] = cfv[3] + p_ea * u_hi;
      co[4] = cfv[4];
      co[5] = cfv[5];
    }

    if constexpr (STRESS)
    {
      // rows 0/1 of the stress tensor are accumulated after the weak-symmetry correction
      if (r == 0)
      {
#pragma unroll
        for (int q = 0; q < 6; ++q)
          cf0[q] = co[q];
        continue;
      }
      if (r == 1)
      {
        // ---- weak symmetry (se/solve_patch_weaksym.hpp:59-233, se/stressmin_kernel.hpp:76-248) for
        // interior patches: min |tau_0|^2 + |tau_1|^2 over the divergence-free patch space s.t.
        // (as(sigma + tau), lambda_n) = 0 for the hat functions of all patch nodes.  A is the
        // matrix just factorised; X = A^-1 B is solved one constraint column per lane, the Schur
        // complement K = sum_k B_k^T X_k is reduced analytically by the mean-value multiplier
        // (the columns of B sum to zero on an interior patch) to an nc x nc SPD system. ----
        using KS = K2Stress<S>;
        double* Fs = s_tile;            // [4][S]: e*ipv, g*ipv, w*ipv, ipv of the chain rows
        double* Xs = s_tile + KS::O_X;  // [2][S+1][XROW]
        Fs[j] = e * ipv;
        Fs[S + j] = g * ipv;
        Fs[2 * S + j] = w * ipv;
        Fs[3 * S + j] = ipv;
        // moments int_T phi_q^d lambda_n of the cell
        const double* tp = s_p1 + combo_of(fm, fp) * K2P_BLOCK;
        const double J00 = sgn * cur.adj[3], J01 = -sgn * cur.adj[1], J10 = -sgn * cur.adj[2], J11 = sgn * cur.adj[0];
        double Px[4][3], Py[4][3], sj[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < 6; ++q)
#pragma unroll
          for (int n = 0; n < 3; ++n)
          {
            const double t0 = tp[(q * 2) * 3 + n], t1 = tp[(q * 2 + 1) * 3 + n];
            const double px = J00 * t0 + J01 * t1, py = J10 * t0 + J11 * t1;
            sj[n] += cf0[q] * py - co[q] * px;  // (sigma_01 - sigma_10, lambda_n)
            if (q < 4)
            {
              Px[q][n] = px;
              Py[q][n] = py;
            }
          }
        // patch functions of the cell: lo (higher order on E_a-1), Z (d0), hi (higher order on E_a);
        // B_0 = (psi_y, lambda), B_1 = -(psi_x, lambda)
        double blo[2][3], bz[2][3], bhi[2][3];
#pragma unroll
        for (int n = 0; n < 3; ++n)
        {
          double q0x = Px[0][n], q0y = Py[0][n];
          if (rev0)
          {
            q0x = -q0x - Px[1][n];
            q0y = -q0y - Py[1][n];
          }
          blo[0][n] = p_em * Py[1][n];
          blo[1][n] = -p_em * Px[1][n];
          bz[0][n] = p_em * q0y + p_ea * Py[2][n];
          bz[1][n] = -(p_em * q0x + p_ea * Px[2][n]);
          bhi[0][n] = p_ea * Py[3][n];
          bhi[1][n] = -p_ea * Px[3][n];
        }
        // lane c owns the outer node N_c of E_c (lo facet of cell c): node slot 2 of cell c and
        // slot 1 of cell c-1 (cyclic)
        const int ncs = max(nc, 1);
        const int prv = (j + ncs - 1) % ncs, nxt = (j + 1) % ncs;
        double vA[2], vB[2], vC[2], vZ[2];
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
        {
          vA[kk] = __shfl_sync(FULL, blo[kk][1], prv, S);
          vB[kk] = blo[kk][2] + __shfl_sync(FULL, bhi[kk][1], prv, S);
          vC[kk] = bhi[kk][2];
          vZ[kk] = bz[kk][2] + __shfl_sync(FULL, bz[kk][1], prv, S);
          if (!active)
            vA[kk] = vB[kk] = vC[kk] = vZ[kk] = 0.0;
        }
        const double adet = active ? fabs(cur.det) : 0.0;
        double Lc = -(sj[2] + __shfl_sync(FULL, sj[1], prv, S));
        double mN = (adet + __shfl_sync(FULL, adet, prv, S)) * (1.0 / 6.0);
        if (!active)
          Lc = mN = 0.0;
        const double Lc0 = -seg_sum<S>(sj[0]);
        const double area = 0.5 * seg_sum<S>(adet);
        const double Lsum = Lc0 + seg_sum<S>(Lc);  // (shuffles stay outside of divergent code)
        const double mu = valid ? Lsum / area : 0.0;
        double rhs = -(Lc - mu * mN);
        __syncwarp();
        // ---- X[:, N_c] = A^-1 B[:, N_c] for both rows, column c in lane c ----
        double x[2][S + 1], xZ[2];
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
        {
#pragma unroll
          for (int b = 0; b < S; ++b)
            x[kk][b] = (b == prv ? vA[kk] : 0.0) + (b == j ? vB[kk] : 0.0) + (b == nxt ? vC[kk] : 0.0);
          x[kk][S] = 0.0;
          xZ[kk] = vZ[kk];
        }
#pragma unroll
        for (int b = 1; b < S; ++b)
        {
          const double mp_ = Fs[b - 1], gb = Fs[S + b], wb = Fs[2 * S + b];
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
          {
            if (b >= 2)
              x[kk][b] -= mp_ * x[kk][b - 1];
            x[kk][0] -= gb * x[kk][b];
            xZ[kk] -= wb * x[kk][b];
          }
        }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
        {
          const double xf = x[kk][0], xz = xZ[kk];
          x[kk][0] = (xf * S_ZZ - S_FZ * xz) * idet;
          xZ[kk] = (S_FF * xz - S_FZ * xf) * idet;
        }
#pragma unroll
        for (int b = S - 1; b >= 1; --b)
        {
          const double mb = Fs[b], gb = Fs[S + b], wb = Fs[2 * S + b], ib = Fs[3 * S + b];
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            x[kk][b] = x[kk][b] * ib - mb * x[kk][b + 1] - gb * x[kk][0] - wb * xZ[kk];
        }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
        {
#pragma unroll
          for (int b = 0; b < S; ++b)
            Xs[(kk * (S + 1) + b) * KS::XROW + j] = x[kk][b];
          Xs[(kk * (S + 1) + S) * KS::XROW + j] = xZ[kk];
        }
        __syncwarp();
        // ---- row r = j of K' = sum_k B_k^T X_k (symmetric positive definite) ----
        double Kr[S];
#pragma unroll
        for (int c2 = 0; c2 < S; ++c2)
        {
          double acc = 0.0;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
          {
            const double* Xk = Xs + kk * (S + 1) * KS::XROW + c2;
            acc += vA[kk] * Xk[prv * KS::XROW] + vB[kk] * Xk[j * KS::XROW] + vC[kk] * Xk[nxt * KS::XROW]
                   + vZ[kk] * Xk[S * KS::XROW];
          }
          Kr[c2] = acc;
        }
        if (!active)
        {
#pragma unroll
          for (int c2 = 0; c2 < S; ++c2)
            Kr[c2] = (c2 == j) ? 1.0 : 0.0;
          rhs = 0.0;
        }
        // ---- K' w = rhs: Gaussian elimination without pivoting, row r in lane r ----
        double myip = 1.0;
#pragma unroll
        for (int pv_ = 0; pv_ < S; ++pv_)
        {
          const double piv = __shfl_sync(FULL, Kr[pv_], pv_, S);
          const double ip = eqlb_rcp(piv);
          if (j == pv_)
            myip = ip;
          if (pv_ < S - 1)
          {
            const double prhs = __shfl_sync(FULL, rhs, pv_, S);
            const double lfac = (j > pv_) ? Kr[pv_] * ip : 0.0;
            rhs -= lfac * prhs;
#pragma unroll
            for (int c2 = pv_ + 1; c2 < S; ++c2)
              Kr[c2] -= lfac * __shfl_sync(FULL, Kr[c2], pv_, S);
          }
        }
        double wv[S];
#pragma unroll
        for (int pv_ = S - 1; pv_ >= 0; --pv_)
        {
          wv[pv_] = __shfl_sync(FULL, rhs * myip, pv_, S);
          if (j < pv_)
            rhs -= Kr[pv_] * wv[pv_];
        }
        // ---- tau_k = -X_k w on the own facet row and on d0; correct both rows ----
        double ub[2], uz[2];
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
        {
          const double* Xk = Xs + kk * (S + 1) * KS::XROW;
          double a0 = 0.0, a1 = 0.0;
#pragma unroll
          for (int c2 = 0; c2 < S; ++c2)
          {
            a0 -= Xk[j * KS::XROW + c2] * wv[c2];
            a1 -= Xk[S * KS::XROW + c2] * wv[c2];
          }
          ub[kk] = a0;
          uz[kk] = a1;
        }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
        {
          const double uhi = __shfl_sync(FULL, ub[kk], nxt, S);
          if (active)
          {
            double um0 = p_em * uz[kk], um1 = p_em * ub[kk];
            if (rev0)
            {
              const double t0 = -um0, t1 = -um0 + um1;
              um0 = t0;
              um1 = t1;
            }
            double* tgt = kk ? co : cf0;
            tgt[0] += um0;
            tgt[1] += um1;
            tgt[2] += p_ea * uz[kk];
            tgt[3] += p_ea * uhi;
          }
        }
        __syncwarp();
        if (active)
        {
          double* d0 = ptrs.S[0] + (size_t)c * nrt;
          atomicAdd(d0 + fm * 2, cf0[0]);
          atomicAdd(d0 + fm * 2 + 1, cf0[1]);
          atomicAdd(d0 + fp * 2, cf0[2]);
          atomicAdd(d0 + fp * 2 + 1, cf0[3]);
          atomicAdd(d0 + 6, cf0[4]);
          atomicAdd(d0 + 7, cf0[5]);
        }
      }
    }

    // ---- accumulate ----
    if constexpr (EV)
    {
      if (dep_pending)
      {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        dep_pending = false;
      }
    }
    if (active)
    {
      if (EV)
      {
        for (int side = (first_c && !internal) ? 0 : 1; side < 2; ++side)
        {
          const bool refl = (info & (side ? 512 : 256)) != 0;
          const double cl0 = side ? co[2] : co[0], cl1 = side ? co[3] : co[1];
          const double cg0 = refl ? -cl0 : cl0;
          const double cg1 = refl ? (-cl0 + cl1) : cl1;
          double* d = sig + (size_t)(side ? rc.w : rc.z) * k;
          if (use_atomics)
          {
            atomicAdd(d, cg0);
            atomicAdd(d + 1, cg1);
          }
          else
          {
            double2 vv = *reinterpret_cast<double2*>(d);
            vv.x += cg0;
            vv.y += cg1;
            *reinterpret_cast<double2*>(d) = vv;
          }
        }
        double* dstc = sig + (size_t)nfct * k + (size_t)c * 2;
        if (use_atomics)
        {
          atomicAdd(dstc, co[4]);
          atomicAdd(dstc + 1, co[5]);
        }
        else
        {
          double2 vv = *reinterpret_cast<double2*>(dstc);
          vv.x += co[4];
          vv.y += co[5];
          *reinterpret_cast<double2*>(dstc) = vv;
        }
      }
      else
      {
        double* d = sig + (size_t)c * nrt;
        if (use_atomics)
        {
          atomicAdd(d + fm * 2, co[0]);
          atomicAdd(d + fm * 2 + 1, co[1]);
          atomicAdd(d + fp * 2, co[2]);
          atomicAdd(d + fp * 2 + 1, co[3]);
          atomicAdd(d + 6, co[4]);
          atomicAdd(d + 7, co[5]);
        }
        else
        {
          double2* d2 = reinterpret_cast<double2*>(d);
          double2 vlo = d2[fm], vhi = d2[fp], vdv = d2[3];
          vlo.x += co[0];
          vlo.y += co[1];
          vhi.x += co[2];
          vhi.y += co[3];
          vdv.x += co[4];
          vdv.y += co[5];
          d2[fm] = vlo;
          d2[fp] = vhi;
          d2[3] = vdv;
        }
      }
    }
  }
  }
}


// ---------------------------------------------------------------------------
// Thread-per-patch variant with warp-cooperative staging (round 2).
// ncu on the lane-per-cell kernel above (profiles/r2a_*): the LSU data pipe (warp shuffles of the
// scans / lane-to-lane elimination + per-lane table reads) is 70 % busy, issue slots 54 %, the
// FP64 pipe 43 %; 11 % of all instructions are shuffles, 11 % selects of predicated lane code, and
// the serial elimination runs at 1/S lane efficiency.  Here one THREAD owns a patch:
//   load phase    the warp copies the J / G / f records of the 32*S patch cells of its batch into
//                 shared memory, lane per cell (coalesced 512-byte requests, cp.async);
//   compute phase every thread walks the fan of its own patch out of shared memory: explicit sweep,
//                 cell tensors, bordered tridiagonal elimination, back substitution - no shuffles,
//                 no lane predication, table reads are warp-uniform on structured meshes;
//                 the 13-double slot of a finished cell is reused for the state the back
//                 substitution needs (cfv, chain row), then for the corrector of the cell;
//   output phase  lane per cell again: coalesced accumulation into sigma (RED).
// Same mathematics and the same records / tables as patch_k2w_kernel (single RHS, no stress).
// Staging area per warp: [a][5 pair fields][32] double2 + [a][3 single fields][32] double + [a][32] info
// ints; the patch column is XOR-swizzled with the fan position a, so that both the lane-per-cell copies
// (8 cells of one patch per request) and the thread-per-patch reads spread over the banks.
// ---------------------------------------------------------------------------
template <int S>
struct K2T
{
  static constexpr int ROW = 32;
  static constexpr int PAIRS = S * 5 * ROW * 2;          // doubles
  static constexpr int SINGLES = S * 3 * ROW;            // doubles
  static constexpr int INFO = (S * ROW + 1) / 2;         // ints stored in double slots
  static constexpr int WARP_DOUBLES = PAIRS + SINGLES + INFO + ((PAIRS + SINGLES + INFO) & 1);
};

template <int S>
__device__ __forceinline__ double2* k2t_pair(double* w, int a, int f, int p)
{
  return reinterpret_cast<double2*>(w) + (a * 5 + f) * K2T<S>::ROW + (p ^ a);
}
template <int S>
__device__ __forceinline__ double* k2t_single(double* w, int a, int f, int p)
{
  return w + K2T<S>::PAIRS + (a * 3 + f) * K2T<S>::ROW + (p ^ a);
}
template <int S>
__device__ __forceinline__ int* k2t_info(double* w, int a, int p)
{
  return reinterpret_cast<int*>(w + K2T<S>::PAIRS + K2T<S>::SINGLES) + a * K2T<S>::ROW + (p ^ a);
}

template <bool EV, int S, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
patch_k2t_kernel(PatchView pv, int first, int count, const double* __restrict__ k2tab, const double* __restrict__ cellJ,
                 RhsPtrs ptrs, const double* __restrict__ bflux, int use_atomics, const int4* __restrict__ rec, int nfct,
                 int nbatch)
{
  extern __shared__ double s_mem[];
  double* s_blk = s_mem;
  double* s_dgm = s_mem + 6 * K2_BLOCK;
  double* s_mono = s_dgm + 9;
  for (int i = threadIdx.x; i < K2_TAB; i += blockDim.x)
    s_mem[i] = k2tab[i];
  __syncthreads();
  double* sw = s_mem + K2_TAB + (threadIdx.x >> 5) * K2T<S>::WARP_DOUBLES;
  // programmatic dependent launch: tables, loads and all arithmetic of the first batch may overlap the
  // tail of the previous colour; sigma is first touched after griddepcontrol.wait
  bool dep_pending = true;
  asm volatile("griddepcontrol.launch_dependents;");

  constexpr int k = 2, nrt = 8;
  const int lane = threadIdx.x & 31;
  const double* __restrict__ G = ptrs.G[0];
  const double* __restrict__ Fv = ptrs.F[0];
  double* __restrict__ sig = ptrs.S[0];
  const int wstride = gridDim.x * (blockDim.x >> 5);
  for (int bt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); bt < nbatch; bt += wstride)
  {
    const int pbase = bt * 32;  // first patch of the batch (within the launch segment)
    // ---------------- load phase: lane per cell ----------------
#pragma unroll
    for (int r = 0; r < S; ++r)
    {
      const int idx = r * 32 + lane;       // cell slot within the batch
      const int pl = idx / S, a = idx % S; // patch within the batch, position in the fan
      int4 rc = make_int4(0, 0, 0, 0);
      if (pbase + pl < count)
        rc = rec[(size_t)(pbase + pl) * S + a];
      const int32_t c = rc.x;
      const double2* js = reinterpret_cast<const double2*>(cellJ) + 2 * (size_t)c;
      const double2* gs = reinterpret_cast<const double2*>(G) + 3 * (size_t)c;
      const double* fs = Fv + 3 * (size_t)c;
      cp_async16(k2t_pair<S>(sw, a, 0, pl), js);
      cp_async16(k2t_pair<S>(sw, a, 1, pl), js + 1);
      cp_async16(k2t_pair<S>(sw, a, 2, pl), gs);
      cp_async16(k2t_pair<S>(sw, a, 3, pl), gs + 1);
      cp_async16(k2t_pair<S>(sw, a, 4, pl), gs + 2);
      cp_async8(k2t_single<S>(sw, a, 0, pl), fs);
      cp_async8(k2t_single<S>(sw, a, 1, pl), fs + 1);
      cp_async8(k2t_single<S>(sw, a, 2, pl), fs + 2);
      *k2t_info<S>(sw, a, pl) = rc.y;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    // ---------------- compute phase: thread per patch ----------------
    {
      const int pl = lane;
      const bool valid = pbase + pl < count;
      const size_t ip = (size_t)first + (valid ? pbase + pl : 0);
      const int nc = valid ? (*k2t_info<S>(sw, 0, pl) >> 16) : 0;
      const uint8_t ri = valid ? pv.rhsinfo[ip] : 0;
      const int ptype = ri & 3;
      const bool bc_e0 = (ri & 8) != 0, bc_en = (ri & 16) != 0;
      const bool internal = (ptype == EQLB_PATCH_INTERNAL);
      const bool req_bc = (ptype == EQLB_PATCH_ESSNT_DUAL || ptype == EQLB_PATCH_MIXED);
      const bool mark_z = req_bc, mark_f0 = req_bc, mark_fn = (ptype == EQLB_PATCH_ESSNT_DUAL);
      const int nch = internal ? nc - 1 : nc;

      auto raw_of = [&](int a)
      {
        K2Raw r;
        r.j0 = *k2t_pair<S>(sw, a, 0, pl);
        r.j1 = *k2t_pair<S>(sw, a, 1, pl);
        r.g0 = *k2t_pair<S>(sw, a, 2, pl);
        r.g1 = *k2t_pair<S>(sw, a, 3, pl);
        r.g2 = *k2t_pair<S>(sw, a, 4, pl);
        r.f0 = *k2t_single<S>(sw, a, 0, pl);
        r.f1 = *k2t_single<S>(sw, a, 1, pl);
        r.f2 = *k2t_single<S>(sw, a, 2, pl);
        return r;
      };
      // boundary value of a patch-boundary cell (first / last cell of a boundary patch)
      auto bc_of = [&](int a, int info, bool first_c, double& bv0, double& bv1) -> bool
      {
        bv0 = bv1 = 0.0;
        const bool last_c = (a == nc - 1);
        if (internal || !(first_c || last_c))
          return false;
        bool has_bc = false;
        if (ptype == EQLB_PATCH_ESSNT_DUAL)
          has_bc = true;
        else if (ptype == EQLB_PATCH_MIXED)
          has_bc = first_c ? bc_e0 : bc_en;
        if (has_bc)
        {
          const int fm = (info >> 2) & 3, fp = (info >> 4) & 3;
          const int32_t c = rec[(size_t)(pbase + pl) * S + a].x;
          const double* blk = s_blk + combo_of(fm, fp) * K2_BLOCK;
          patch_bc(bflux + (size_t)c * nrt + (first_c ? fm : fp) * k, blk + K2_O_BC + (first_c ? 0 : 4), bv0, bv1);
        }
        return has_bc;
      };
      // sign of det J and the prefactors of the two patch facets of a cell
      auto signs_of = [&](int a, int info, double& pm, double& pp) -> double
      {
        const double2 j0 = *k2t_pair<S>(sw, a, 0, pl), j1 = *k2t_pair<S>(sw, a, 1, pl);
        const double det = j0.x * j1.y - j0.y * j1.x;
        const double sgn = det > 0.0 ? 1.0 : -1.0;
        const int fm = (info >> 2) & 3, fp = (info >> 4) & 3;
        pm = (fm == 1) ? sgn : -sgn;
        pp = (fp == 1) ? sgn : -sgn;
        return det;
      };

      // ---- EV: mean-value shift (ev/assembly.hpp:283-298) needs sums over the patch first ----
      double lam = 0.0;
      if (EV)
      {
        double tot = 0.0, area2 = 0.0;
        for (int a = 0; a < nc; ++a)
        {
          const int info = *k2t_info<S>(sw, a, pl) & 0xffff;
          K2Cell cl;
          load_cell<true>(cl, 0, info, raw_of(a), s_blk, s_dgm);
          const double sgn = cl.det > 0.0 ? 1.0 : -1.0;
          tot += sgn * cl.cm[0];
          area2 += fabs(cl.det);
          double bv0, bv1;
          if (bc_of(a, info, a == 0, bv0, bv1) && ptype == EQLB_PATCH_ESSNT_DUAL)
            tot -= ((a == 0) ? cl.pm : cl.pp) * bv0;
        }
        if (valid && nc > 0 && (internal || ptype == EQLB_PATCH_ESSNT_DUAL))
          lam = tot * eqlb_rcp(0.5 * area2);
      }

      // neighbour data: prefactor pp of the previous cell (cyclic on interior patches), lo-facet data of
      // the next cell (cyclic)
      double pp_prev = 0.0, mp0_prev = 0.0;
      double f_pm = 0.0, f_m0 = 0.0, f_m1 = 0.0;  // first cell
      if (nc > 0)
      {
        double pmx, ppx;
        if (internal)
        {
          signs_of(nc - 1, *k2t_info<S>(sw, nc - 1, pl) & 0xffff, pmx, ppx);
          pp_prev = ppx;
        }
      }
      double c_run = 0.0;                                    // running sum of the explicit sweep
      double bFF = 0.0, bFZ = 0.0, bZZ = 0.0, bLF = 0.0, bLZ = 0.0;  // border (E_0, d0)
      double oD = 0.0, oG = 0.0, oW = 0.0, oL = 0.0;         // elimination carry
      double h22 = 0.0, h12 = 0.0, hl2 = 0.0, h02 = 0.0;     // hi-facet part of the previous cell
      K2Cell cur, nxt;
      int info_n = (nc > 0) ? (*k2t_info<S>(sw, 0, pl) & 0xffff) : 0;
      if (nc > 0)
      {
        load_cell<EV>(nxt, 0, info_n, raw_of(0), s_blk, s_dgm);
        f_pm = nxt.pm;
        f_m0 = nxt.mm[0];
        f_m1 = nxt.mm[1];
      }

      auto eliminate = [&](int b, double D, double e, double g, double w, double l)
      {
        D -= oD;
        g -= oG;
        w -= oW;
        l -= oL;
        const double ipv = eqlb_rcp(D);
        const double ei = e * ipv, gi = g * ipv, wi = w * ipv;
        oD = ei * e;
        oG = ei * g;
        oW = ei * w;
        oL = ei * l;
        bFF -= gi * g;
        bFZ -= gi * w;
        bZZ -= wi * w;
        bLF -= gi * l;
        bLZ -= wi * l;
        // chain row b lives in the (finished) slot b-1
        *k2t_pair<S>(sw, b - 1, 3, pl) = make_double2(ipv, e);
        *k2t_pair<S>(sw, b - 1, 4, pl) = make_double2(g, w);
        *k2t_single<S>(sw, b - 1, 0, pl) = l;
      };

      for (int a = 0; a < nc; ++a)
      {
        cur = nxt;
        const int info = info_n;
        const bool first_c = (a == 0), last_c = (a == nc - 1);
        // lo-facet data of the next cell of the fan (the first cell closes the cycle)
        double n_pm = f_pm, n_m0 = f_m0, n_m1 = f_m1;
        if (!last_c)
        {
          info_n = *k2t_info<S>(sw, a + 1, pl) & 0xffff;
          load_cell<EV>(nxt, 0, info_n, raw_of(a + 1), s_blk, s_dgm);
          n_pm = nxt.pm;
          n_m0 = nxt.mm[0];
          n_m1 = nxt.mm[1];
        }
        const int fm = (info >> 2) & 3, fp = (info >> 4) & 3;
        const bool rev0 = (info & 64) != 0, rev1 = (info & 128) != 0;
        const double* blk = s_blk + combo_of(fm, fp) * K2_BLOCK;
        const double sgn = cur.det > 0.0 ? 1.0 : -1.0;
        const bool on_bnd = !internal && (first_c || last_c);
        double bv0, bv1;
        const bool has_bc = bc_of(a, info, first_c, bv0, bv1);
        if (EV)
        {
#pragma unroll
          for (int t = 0; t < K2_NT; ++t)
            cur.cm[t] -= lam * cur.det * s_mono[t];
        }

        // ---- step 1: explicit sweep ----
        double cfv[6] = {0.0, 0.0, 0.0, 0.0, cur.cm[1], cur.cm[2]};
        double surf = 0.0;
        if (!EV)
        {
          if (!first_c)
            surf = -cur.mm[0] - pp_prev * cur.pm * mp0_prev;
          else if (!internal && (has_bc || ptype == EQLB_PATCH_MIXED))
          {
            cfv[1] += ((ptype == EQLB_PATCH_MIXED && !has_bc) ? 1.0 : -1.0) * cur.mm[1];
            if (has_bc)
              surf = -cur.mm[0];
          }
        }
        const double vol = sgn * cur.cm[0];
        const double t_add = cur.pm * surf + ((has_bc && first_c) ? cur.pm * bv0 : 0.0);
        c_run += vol - t_add;
        const double c_p = c_run;
        const double c_m = vol - c_p;
        if (has_bc)
        {
          if (first_c)
            cfv[1] += bv1;
          else
            cfv[3] += bv1;
        }
        if (!EV)
        {
          if (on_bnd && last_c)
            cfv[3] += (has_bc ? -1.0 : 1.0) * cur.mp[1];
          else
          {
            const double tau = -cur.pp * n_pm;
            const double mt1 = rev1 ? (n_m0 - n_m1) : n_m1;
            double hh = tau * mt1 - cur.mp[1];
            if (rev1 && !last_c)
              hh += -(tau * n_m0 - cur.mp[0]) + n_pm * c_p;
            cfv[3] += hh;
          }
        }
        else if (rev1 && !last_c)
          cfv[3] += n_pm * c_p;
        cfv[0] += cur.pm * c_m;
        cfv[2] += cur.pp * c_p;

        // ---- cell block of the RT mass matrix and load ----
        double MB[4][6];
        {
          const double2* tm = reinterpret_cast<const double2*>(blk + K2_O_MASS);
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int s2 = 0; s2 < 3; ++s2)
            {
              const int c0 = 2 * s2, c1 = 2 * s2 + 1;
              if (c1 < q)
              {
                MB[q][c0] = MB[c0][q];
                MB[q][c1] = MB[c1][q];
                continue;
              }
              const double2 a0 = tm[q * 3 + s2], a1 = tm[12 + q * 3 + s2], a2 = tm[24 + q * 3 + s2];
              MB[q][c0] = (c0 < q) ? MB[c0][q] : cur.g[0] * a0.x + cur.g[1] * a1.x + cur.g[2] * a2.x;
              MB[q][c1] = cur.g[0] * a0.y + cur.g[1] * a1.y + cur.g[2] * a2.y;
            }
        }
        double y[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
          double s2 = 0.0;
#pragma unroll
          for (int c2 = 0; c2 < 6; ++c2)
            s2 += MB[q][c2] * cfv[c2];
          y[q] = s2;
        }
        if (EV)
        {
          const double2* hh = reinterpret_cast<const double2*>(blk + K2_O_H);
#pragma unroll
          for (int mI = 0; mI < 3; ++mI)
          {
            const double gx = cur.G[2 * mI], gy = cur.G[2 * mI + 1];
            const double jg0 = sgn * (cur.adj[3] * gx - cur.adj[2] * gy);
            const double jg1 = sgn * (-cur.adj[1] * gx + cur.adj[0] * gy);
#pragma unroll
            for (int q = 0; q < 4; ++q)
            {
              const double2 hv = hh[mI * 4 + q];
              y[q] -= jg0 * hv.x + jg1 * hv.y;
            }
          }
        }
        if (rev0)
        {
#pragma unroll
          for (int s2 = 0; s2 < 4; ++s2)
            MB[0][s2] = -MB[0][s2] - MB[1][s2];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            MB[q][0] = -MB[q][0] - MB[q][1];
          y[0] = -y[0] - y[1];
        }
        const double p_ea = -cur.pp;
        const double p_em = rev0 ? -pp_prev : cur.pm;
        const double pe = p_em * p_ea;
        const bool hi_is_F = internal && last_c;
        const bool m_lo = first_c ? mark_f0 : false;
        const bool m_hi = (!internal && last_c) ? mark_fn : false;
        double T00 = MB[1][1], T22 = MB[3][3], T02 = pe * MB[1][3];
        double T11 = MB[0][0] + MB[2][2] + 2.0 * pe * MB[0][2];
        double T01 = MB[1][0] + pe * MB[1][2], T12 = pe * MB[3][0] + MB[3][2];
        double l0 = -p_em * y[1], l1 = -(p_em * y[0] + p_ea * y[2]), l2 = -p_ea * y[3];
        if (m_lo)
          T00 = T01 = T02 = l0 = 0.0;
        if (m_hi)
          T22 = T12 = T02 = l2 = 0.0;
        if (mark_z)
          T11 = T01 = T12 = l1 = 0.0;

        // the slot of cell a is free now: state for the map-back
        *k2t_pair<S>(sw, a, 0, pl) = make_double2(cfv[0], cfv[1]);
        *k2t_pair<S>(sw, a, 1, pl) = make_double2(cfv[2], cfv[3]);
        *k2t_pair<S>(sw, a, 2, pl) = make_double2(cfv[4], cfv[5]);
        *k2t_single<S>(sw, a, 1, pl) = p_em;
        *k2t_single<S>(sw, a, 2, pl) = p_ea;

        // ---- patch system: chain facet E_a (a >= 1) is complete with this cell ----
        bZZ += T11;
        bLZ += l1;
        if (first_c)
        {
          bFF += T00;
          bFZ += T01;
          bLF += l0;
        }
        else
        {
          const double g = (a == 1 ? h02 : 0.0) + (hi_is_F ? T02 : 0.0);
          const double e = (!hi_is_F && a < nch) ? T02 : 0.0;
          eliminate(a, T00 + h22, e, g, T01 + h12, l0 + hl2);
        }
        if (hi_is_F)
        {
          bFF += T22;
          bFZ += T12;
          bLF += l2;
        }
        h22 = T22;
        h12 = T12;
        hl2 = l2;
        h02 = T02;
        pp_prev = cur.pp;
        mp0_prev = cur.mp[0];
      }
      if (!internal && nc > 0)
      {
        // last facet E_nc of a boundary patch: fed by the last cell only
        if (mark_fn)
          eliminate(nc, 1.0, 0.0, 0.0, 0.0, 0.0);  // identity row (the carry is zero: the last cell's hi part was masked)
        else
          eliminate(nc, h22, 0.0, 0.0, h12, hl2);
      }

      // ---- border 2 x 2 system ----
      double S_FF = bFF, S_FZ = bFZ, S_ZZ = bZZ;
      if (mark_z)
        S_ZZ = 1.0;
      if (mark_f0)
        S_FF = 1.0;
      double u_F = 0.0, u_Z = 0.0;
      if (valid && nc > 0)
      {
        const double idet = eqlb_rcp(S_FF * S_ZZ - S_FZ * S_FZ);
        u_F = (bLF * S_ZZ - S_FZ * bLZ) * idet;
        u_Z = (S_FF * bLZ - S_FZ * bLF) * idet;
      }

      // ---- back substitution + map back, descending along the fan ----
      double u_up = 0.0;  // u of chain facet a+1
      if (!internal && nc > 0)
      {
        const double2 ie = *k2t_pair<S>(sw, nc - 1, 3, pl), gw = *k2t_pair<S>(sw, nc - 1, 4, pl);
        const double l = *k2t_single<S>(sw, nc - 1, 0, pl);
        u_up = (l - gw.x * u_F - gw.y * u_Z) * ie.x;
      }
      for (int a = nc - 1; a >= 0; --a)
      {
        const bool first_c = (a == 0), hi_is_F = internal && (a == nc - 1);
        double u_lo = u_F;
        if (!first_c)
        {
          const double2 ie = *k2t_pair<S>(sw, a - 1, 3, pl), gw = *k2t_pair<S>(sw, a - 1, 4, pl);
          const double l = *k2t_single<S>(sw, a - 1, 0, pl);
          u_lo = (l - ie.y * u_up - gw.x * u_F - gw.y * u_Z) * ie.x;
        }
        const double u_hi = hi_is_F ? u_F : u_up;
        const double2 c01 = *k2t_pair<S>(sw, a, 0, pl), c23 = *k2t_pair<S>(sw, a, 1, pl);
        const double p_em = *k2t_single<S>(sw, a, 1, pl), p_ea = *k2t_single<S>(sw, a, 2, pl);
        const bool rev0 = (*k2t_info<S>(sw, a, pl) & 64) != 0;
        double um0 = p_em * u_Z, um1 = p_em * u_lo;
        if (rev0)
        {
          const double t0 = -um0, t1 = -um0 + um1;
          um0 = t0;
          um1 = t1;
        }
        *k2t_pair<S>(sw, a, 0, pl) = make_double2(c01.x + um0, c01.y + um1);
        *k2t_pair<S>(sw, a, 1, pl) = make_double2(c23.x + p_ea * u_Z, c23.y + p_ea * u_hi);
        u_up = u_lo;
      }
    }
    __syncwarp();

    // ---------------- output phase: lane per cell ----------------
    if (dep_pending)
    {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      dep_pending = false;
    }
#pragma unroll
    for (int r = 0; r < S; ++r)
    {
      const int idx = r * 32 + lane;
      const int pl = idx / S, a = idx % S;
      if (pbase + pl >= count)
        continue;
      const int4 rc = rec[(size_t)(pbase + pl) * S + a];
      const int nc = rc.y >> 16, info = rc.y & 0xffff;
      if (a >= nc)
        continue;
      const int32_t c = rc.x;
      const int fm = (info >> 2) & 3, fp = (info >> 4) & 3;
      const double2 c01 = *k2t_pair<S>(sw, a, 0, pl), c23 = *k2t_pair<S>(sw, a, 1, pl), c45 = *k2t_pair<S>(sw, a, 2, pl);
      if (EV)
      {
        const bool internal = (pv.rhsinfo[(size_t)first + pbase + pl] & 3) == EQLB_PATCH_INTERNAL;
        for (int side = (a == 0 && !internal) ? 0 : 1; side < 2; ++side)
        {
          const bool refl = (info & (side ? 512 : 256)) != 0;
          const double cl0 = side ? c23.x : c01.x, cl1 = side ? c23.y : c01.y;
          const double cg0 = refl ? -cl0 : cl0;
          const double cg1 = refl ? (-cl0 + cl1) : cl1;
          double* d = sig + (size_t)(side ? rc.w : rc.z) * k;
          if (use_atomics)
          {
            atomicAdd(d, cg0);
            atomicAdd(d + 1, cg1);
          }
          else
          {
            double2 vv = *reinterpret_cast<double2*>(d);
            vv.x += cg0;
            vv.y += cg1;
            *reinterpret_cast<double2*>(d) = vv;
          }
        }
        double* dstc = sig + (size_t)nfct * k + (size_t)c * 2;
        if (use_atomics)
        {
          atomicAdd(dstc, c45.x);
          atomicAdd(dstc + 1, c45.y);
        }
        else
        {
          double2 vv = *reinterpret_cast<double2*>(dstc);
          vv.x += c45.x;
          vv.y += c45.y;
          *reinterpret_cast<double2*>(dstc) = vv;
        }
      }
      else
      {
        double* d = sig + (size_t)c * nrt;
        if (use_atomics)
        {
          atomicAdd(d + fm * 2, c01.x);
          atomicAdd(d + fm * 2 + 1, c01.y);
          atomicAdd(d + fp * 2, c23.x);
          atomicAdd(d + fp * 2 + 1, c23.y);
          atomicAdd(d + 6, c45.x);
          atomicAdd(d + 7, c45.y);
        }
        else
        {
          double2* d2 = reinterpret_cast<double2*>(d);
          double2 vlo = d2[fm], vhi = d2[fp], vdv = d2[3];
          vlo.x += c01.x;
          vlo.y += c01.y;
          vhi.x += c23.x;
          vhi.y += c23.y;
          vdv.x += c45.x;
          vdv.y += c45.y;
          d2[fm] = vlo;
          d2[fp] = vhi;
          d2[3] = vdv;
        }
      }
    }
    __syncwarp();  // the staging area is overwritten by the next batch
  }
}

} // namespace

// gather the reference tables per local facet pair (fm, fp); v = 3 - fm - fp
void build_k2_tables(eqlb_handle* h, const eqlb_tables* t)
{
  std::vector<double> tab(K2_TAB_STRESS, 0.0);
  const int nrt = 8, ndg = 3, nt = 3, k = 2;
  for (int fm = 0; fm < 3; ++fm)
    for (int fp = 0; fp < 3; ++fp)
    {
      if (fm == fp)
        continue;
      const int v = 3 - fm - fp;
      double* blk = tab.data() + combo_of(fm, fp) * K2_BLOCK;
      auto rdof = [&](int q) { return q < 2 ? fm * 2 + q : (q < 4 ? fp * 2 + (q - 2) : 6 + (q - 4)); };
      for (int m = 0; m < 3; ++m)
        for (int q = 0; q < 4; ++q)
          for (int s = 0; s < 6; ++s)
            blk[K2_O_MASS + (m * 4 + q) * 6 + s] = t->rt_mass[((size_t)m * nrt + rdof(q)) * nrt + rdof(s)];
      for (int m = 0; m < ndg; ++m)
        for (int q = 0; q < 4; ++q)
          for (int d = 0; d < 2; ++d)
            blk[K2_O_H + (m * 4 + q) * 2 + d] = t->hat_dg_rt[(((size_t)v * ndg + m) * nrt + rdof(q)) * 2 + d];
      for (int tt = 0; tt < nt; ++tt)
        for (int i = 0; i < ndg; ++i)
        {
          blk[K2_O_CMF + tt * 3 + i] = t->cell_mom_f[((size_t)v * nt + tt) * ndg + i];
          for (int d = 0; d < 2; ++d)
            blk[K2_O_CMG + (tt * 3 + i) * 2 + d] = t->cell_mom_g[(((size_t)v * nt + tt) * ndg + i) * 2 + d];
        }
      // weak symmetry: int phi_q^d lambda_n, nodes ordered (patch node, outer node of E_a, outer node of E_a-1)
      {
        double* pb = tab.data() + K2_TAB + combo_of(fm, fp) * K2P_BLOCK;
        const int vl[3] = {v, fm, fp};
        for (int q = 0; q < 6; ++q)
          for (int d = 0; d < 2; ++d)
            for (int n = 0; n < 3; ++n)
              pb[(q * 2 + d) * 3 + n] = t->rt_p1[((size_t)rdof(q) * 2 + d) * 3 + vl[n]];
      }
      for (int side = 0; side < 2; ++side)
      {
        const int f = side ? fp : fm;
        for (int j = 0; j < k; ++j)
        {
          for (int i = 0; i < ndg; ++i)
            blk[K2_O_FM + side * 6 + j * 3 + i] = t->fct_mom[(((size_t)f * 3 + v) * k + j) * ndg + i];
          for (int i = 0; i < k; ++i)
            blk[K2_O_BC + side * 4 + j * 2 + i] = t->bc_mat[(((size_t)f * 3 + v) * k + j) * k + i];
        }
      }
    }
  for (int i = 0; i < 9; ++i)
    tab[6 * K2_BLOCK + i] = t->dg_mono[i];
  for (int i = 0; i < 3; ++i)
    tab[6 * K2_BLOCK + 9 + i] = t->mono_int[i];
  h->d_k2tab.upload(tab.data(), tab.size());
}

template <bool EV>
static void launch_k2_range(eqlb_handle* h, const RhsPtrs& ptrs, int first, int count, int use_atomics, int maxnf, int lanes,
                            int64_t recoff, bool stress, bool pdl)
{
  if (count <= 0)
    return;
  const int bs = 128;
  const PatchView pv = h->patch_view();
  const size_t bstride = (size_t)h->ncell * h->nrt;
  if (lanes > 0 && recoff >= 0)
  {
    // warp-cooperative kernel: S lanes per patch
    const size_t smem = (size_t)K2_TAB * sizeof(double);
    const int S = lanes;
    const int nwt = (count + (32 / S) - 1) / (32 / S);  // warp tiles
    static const int waves = getenv("EQLB_K2W_WAVES") ? atoi(getenv("EQLB_K2W_WAVES")) : 1;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
    if (stress)
    {
      // SE with the weak-symmetry stage fused in (interior patches)
      if constexpr (!EV)
      {
        // resident CTAs per SM: 3 (168 registers, no spills) or 4 (128 registers, 20..120 spilled doubles);
        // EQLB_K2S_MINB selects, measurements in profiles/r2_kernel_experiments.md
        static const int minb_env = getenv("EQLB_K2S_MINB") ? atoi(getenv("EQLB_K2S_MINB")) : 3;
        const int minb = (minb_env == 4) ? 4 : 3;
        const int tile = (S == 4) ? K2Stress<4>::TILE : (S == 8 ? K2Stress<8>::TILE : K2Stress<16>::TILE);
        const size_t smem_s = ((size_t)K2_TAB_STRESS + (size_t)(bs / S) * tile) * sizeof(double);
        auto kern = (minb == 3) ? ((S == 4) ? patch_k2w_kernel<false, 4, 3, true>
                                            : (S == 8 ? patch_k2w_kernel<false, 8, 3, true> : patch_k2w_kernel<false, 16, 3, true>))
                                : ((S == 4) ? patch_k2w_kernel<false, 4, 4, true>
                                            : (S == 8 ? patch_k2w_kernel<false, 8, 4, true> : patch_k2w_kernel<false, 16, 4, true>));
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
        const int grid = std::max(1, std::min((nwt + 3) / 4, nsm * minb));
        kern<<<grid, bs, smem_s, h->stream>>>(pv, first, count, h->d_k2tab.p, h->d_cellJ.p, h->nrhs, ptrs, h->d_bflux.p, bstride,
                                              1, h->d_prec.p + recoff, h->nfct, nwt);
      }
      CUDA_CHECK(cudaGetLastError());
      h->launches++;
      return;
    }
    // 5 CTAs/SM (<= 96 registers) measured best on B200: 3 -> 0.87, 4 -> 0.77, 5 -> 0.76, 6 -> 0.75/0.81 ms (EV/SE)
    // EQLB_K2T=1: thread-per-patch kernel with warp-cooperative staging (single RHS).  Measured on B200
    // (profiles/r2_kernel_experiments.md): S=4 segments 12 warps/SM: on par with the lane-per-cell kernel;
    // S=8 segments (27.6 KB staging per warp -> 7 warps/SM): 48 % slower - the serial per-patch instruction
    // stream needs more resident warps than the staging area leaves room for.  Opt-in, default off.
    static const bool k2t_env = getenv("EQLB_K2T") && atoi(getenv("EQLB_K2T")) != 0;
    // EQLB_K2T_MASK: bit mask of the lane counts S (4 | 8 | 16) that use the thread-per-patch kernel
    static const int k2t_mask = getenv("EQLB_K2T_MASK") ? atoi(getenv("EQLB_K2T_MASK")) : 28;
    if (k2t_env && h->nrhs == 1 && (k2t_mask & S))
    {
      // warps per CTA / CTAs per SM sized by the staging area (227 KB per SM): S=4: 2 x 6, S=8: 1 x 7, S=16: 1 x 3 warps
      const int nbatch = (count + 31) / 32;
      const double* bfl = h->d_bflux.p;
      const int4* recp = h->d_prec.p + recoff;
      auto go = [&](auto kern, int nt, int minb_t, size_t warp_doubles)
      {
        const size_t smem_t = ((size_t)K2_TAB + (size_t)(nt / 32) * warp_doubles) * sizeof(double);
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
        const int wpc = nt / 32;
        const int grid = std::max(1, std::min((nbatch + wpc - 1) / wpc, nsm * minb_t));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(nt);
        cfg.dynamicSmemBytes = smem_t;
        cfg.stream = h->stream;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = pdl ? 1 : 0;
        CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, pv, first, count, (const double*)h->d_k2tab.p, (const double*)h->d_cellJ.p, ptrs,
                                      bfl, use_atomics, recp, h->nfct, nbatch));
      };
      if (S == 4)
        go(patch_k2t_kernel<EV, 4, 192, 2>, 192, 2, K2T<4>::WARP_DOUBLES);
      else if (S == 8)
        go(patch_k2t_kernel<EV, 8, 224, 1>, 224, 1, K2T<8>::WARP_DOUBLES);
      else
        go(patch_k2t_kernel<EV, 16, 96, 1>, 96, 1, K2T<16>::WARP_DOUBLES);
      CUDA_CHECK(cudaGetLastError());
      h->launches++;
      return;
    }
    constexpr int minb = 5;
    // single RHS: cell data software-pipelined through shared memory (cp.async, one tile ahead)
    // (EQLB_K2_PIPE=1; measured 0.6943 vs 0.6945 ms/step: the load latency it hides was already covered by
    //  the other resident warps, and the copies add LSU traffic - opt-in, default off)
    static const bool pipe_env = getenv("EQLB_K2_PIPE") && atoi(getenv("EQLB_K2_PIPE")) != 0;
    const bool pipe = pipe_env && h->nrhs == 1;
    auto kern = pipe ? ((S == 4) ? patch_k2w_kernel<EV, 4, minb, false, true>
                                 : (S == 8 ? patch_k2w_kernel<EV, 8, minb, false, true> : patch_k2w_kernel<EV, 16, minb, false, true>))
                     : ((S == 4) ? patch_k2w_kernel<EV, 4, minb, false>
                                 : (S == 8 ? patch_k2w_kernel<EV, 8, minb, false> : patch_k2w_kernel<EV, 16, minb, false>));
    const size_t smem_p = smem + (pipe ? (size_t)(bs / 32) * 2 * K2_STAGE_DOUBLES * sizeof(double) : 0);
    const int grid = std::max(1, std::min((nwt + 3) / 4, nsm * minb * waves));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(bs);
    cfg.dynamicSmemBytes = smem_p;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = pdl ? 1 : 0;  // only behind one of our own launches of the same call (launch_patch_t)
    const double* bfl = h->d_bflux.p;
    const int4* recp = h->d_prec.p + recoff;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, pv, first, count, (const double*)h->d_k2tab.p, (const double*)h->d_cellJ.p, h->nrhs,
                                  ptrs, bfl, bstride, use_atomics, recp, h->nfct, nwt));
  }
  else
    throw EqlbError(EQLB_ERR_STATE, "degree-2 kernel: segment without lane records");
  CUDA_CHECK(cudaGetLastError());
  h->launches++;
}

void launch_k2(eqlb_handle* h, bool ev, const RhsPtrs& ptrs, int first, int count, int use_atomics, int maxnf, int lanes,
               int64_t recoff, bool stress, bool pdl)
{
  if (stress && (ev || lanes <= 0 || recoff < 0 || h->nrhs < 2))
    throw EqlbError(EQLB_ERR_STATE, "degree-2 stress kernel: needs SE, lane records and at least two rows");
  if (ev)
    launch_k2_range<true>(h, ptrs, first, count, use_atomics, maxnf, lanes, recoff, false, pdl);
  else
    launch_k2_range<false>(h, ptrs, first, count, use_atomics, maxnf, lanes, recoff, stress, pdl);
}
