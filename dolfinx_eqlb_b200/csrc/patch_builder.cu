// Device patch builder: ordered cell/facet fan around every vertex, patch
// classification per RHS, reversed-facet flags; plus the expansion of the SE
// 4-plane DOF map for bit-exact comparison with the reference layout.
//
// Reference semantics (not code): se::OrientedPatch::initialize_patch
// `se/Patch.cpp:406-635`, next_facet `:682-759`, reversion_required `:106-128`,
// Patch::create_subdofmap `se/Patch.hpp:792-898`, set_assembly_informations
// `:710-789`, reversed-facet detection `se/solve_patch_semiexplt.hpp:325-389`,
// set_boundary_markers `se/assembly.hpp:46-98`.
//
// One thread per patch: the walk around a vertex is inherently sequential
// (<= 16 steps) while the ~1e6..3e7 patches are independent; all loads are 4-byte
// gathers from CSR arrays that stay L2 resident.  The builder runs once per
// eqlb_set_bcs, not per equilibration.
#include <cmath>

#include <cub/device/device_radix_sort.cuh>

#include "eqlb_internal.cuh"

namespace
{

struct Fan
{
  int nc, nf;
  int32_t cells[EQLB_NCMAX + 2];
  int32_t fcts[EQLB_NCMAX + 2];
  int8_t inod[EQLB_NCMAX + 2];
  int8_t fl[2 * (EQLB_NCMAX + 1)];
  int8_t type0;
  int32_t fct_ep[2], fct_ef[2];
};

__device__ __forceinline__ int local_index3(const int32_t* __restrict__ arr3, int32_t val)
{
  return (arr3[0] == val) ? 0 : ((arr3[1] == val) ? 1 : 2);
}

// Ordered fan of cells / facets around `node`.
__device__ void build_fan(const MeshView& m, const int8_t* __restrict__ facet_type0, int node, Fan& F)
{
  const int c0 = m.node_cell_off[node], c1 = m.node_cell_off[node + 1];
  const int f0 = m.node_fct_off[node], f1 = m.node_fct_off[node + 1];
  const int nc = c1 - c0, nf = f1 - f0;
  F.nc = nc;
  F.nf = nf;
  F.type0 = EQLB_PATCH_INTERNAL;
  F.fct_ep[0] = F.fct_ep[1] = F.fct_ef[0] = F.fct_ef[1] = -1;
  int32_t fct_first = m.node_fct[f0];
  if (nf > nc)
  {
    for (int i = f0; i < f1; ++i)
    {
      const int32_t id = m.node_fct[i];
      const int8_t ft = facet_type0[id];
      if (ft == EQLB_FCT_ESSNT_PRIMAL)
      {
        if (F.fct_ep[0] < 0)
          F.fct_ep[0] = id;
        else
          F.fct_ep[1] = id;
      }
      else if (ft == EQLB_FCT_ESSNT_DUAL)
      {
        if (F.fct_ef[0] < 0)
          F.fct_ef[0] = id;
        else
          F.fct_ef[1] = id;
      }
    }
    if (F.fct_ef[0] < 0)
    {
      F.type0 = EQLB_PATCH_ESSNT_PRIMAL;
      fct_first = F.fct_ep[0];
    }
    else
    {
      F.type0 = (F.fct_ep[0] < 0) ? EQLB_PATCH_ESSNT_DUAL : EQLB_PATCH_MIXED;
      fct_first = F.fct_ef[0];
    }
  }
  const bool internal = (F.type0 == EQLB_PATCH_INTERNAL);
  int lloop = nc + 1;
  if (internal)
  {
    F.fcts[1] = fct_first;
    F.cells[1] = m.fct_cell[m.fct_cell_off[fct_first] + 1];
  }
  else
  {
    F.fcts[0] = fct_first;
    const int32_t c = m.fct_cell[m.fct_cell_off[fct_first]];
    F.cells[1] = c;
    const int lf = local_index3(m.cell_fct + 3 * c, fct_first);
    F.fl[0] = lf;
    F.fl[1] = lf;
    const int v = local_index3(m.cell_node + 3 * c, node);
    // next facet: the facet of the cell, other than lf, that contains the node
    // (facet f is opposite vertex f)  == result of se/Patch.cpp:682-759
    F.fcts[1] = m.cell_fct[3 * c + (3 - lf - v)];
    lloop = nc;
  }
  for (int a = 1; a < lloop; ++a)
  {
    const int32_t fct_a = F.fcts[a], cell_a = F.cells[a];
    const int o = m.fct_cell_off[fct_a];
    const int32_t ca = m.fct_cell[o], cb = m.fct_cell[o + 1];
    const int32_t cell_ap1 = (ca == cell_a) ? cb : ca;
    F.cells[a + 1] = cell_ap1;
    const int lf_ap1 = local_index3(m.cell_fct + 3 * cell_ap1, fct_a);
    F.fl[2 * a] = local_index3(m.cell_fct + 3 * cell_a, fct_a);
    F.fl[2 * a + 1] = lf_ap1;
    F.inod[a] = local_index3(m.cell_node + 3 * cell_a, node);
    const int v = local_index3(m.cell_node + 3 * cell_ap1, node);
    F.fcts[a + 1] = m.cell_fct[3 * cell_ap1 + (3 - lf_ap1 - v)];
  }
  if (!internal)
  {
    F.inod[nc] = local_index3(m.cell_node + 3 * F.cells[nc], node);
    const int lf = local_index3(m.cell_fct + 3 * F.cells[nc], F.fcts[nc]);
    F.fl[2 * nc] = lf;
    F.fl[2 * nc + 1] = lf;
  }
  else
  {
    F.cells[0] = F.cells[nc];
    F.cells[nc + 1] = F.cells[1];
    F.inod[0] = F.inod[nc];
    F.inod[nc + 1] = F.inod[1];
    F.fcts[0] = F.fcts[nf];
    F.fl[0] = F.fl[2 * nf];
    F.fl[1] = F.fl[2 * nf + 1];
  }
}

// patch type of RHS i (se/Patch.cpp:487-526)
__device__ __forceinline__ int8_t patch_type_rhs(const Fan& F, const int8_t* __restrict__ ft_i, int i)
{
  if (F.type0 == EQLB_PATCH_INTERNAL)
    return EQLB_PATCH_INTERNAL;
  if (i == 0)
    return F.type0;
  int32_t fa, fb;
  if (F.type0 == EQLB_PATCH_ESSNT_PRIMAL)
  {
    fa = F.fct_ep[0];
    fb = F.fct_ep[1];
  }
  else if (F.type0 == EQLB_PATCH_ESSNT_DUAL)
  {
    fa = F.fct_ef[0];
    fb = F.fct_ef[1];
  }
  else
  {
    fa = F.fct_ef[0];
    fb = F.fct_ep[0];
  }
  if (ft_i[fa] == ft_i[fb])
    return (ft_i[fa] == EQLB_FCT_ESSNT_PRIMAL) ? EQLB_PATCH_ESSNT_PRIMAL : EQLB_PATCH_ESSNT_DUAL;
  return EQLB_PATCH_MIXED;
}

// reversed(a, 0/1) flags (se/solve_patch_semiexplt.hpp:325-389)
__device__ __forceinline__ void reversed_flags(const MeshView& m, const Fan& F, int a, bool& r0, bool& r1)
{
  const bool bnd = (F.type0 != EQLB_PATCH_INTERNAL);
  const int32_t c = F.cells[a];
  r0 = false;
  r1 = false;
  if (!(bnd && a == 1))
  {
    const int32_t c_am1 = F.cells[a - 1];
    // local id of E_{a-1} in T_{a-1} = fl[2(a-1)] ; for the interior patch a-1 == 0
    // maps to fl[0] (= fl[2 nf])
    r0 = m.fct_perms[3 * c_am1 + F.fl[2 * (a - 1)]] != m.fct_perms[3 * c + F.fl[2 * a - 1]];
  }
  if (!(bnd && a == F.nc))
  {
    const int32_t c_ap1 = F.cells[a + 1];
    r1 = m.fct_perms[3 * c + F.fl[2 * a]] != m.fct_perms[3 * c_ap1 + F.fl[2 * a + 1]];
  }
}

__global__ void patch_builder_kernel(MeshView m, const int8_t* __restrict__ facet_type, int nrhs,
                                     const int32_t* __restrict__ order, int npatch, size_t stride, int ncmax,
                                     // compact records (colour order)
                                     int32_t* __restrict__ pnode, uint8_t* __restrict__ pncells,
                                     int32_t* __restrict__ pcell, uint16_t* __restrict__ pinfo,
                                     uint8_t* __restrict__ prhs, int4* __restrict__ prec,
                                     const int64_t* __restrict__ seginfo, int nseg,
                                     // expanded, reference layout (node order); may be null
                                     int32_t* __restrict__ x_ncells, int32_t* __restrict__ x_cells,
                                     int32_t* __restrict__ x_fcts, int8_t* __restrict__ x_inod,
                                     int8_t* __restrict__ x_fl, int8_t* __restrict__ x_type,
                                     uint8_t* __restrict__ x_rev, uint8_t* __restrict__ x_reversion)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npatch)
    return;
  const int node = order[i];
  Fan F;
  build_fan(m, facet_type, node, F);
  const int nc = F.nc;
  const bool internal = (F.type0 == EQLB_PATCH_INTERNAL);

  if (pnode)
  {
    pnode[i] = node;
    pncells[i] = (uint8_t)nc;
  }
  for (int a = 1; a <= nc; ++a)
  {
    bool r0, r1;
    reversed_flags(m, F, a, r0, r1);
    if (pcell)
    {
      pcell[(size_t)(a - 1) * stride + i] = F.cells[a];
      const int rho_m = m.fct_perms[3 * F.cells[a] + F.fl[2 * a - 1]] ? 256 : 0;
      const int rho_p = m.fct_perms[3 * F.cells[a] + F.fl[2 * a]] ? 512 : 0;
      pinfo[(size_t)(a - 1) * stride + i] = (uint16_t)(F.inod[a] | (F.fl[2 * a - 1] << 2) | (F.fl[2 * a] << 4)
                                                       | (r0 ? 64 : 0) | (r1 ? 128 : 0) | rho_m | rho_p);
    }
    if (x_rev)
    {
      x_rev[((size_t)node * ncmax + (a - 1)) * 2] = r0;
      x_rev[((size_t)node * ncmax + (a - 1)) * 2 + 1] = r1;
    }
  }
  if (prec)
  {
    // lane records of the warp-cooperative kernels (only for the eligible head of a segment)
    for (int sg = 0; sg < nseg; ++sg)
    {
      const int64_t first = seginfo[4 * sg], nfast = seginfo[4 * sg + 1], lanes = seginfo[4 * sg + 2];
      if (lanes == 0 || i < first || i >= first + nfast)
        continue;
      int4* dst = prec + seginfo[4 * sg + 3] + (i - first) * lanes;
      for (int a = 1; a <= (int)lanes; ++a)
      {
        int4 rc = make_int4(0, nc << 16, 0, 0);
        if (a <= nc)
        {
          bool r0, r1;
          reversed_flags(m, F, a, r0, r1);
          const int rho_m = m.fct_perms[3 * F.cells[a] + F.fl[2 * a - 1]] ? 256 : 0;
          const int rho_p = m.fct_perms[3 * F.cells[a] + F.fl[2 * a]] ? 512 : 0;
          rc.x = F.cells[a];
          rc.y |= F.inod[a] | (F.fl[2 * a - 1] << 2) | (F.fl[2 * a] << 4) | (r0 ? 64 : 0) | (r1 ? 128 : 0) | rho_m | rho_p;
          rc.z = F.fcts[a - 1];
          rc.w = F.fcts[a];
        }
        dst[a - 1] = rc;
      }
      break;
    }
  }
  int8_t type_prev = 0;
  for (int r = 0; r < nrhs; ++r)
  {
    const int8_t* ft = facet_type + (size_t)r * m.nfct;
    const int8_t tp = patch_type_rhs(F, ft, r);
    bool bc0 = false, bcn = false, rev = false;
    if (!internal)
    {
      bc0 = ft[F.fcts[0]] == EQLB_FCT_ESSNT_DUAL;
      bcn = ft[F.fcts[nc]] == EQLB_FCT_ESSNT_DUAL;
      const bool req = (tp == EQLB_PATCH_ESSNT_DUAL || tp == EQLB_PATCH_MIXED);
      if (r > 0 && req && (tp != type_prev || tp == EQLB_PATCH_MIXED) && !bc0)
        rev = true;
    }
    type_prev = tp;
    if (prhs)
      prhs[(size_t)r * stride + i] = (uint8_t)(tp | (rev ? 4 : 0) | (bc0 ? 8 : 0) | (bcn ? 16 : 0));
    if (x_type)
      x_type[(size_t)node * nrhs + r] = tp;
    if (x_reversion)
      x_reversion[(size_t)node * nrhs + r] = rev;
  }
  if (x_ncells)
    x_ncells[node] = nc;
  const int lo = internal ? 0 : 1;
  const int hi = internal ? nc + 2 : nc + 1;
  if (x_cells)
    for (int a = lo; a < hi; ++a)
      x_cells[(size_t)node * (ncmax + 2) + a] = F.cells[a];
  if (x_inod)
    for (int a = lo; a < hi; ++a)
      x_inod[(size_t)node * (ncmax + 2) + a] = F.inod[a];
  if (x_fcts)
    for (int a = 0; a < F.nf + (internal ? 1 : 0); ++a)
      x_fcts[(size_t)node * (ncmax + 2) + a] = F.fcts[a];
  if (x_fl)
    for (int a = 0; a < 2 * F.nf + (internal ? 2 : 0); ++a)
      x_fl[(size_t)node * 2 * (ncmax + 1) + a] = F.fl[a];
}

// ---------------------------------------------------------------------------
// SE DOF-map expansion (parity evidence only; the hot kernel evaluates the same
// index arithmetic on the fly instead of reading 4 planes of int32 from HBM)
// ---------------------------------------------------------------------------
__global__ void se_dofmap_kernel(MeshView m, const int8_t* __restrict__ facet_type, int nrhs, int npatch, int ncmax,
                                 int k, int nrt, int nadd, int ndiv, int ndg_fct, int p, bool stress,
                                 const int32_t* __restrict__ closure, const double* __restrict__ cellJ,
                                 int32_t* __restrict__ dofmap, int32_t* __restrict__ projflux,
                                 int8_t* __restrict__ bmarkers, int ndpc, int hzmax,
                                 const uint8_t* __restrict__ owned)
{
  const int node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= npatch || (owned && !owned[node]))
    return;
  Fan F;
  build_fan(m, facet_type, node, F);
  const int nc = F.nc, nf = F.nf;
  const bool internal = (F.type0 == EQLB_PATCH_INTERNAL);
  const int offs1 = k, offs2 = 2 * k, offs3 = 2 * k + nadd, offs4 = stress ? offs3 + 3 : offs3;
  const bool out[3] = {false, true, false};
  int32_t* dm = dofmap ? dofmap + (size_t)node * 4 * (ncmax + 2) * ndpc : nullptr;
#define DM(pl, a, i) dm[((size_t)(pl) * (ncmax + 2) + (a)) * ndpc + (i)]
  if (dm)
  {
    for (int pl = 0; pl < 4; ++pl)
      for (int a = 0; a < nc + 2; ++a)
        for (int i = 0; i < ndpc; ++i)
          DM(pl, a, i) = 0;
    for (int a = 1; a <= nc; ++a)
    {
      const int32_t cell = F.cells[a];
      const int fl_m = F.fl[2 * a - 1], fl_p = F.fl[2 * a];
      const int gdof = cell * nrt;
      if (k == 1)
      {
        DM(0, a, 0) = fl_m;
        DM(0, a, 1) = fl_p;
        DM(1, a, 0) = gdof + fl_m;
        DM(1, a, 1) = gdof + fl_p;
      }
      else
      {
        const int pd_m = (a - 1) * (k - 1);
        const int pd_p = (internal && a == nc) ? 0 : pd_m + k - 1;
        for (int ii = 0; ii < k; ++ii)
        {
          DM(0, a, ii) = fl_m * k + ii;
          DM(0, a, offs1 + ii) = fl_p * k + ii;
          DM(1, a, ii) = gdof + fl_m * k + ii;
          DM(1, a, offs1 + ii) = gdof + fl_p * k + ii;
          DM(2, a, ii) = ii == 0 ? 0 : pd_m + ii;
          DM(2, a, offs1 + ii) = ii == 0 ? 0 : pd_p + ii;
        }
        for (int ii = 0; ii < nadd; ++ii)
        {
          DM(0, a, offs2 + ii) = 3 * k + ndiv + ii;
          DM(1, a, offs2 + ii) = gdof + 3 * k + ndiv + ii;
          DM(2, a, offs2 + ii) = nf * (k - 1) + 1 + (a - 1) * nadd + ii;
          DM(3, a, offs2 + ii) = 1;
        }
        for (int ii = 0; ii < ndiv; ++ii)
        {
          DM(0, a, offs4 + ii) = 3 * k + ii;
          DM(1, a, offs4 + ii) = gdof + 3 * k + ii;
        }
      }
    }
    if (stress)
    {
      // weak-symmetry constraint DOFs (se/Patch.hpp:621-708): slot offs3 = patch node,
      // offs3+1 = outer node of E_a, offs3+2 = outer node of E_{a-1}
      for (int a = 1; a <= nc; ++a)
      {
        const int32_t cell = F.cells[a];
        DM(0, a, offs3) = F.inod[a];
        DM(2, a, offs3) = 0;
        DM(3, a, offs3) = 1;
        // outer node of E_a
        {
          const int32_t fct = F.fcts[a];
          const int32_t n0 = m.fct_node[2 * fct], n1 = m.fct_node[2 * fct + 1];
          const int32_t nd = (n0 == node) ? n1 : n0;
          DM(0, a, offs3 + 1) = local_index3(m.cell_node + 3 * cell, nd);
          DM(2, a, offs3 + 1) = (!internal && a == nc) ? nf : a;
          DM(3, a, offs3 + 1) = 1;
        }
        // outer node of E_{a-1}
        {
          const int32_t fct = F.fcts[a - 1];
          const int32_t n0 = m.fct_node[2 * fct], n1 = m.fct_node[2 * fct + 1];
          const int32_t nd = (n0 == node) ? n1 : n0;
          DM(0, a, offs3 + 2) = local_index3(m.cell_node + 3 * cell, nd);
          DM(2, a, offs3 + 2) = (a == 1) ? (internal ? nc : nf - 1) : a - 1;
          DM(3, a, offs3 + 2) = 1;
        }
      }
    }
    // prefactors (plane 3), set_assembly_informations
    for (int a = 1; a <= nc; ++a)
    {
      bool r0, r1;
      reversed_flags(m, F, a, r0, r1);
      const double* Jc = cellJ + 4 * (size_t)F.cells[a];
      const double det = Jc[0] * Jc[3] - Jc[1] * Jc[2];
      const int fl_m = F.fl[2 * a - 1], fl_p = F.fl[2 * a];
      int p_m, p_p;
      if (det < 0)
      {
        p_m = r0 ? DM(3, a - 1, k) : (out[fl_m] ? -1 : 1);
        p_p = out[fl_p] ? 1 : -1;
      }
      else
      {
        p_m = r0 ? DM(3, a - 1, k) : (out[fl_m] ? 1 : -1);
        p_p = out[fl_p] ? -1 : 1;
      }
      for (int i = 0; i < k; ++i)
      {
        DM(3, a, i) = p_m;
        DM(3, a, k + i) = p_p;
      }
    }
    if (internal)
    {
      bool r0, r1;
      reversed_flags(m, F, 1, r0, r1);
      if (r0)
        for (int i = 0; i < k; ++i)
          DM(3, 1, i) = DM(3, nc, k + i);
      for (int pl = 0; pl < 4; ++pl)
        for (int ii = 0; ii < ndpc; ++ii)
        {
          DM(pl, 0, ii) = DM(pl, nc, ii);
          DM(pl, nc + 1, ii) = DM(pl, 1, ii);
        }
    }
  }
#undef DM
  if (projflux)
  {
    int32_t* pf = projflux + (size_t)node * (ncmax + 1) * 2 * ndg_fct;
    for (int a = 0; a < (nc + 1) * 2 * ndg_fct; ++a)
      pf[a] = 0;
    if (k > 1 && p > 0)
    {
      for (int a = 1; a <= nc; ++a)
        for (int i = 0; i < ndg_fct; ++i)
        {
          pf[(a - 1) * 2 * ndg_fct + i] = closure[F.fl[2 * a - 1] * ndg_fct + i];
          pf[a * 2 * ndg_fct + ndg_fct + i] = closure[F.fl[2 * a] * ndg_fct + i];
        }
      for (int i = 0; i < ndg_fct; ++i)
      {
        if (internal)
        {
          pf[nc * 2 * ndg_fct + i] = pf[i];
          pf[ndg_fct + i] = pf[nc * 2 * ndg_fct + ndg_fct + i];
        }
        else
        {
          pf[ndg_fct + i] = pf[i];
          pf[nc * 2 * ndg_fct + i] = pf[nc * 2 * ndg_fct + ndg_fct + i];
        }
      }
    }
  }
  if (bmarkers)
  {
    const int hz = 1 + (k - 1) * nf + nadd * nc;
    int8_t type_prev = 0;
    for (int r = 0; r < nrhs; ++r)
    {
      const int8_t* ft = facet_type + (size_t)r * m.nfct;
      const int8_t tp = patch_type_rhs(F, ft, r);
      bool rev = false;
      if (!internal)
      {
        const bool bc0 = ft[F.fcts[0]] == EQLB_FCT_ESSNT_DUAL;
        const bool req = (tp == EQLB_PATCH_ESSNT_DUAL || tp == EQLB_PATCH_MIXED);
        if (r > 0 && req && (tp != type_prev || tp == EQLB_PATCH_MIXED) && !bc0)
          rev = true;
      }
      type_prev = tp;
      int8_t* bm = bmarkers + ((size_t)node * nrhs + r) * hzmax;
      for (int j = 0; j < hz; ++j)
        bm[j] = 0;
      if (tp == EQLB_PATCH_ESSNT_DUAL || tp == EQLB_PATCH_MIXED)
      {
        bm[0] = 1;
        const int offset_En = nc * (k - 1);
        for (int j = 1; j < k; ++j)
        {
          if (tp == EQLB_PATCH_ESSNT_DUAL)
          {
            bm[j] = 1;
            bm[j + offset_En] = 1;
          }
          else if (rev)
            bm[offset_En + j] = 1;
          else
            bm[j] = 1;
        }
      }
    }
  }
}

__global__ void cellJ_kernel(int ncell, const double* __restrict__ x, const int32_t* __restrict__ cell_node,
                             double* __restrict__ cellJ)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell)
    return;
  const int32_t n0 = cell_node[3 * c], n1 = cell_node[3 * c + 1], n2 = cell_node[3 * c + 2];
  const double x0 = x[3 * n0], y0 = x[3 * n0 + 1];
  double2* out = reinterpret_cast<double2*>(cellJ + 4 * (size_t)c);
  out[0] = make_double2(x[3 * n1] - x0, x[3 * n2] - x0);          // J00 J01
  out[1] = make_double2(x[3 * n1 + 1] - y0, x[3 * n2 + 1] - y0);  // J10 J11
}


// ---------------------------------------------------------------------------
// EV patch ordering and sub-DOF maps (ev/Patch.cpp:83-309, 360-437, 482-676).
// Parity evidence only: the EV hot kernel works on the SE fan and the conforming
// global numbering directly.  The EV convention differs from the SE one: the fan
// starts at the first facet of node->facet (interior) or at the flux / Dirichlet
// boundary facet, cells are stored 0-based and the interior patch stores the cell
// BEHIND facet ii at ii+1 (wrap to 0).
//   mixed element dofs: [3k facet][k^2-k cell flux][ndg DG]; global numbering of the
//   conforming space: facet*k + j | nfct*k + cell*(k^2-k) + i | nflux + cell*ndg + q.
// ---------------------------------------------------------------------------
__global__ void ev_dofmap_kernel(MeshView m, const int8_t* __restrict__ facet_type, int npatch, int ncmax, int k, int nrt,
                                 int ndg, const uint8_t* __restrict__ owned, int32_t* __restrict__ o_ncells,
                                 int32_t* __restrict__ o_cells, int32_t* __restrict__ o_fcts, int8_t* __restrict__ o_inod,
                                 int32_t* __restrict__ o_elmt, int32_t* __restrict__ o_patch, int32_t* __restrict__ o_global,
                                 int32_t* __restrict__ o_lpatch, int32_t* __restrict__ o_lglobal)
{
  const int node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= npatch || (owned && !owned[node]))
    return;
  const int nc = m.node_cell_off[node + 1] - m.node_cell_off[node];
  const int nf = m.node_fct_off[node + 1] - m.node_fct_off[node];
  const int32_t* pf = m.node_fct + m.node_fct_off[node];
  const int nflux_cell = nrt - 3 * k, ndof_cell = nflux_cell + ndg, nz = nrt + ndg - k;
  const int nflux = m.nfct * k + m.ncell * nflux_cell;
  const int lenf = ncmax * nflux_cell + (ncmax + 1) * k;
  int32_t* cells = o_cells + (size_t)node * ncmax;
  int32_t* fcts = o_fcts + (size_t)node * (ncmax + 1);
  int8_t* inod = o_inod + (size_t)node * ncmax;
  int32_t* d_el = o_elmt + (size_t)node * ncmax * nz;
  int32_t* d_pa = o_patch + (size_t)node * ncmax * nz;
  int32_t* d_gl = o_global + (size_t)node * ncmax * nz;
  int32_t* l_pa = o_lpatch + (size_t)node * lenf;
  int32_t* l_gl = o_lglobal + (size_t)node * lenf;
  o_ncells[node] = nc;

  // sorted facet ids of the patch (insertion sort, <= 17 entries)
  int32_t srt[EQLB_NCMAX + 1];
  for (int i = 0; i < nf; ++i)
  {
    int32_t v = pf[i];
    int q = i;
    while (q > 0 && srt[q - 1] > v)
    {
      srt[q] = srt[q - 1];
      --q;
    }
    srt[q] = v;
  }
  auto in_patch = [&](int32_t f)
  {
    for (int i = 0; i < nf; ++i)
      if (srt[i] == f)
        return true;
    return false;
  };
  // patch type of RHS 0 and first facet
  const bool bnd = nf > nc;
  int32_t fct_i = pf[0];
  if (bnd)
  {
    int32_t ef = -1, ep = -1;
    for (int i = 0; i < nf; ++i)
    {
      const int8_t ft = facet_type[pf[i]];
      if (ft == EQLB_FCT_ESSNT_PRIMAL && ep < 0)
        ep = pf[i];
      else if (ft == EQLB_FCT_ESSNT_DUAL && ef < 0)
        ef = pf[i];
    }
    fct_i = (ef < 0) ? ep : ef;
  }
  auto gdof = [&](int32_t cell, int ldof) -> int32_t
  {
    if (ldof < 3 * k)
      return m.cell_fct[3 * cell + ldof / k] * k + ldof % k;
    if (ldof < nrt)
      return m.nfct * k + cell * nflux_cell + (ldof - 3 * k);
    return nflux + cell * ndg + (ldof - nrt);
  };
  int32_t cell_i = -1, dof_patch = 0, offs_l = 0;
  for (int ii = 0; ii < nc; ++ii)
  {
    // cell behind facet ii and the local ids of that facet in both cells
    const int o = m.fct_cell_off[fct_i];
    int32_t c_new, c_old;
    if (bnd && ii == 0)
    {
      c_new = m.fct_cell[o];
      c_old = c_new;
    }
    else if (m.fct_cell[o] == cell_i)
    {
      c_new = m.fct_cell[o + 1];
      c_old = m.fct_cell[o];
    }
    else
    {
      c_new = m.fct_cell[o];
      c_old = m.fct_cell[o + 1];
    }
    const int lf_new = local_index3(m.cell_fct + 3 * c_new, fct_i);
    const int lf_old = local_index3(m.cell_fct + 3 * c_old, fct_i);
    const int8_t vloc = (int8_t)local_index3(m.cell_node + 3 * c_new, node);
    // next facet: of the two other facets of the cell the one that belongs to the patch
    const int32_t* cf = m.cell_fct + 3 * c_new;
    const int32_t fa = cf[lf_new == 0 ? 1 : 0], fb = cf[lf_new == 2 ? 1 : 2];
    const int32_t e0 = min(fa, fb), e1 = max(fa, fb);
    int32_t fct_next;
    if (e0 < srt[0])
      fct_next = e1;
    else if (e1 > srt[nf - 1])
      fct_next = e0;
    else
      fct_next = in_patch(e0) ? e0 : e1;
    // storage position of the cell and offsets of its dof block
    int32_t offs_p, offs_f;
    if (bnd)
    {
      cells[ii] = c_new;
      inod[ii] = vloc;
      offs_f = (ii == 0) ? k : (ii - 1) * nz + k;
      offs_p = ii * nz;
    }
    else if (ii < nf - 1)
    {
      cells[ii + 1] = c_new;
      cells[ii] = c_old;
      inod[ii + 1] = vloc;
      offs_f = ii * nz + k;
      offs_p = (ii + 1) * nz;
    }
    else
    {
      cells[0] = c_new;
      inod[0] = vloc;
      offs_f = ii * nz + k;
      offs_p = 0;
    }
    cell_i = c_new;
    for (int jj = 0; jj < k; ++jj)
    {
      const int ldof = lf_new * k + jj;
      const int32_t g = gdof(cell_i, ldof);
      d_el[offs_p] = ldof;
      d_el[offs_f + jj] = lf_old * k + jj;
      d_pa[offs_p] = dof_patch;
      d_pa[offs_f + jj] = dof_patch;
      d_gl[offs_p] = g;
      d_gl[offs_f + jj] = g;
      l_pa[offs_l] = dof_patch;
      l_gl[offs_l] = g;
      ++dof_patch;
      ++offs_p;
      ++offs_l;
    }
    offs_p += k;
    for (int jj = 0; jj < ndof_cell; ++jj)
    {
      const int ldof = 3 * k + jj;
      const int32_t g = gdof(cell_i, ldof);
      d_el[offs_p] = ldof;
      d_pa[offs_p] = dof_patch;
      d_gl[offs_p] = g;
      if (jj < nflux_cell)
      {
        l_pa[offs_l] = dof_patch;
        l_gl[offs_l] = g;
        ++offs_l;
      }
      ++dof_patch;
      ++offs_p;
    }
    fcts[ii] = fct_i;
    fct_i = fct_next;
  }
  if (bnd)
  {
    const int lf = local_index3(m.cell_fct + 3 * cell_i, fct_i);
    int32_t offs_p = (nc - 1) * nz + k;
    for (int jj = 0; jj < k; ++jj)
    {
      const int ldof = lf * k + jj;
      const int32_t g = gdof(cell_i, ldof);
      d_el[offs_p] = ldof;
      d_pa[offs_p] = dof_patch;
      d_gl[offs_p] = g;
      l_pa[offs_l] = dof_patch;
      l_gl[offs_l] = g;
      ++dof_patch;
      ++offs_p;
      ++offs_l;
    }
    fcts[nf - 1] = fct_i;
  }
}

} // namespace

MeshView eqlb_handle::mesh_view() const
{
  MeshView m;
  m.nnode = nnode;
  m.ncell = ncell;
  m.nfct = nfct;
  m.x = d_x.p;
  m.cell_node = d_cell_node.p;
  m.cell_fct = d_cell_fct.p;
  m.fct_node = d_fct_node.p;
  m.fct_cell_off = d_fct_cell_off.p;
  m.fct_cell = d_fct_cell.p;
  m.node_cell_off = d_node_cell_off.p;
  m.node_cell = d_node_cell.p;
  m.node_fct_off = d_node_fct_off.p;
  m.node_fct = d_node_fct.p;
  m.fct_perms = d_fct_perms.p;
  return m;
}

PatchView eqlb_handle::patch_view() const
{
  PatchView v;
  v.npatch = nactive;
  v.ncmax = ncmax;
  v.nrhs = nrhs;
  v.stride = pstride;
  v.node = d_pnode.p;
  v.ncells = d_pncells.p;
  v.cell = d_pcell.p;
  v.info = d_pinfo.p;
  v.rhsinfo = d_prhs.p;
  v.rec = d_prec.p;
  return v;
}

void launch_compute_cellJ(eqlb_handle* h)
{
  h->d_cellJ.alloc((size_t)h->ncell * 4);
  const int bs = 256;
  cellJ_kernel<<<(h->ncell + bs - 1) / bs, bs, 0, h->stream>>>(h->ncell, h->d_x.p, h->d_cell_node.p, h->d_cellJ.p);
  CUDA_CHECK(cudaGetLastError());
  h->launches++;
}

void launch_patch_builder(eqlb_handle* h, int32_t* x_ncells, int32_t* x_cells, int32_t* x_fcts, int8_t* x_inod,
                          int8_t* x_fl, int8_t* x_type, uint8_t* x_rev, uint8_t* x_reversion)
{
  const bool expand = x_ncells || x_cells || x_fcts || x_inod || x_fl || x_type || x_rev || x_reversion;
  const int bs = 128;
  if (h->nactive == 0)
    return;
  if (h->d_order.n < (size_t)h->nactive)
    throw EqlbError(EQLB_ERR_STATE, "patch builder: launch order not available");
  patch_builder_kernel<<<(h->nactive + bs - 1) / bs, bs, 0, h->stream>>>(
      h->mesh_view(), h->d_facet_type.p, h->nrhs, h->d_order.p, h->nactive, h->pstride, h->ncmax,
      expand ? nullptr : h->d_pnode.p, h->d_pncells.p, expand ? nullptr : h->d_pcell.p, h->d_pinfo.p,
      expand ? nullptr : h->d_prhs.p, expand ? nullptr : h->d_prec.p, h->d_seginfo.p, h->nsub, x_ncells, x_cells, x_fcts, x_inod, x_fl, x_type, x_rev, x_reversion);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->launches++;
}

void launch_ev_dofmaps(eqlb_handle* h, int32_t* d_ncells, int32_t* d_cells, int32_t* d_fcts, int8_t* d_inod, int32_t* d_elmt,
                       int32_t* d_patch, int32_t* d_global, int32_t* d_lpatch, int32_t* d_lglobal)
{
  DevBuf<uint8_t> d_owned;
  d_owned.upload(h->h_owned.data(), h->h_owned.size());
  const int bs = 128;
  ev_dofmap_kernel<<<(h->nnode + bs - 1) / bs, bs, 0, h->stream>>>(h->mesh_view(), h->d_facet_type.p, h->nnode, h->ncmax, h->k,
                                                                  h->nrt, h->ndg, d_owned.p, d_ncells, d_cells, d_fcts, d_inod,
                                                                  d_elmt, d_patch, d_global, d_lpatch, d_lglobal);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->launches++;
}

void launch_se_dofmaps(eqlb_handle* h, int32_t* d_dofmap, int32_t* d_projflux, int8_t* d_bmarkers, int ndpc, int hzmax)
{
  DevBuf<int32_t> d_closure;
  std::vector<int32_t> closure(3 * h->ndg_fct);
  // closure dofs of P_p facets: vertices (a,b) of facet f then edge interior dofs
  const int fv[3][2] = {{1, 2}, {0, 2}, {0, 1}};
  for (int f = 0; f < 3; ++f)
  {
    if (h->p == 0)
      closure[f] = 0;
    else
    {
      closure[f * h->ndg_fct] = fv[f][0];
      closure[f * h->ndg_fct + 1] = fv[f][1];
      for (int i = 0; i < h->p - 1; ++i)
        closure[f * h->ndg_fct + 2 + i] = 3 + f * (h->p - 1) + i;
    }
  }
  d_closure.upload(closure.data(), closure.size());
  DevBuf<uint8_t> d_owned;
  d_owned.upload(h->h_owned.data(), h->h_owned.size());
  const int bs = 128;
  se_dofmap_kernel<<<(h->nnode + bs - 1) / bs, bs, 0, h->stream>>>(
      h->mesh_view(), h->d_facet_type.p, h->nrhs, h->nnode, h->ncmax, h->k, h->nrt, h->nadd, h->ndiv, h->ndg_fct, h->p,
      (h->flags & EQLB_FLAG_STRESS) != 0, d_closure.p, h->d_cellJ.p, d_dofmap, d_projflux, d_bmarkers, ndpc, hzmax,
      d_owned.p);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->launches++;
}


// ---------------------------------------------------------------------------------------------------
// Greedy vertex colouring on the device, EXACTLY the sequential first-fit colouring in vertex order
// (two vertices of a cell never share a colour): vertex z takes the smallest colour not used by the
// lower-numbered vertices of its cells.  Dataflow formulation: every vertex counts its lower-numbered
// neighbours (with the multiplicity of shared cells); a vertex whose count drops to zero is appended to a
// ready queue; persistent threads pop vertices from the queue, colour them and decrement the counts of their
// higher-numbered neighbours.  The parallelism is the width of the dependency graph (a whole anti-diagonal of
// a row-numbered structured mesh), the run time its depth times a queue round trip - not limited by how many
// vertices fit on the device at once.  A popped vertex always finds all its lower neighbours decided, so
// the result does not depend on the schedule.  colour: -1 = no colour (vertex not owned / grouped).
// Skipped vertices take part in the dataflow (they release their neighbours) but get no colour.
// ---------------------------------------------------------------------------------------------------
namespace
{
__global__ void fill_int_kernel(int* p, int n, int v)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    p[i] = v;
}

// pending[z] = number of (cell, vertex) pairs around z with a lower-numbered vertex; ready vertices -> queue
__global__ void colour_count_kernel(int n, const int32_t* __restrict__ node_cell_off, const int32_t* __restrict__ node_cell,
                                    const int32_t* __restrict__ cell_node, int* pending, int* queue, unsigned* ctl)
{
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= n)
    return;
  int cnt = 0;
  for (int i = node_cell_off[z]; i < node_cell_off[z + 1]; ++i)
  {
    const int32_t* cn = cell_node + 3 * (size_t)node_cell[i];
    for (int j = 0; j < 3; ++j)
      cnt += (cn[j] < z) ? 1 : 0;
  }
  pending[z] = cnt;
  if (cnt == 0)
    queue[atomicAdd(&ctl[3], 1u)] = z;  // ctl[3]: queue tail
}

// Variant for meshes whose dependency wavefront fits the device: every thread owns ONE vertex (handed out in index
// order by a ticket counter, so whatever a thread waits for is resident or finished) and retries until its lower
// neighbours are decided.  No queue, no counters: 3.7 us per wavefront step instead of ~25 us - but only the
// resident vertices (148 SMs x 2048 threads) are in flight, which on a row-numbered n x n mesh limits the wavefront
// to 300 k / n rows: 13 ms at 1024^2, an estimated 0.8 s at 4096^2, where the dataflow kernel takes 50 ms.
__global__ void __launch_bounds__(256)
greedy_colour_ticket_kernel(int n, const int32_t* __restrict__ node_cell_off, const int32_t* __restrict__ node_cell,
                     const int32_t* __restrict__ cell_node, const uint8_t* __restrict__ skip, int* colour, unsigned* ticket,
                     int* status, unsigned* maxcol)
{
  __shared__ unsigned s_base;
  if (threadIdx.x == 0)
    s_base = atomicAdd(ticket, 1u);
  __syncthreads();
  const long zl = (long)s_base * blockDim.x + threadIdx.x;
  if (zl >= n)
    return;
  const int z = (int)zl;
  volatile int* vc = colour;
  if (skip && skip[z])
  {
    vc[z] = -1;
    return;
  }
  const int c0 = node_cell_off[z], c1 = node_cell_off[z + 1];
  unsigned long long used = 0ull;
  // lower-numbered neighbours still undecided: only those are re-read (one bit per (cell slot, local vertex)
  // for patches with up to 21 cells, beyond that every round rescans the whole patch)
  const bool track = (c1 - c0) <= 21;
  unsigned long long pending = ~0ull;
  long spins = 0;
  // The decision is stored INSIDE the loop: lanes that are done must publish their colour before they park at
  // the reconvergence point behind the loop, where they wait for the lanes of the warp that still spin on them.
  bool done = false;
  while (!done)
  {
    unsigned long long still = 0ull;
    for (int i = c0; i < c1; ++i)
    {
      if (track && !((pending >> (3 * (i - c0))) & 7ull))
        continue;
      const int32_t* cn = cell_node + 3 * (size_t)node_cell[i];
      for (int j = 0; j < 3; ++j)
      {
        const int w = cn[j];
        if (w >= z || (track && !((pending >> (3 * (i - c0) + j)) & 1ull)))
          continue;
        const int c = vc[w];
        if (c == -2)
          still |= track ? (1ull << (3 * (i - c0) + j)) : 1ull;
        else if (c >= 0)
          used |= 1ull << c;
      }
    }
    pending = still;
    if (!pending)
    {
      const int c = __ffsll((long long)~used) - 1;
      vc[z] = c;
      __threadfence();
      if ((unsigned)(c + 1) > *(volatile unsigned*)maxcol)
        atomicMax(maxcol, (unsigned)(c + 1));
      done = true;
    }
    else if (++spins > (1L << 22))
    {
      atomicExch(status, 1);  // no progress: reported by the host, never silent
      done = true;
    }
    else
      __nanosleep(32);
  }
}

// ctl: [0] queue head, [1] status, [2] max colour + 1, [3] queue tail
__global__ void __launch_bounds__(256)
greedy_colour_kernel(int n, const int32_t* __restrict__ node_cell_off, const int32_t* __restrict__ node_cell,
                     const int32_t* __restrict__ cell_node, const uint8_t* __restrict__ skip, int* colour, int* pending,
                     int* queue, unsigned* ctl)
{
  volatile int* vq = queue;
  volatile int* vc = colour;
  volatile unsigned* vctl = ctl;
  while (true)
  {
    const unsigned slot = atomicAdd(&ctl[0], 1u);
    if (slot >= (unsigned)n)
      return;
    // the slot is filled by whoever releases the slot-th ready vertex; all poppers are resident (persistent grid)
    // (claiming 32 slots per warp only when the queue is non-empty was measured slower: 108 vs 71 ms at 1024^2)
    int z = vq[slot];
    long spins = 0;
    while (z < 0)
    {
      if (++spins > (1L << 24))
      {
        atomicExch(&ctl[1], 1u);  // no progress: reported by the host, never silent
        return;
      }
      __nanosleep(32);
      z = vq[slot];
    }
    __threadfence();  // the colours published before the release of z are visible
    const int c0 = node_cell_off[z], c1 = node_cell_off[z + 1];
    if (!(skip && skip[z]))
    {
      unsigned long long used = 0ull;
      for (int i = c0; i < c1; ++i)
      {
        const int32_t* cn = cell_node + 3 * (size_t)node_cell[i];
        for (int j = 0; j < 3; ++j)
        {
          const int w = cn[j];
          if (w < z)
          {
            const int c = vc[w];
            if (c >= 0)
              used |= 1ull << c;
          }
        }
      }
      const int c = __ffsll((long long)~used) - 1;
      vc[z] = c;
      if ((unsigned)(c + 1) > vctl[2])
        atomicMax(&ctl[2], (unsigned)(c + 1));
    }
    __threadfence();  // colour before the releases
    for (int i = c0; i < c1; ++i)
    {
      const int32_t* cn = cell_node + 3 * (size_t)node_cell[i];
      for (int j = 0; j < 3; ++j)
      {
        const int u = cn[j];
        if (u > z && atomicSub(&pending[u], 1) == 1)
        {
          // last release of u: the fence orders the colour stores of all earlier releasers (each fenced before its
          // decrement of the same counter) before the queue entry
          __threadfence();
          const unsigned t = atomicAdd(&ctl[3], 1u);
          vq[t] = u;
        }
      }
    }
  }
}
} // namespace

// Colours of all vertices (-1 for skipped vertices).  Two halves so that the caller can overlap the kernel with
// other work (eqlb_create uploads the rest of the mesh meanwhile): `start` queues the kernel on the handle's
// stream, `finish` fetches the result into the host vector and returns the number of colours.
struct ColouringJob
{
  DevBuf<int> d_col, d_pending, d_queue;
  DevBuf<unsigned> d_ctl;  // [queue head, status, max colour + 1, queue tail]
  DevBuf<uint8_t> d_skip;
  bool ticket = false;
};

std::shared_ptr<ColouringJob> device_greedy_colouring_start(eqlb_handle* h, const uint8_t* h_skip)
{
  auto job = std::make_shared<ColouringJob>();
  const int n = h->nnode;
  if (n == 0)
    return job;
  job->d_col.alloc(n);
  job->d_ctl.alloc(4);
  job->d_ctl.zero(h->stream);
  if (h_skip)
  {
    job->d_skip.alloc(n);
    CUDA_CHECK(cudaMemcpyAsync(job->d_skip.p, h_skip, n, cudaMemcpyHostToDevice, h->stream));
  }
  const int nb = (n + 255) / 256;
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
  // ticket kernel while the wavefront of a row-numbered mesh fits the resident threads (~n^1.5 / resident steps),
  // dataflow kernel beyond; EQLB_COLOURING=ticket|dataflow overrides.  Both give the sequential first-fit colouring.
  static const char* force = getenv("EQLB_COLOURING");
  const double steps_ticket = (double)n * std::sqrt((double)n) / ((double)nsm * 2048.0);
  const bool ticket = force ? (force[0] == 't') : steps_ticket < 1.5e5;  // measured crossover on B200: between 2048^2 (ticket 183 vs 215 ms per eqlb_create) and 3072^2 (437 vs 409 ms)
  if (ticket)
  {
    job->ticket = true;
    fill_int_kernel<<<nb, 256, 0, h->stream>>>(job->d_col.p, n, -2);
    greedy_colour_ticket_kernel<<<nb, 256, 0, h->stream>>>(n, h->d_node_cell_off.p, h->d_node_cell.p, h->d_cell_node.p,
                                                            h_skip ? job->d_skip.p : nullptr, job->d_col.p, job->d_ctl.p,
                                                            reinterpret_cast<int*>(job->d_ctl.p + 1), job->d_ctl.p + 2);
    CUDA_CHECK(cudaGetLastError());
    return job;
  }
  job->d_pending.alloc(n);
  job->d_queue.alloc(n);
  fill_int_kernel<<<nb, 256, 0, h->stream>>>(job->d_col.p, n, -1);
  fill_int_kernel<<<nb, 256, 0, h->stream>>>(job->d_queue.p, n, -1);
  colour_count_kernel<<<nb, 256, 0, h->stream>>>(n, h->d_node_cell_off.p, h->d_node_cell.p, h->d_cell_node.p, job->d_pending.p,
                                                 job->d_queue.p, job->d_ctl.p);
  CUDA_CHECK(cudaGetLastError());
  // persistent grid: every CTA must be resident (the poppers wait for each other's releases)
  int per_sm = 1;
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, greedy_colour_kernel, 256, 0));
  const int grid = std::max(1, std::min(nb, nsm * std::max(1, std::min(per_sm, 4))));
  greedy_colour_kernel<<<grid, 256, 0, h->stream>>>(n, h->d_node_cell_off.p, h->d_node_cell.p, h->d_cell_node.p,
                                                    h_skip ? job->d_skip.p : nullptr, job->d_col.p, job->d_pending.p,
                                                    job->d_queue.p, job->d_ctl.p);
  CUDA_CHECK(cudaGetLastError());
  return job;
}

const int* colouring_job_colours(const ColouringJob& job) { return job.d_col.p; }

int device_greedy_colouring_finish(eqlb_handle* h, ColouringJob& job, std::vector<int32_t>& colour)
{
  const int n = h->nnode;
  colour.resize(n);
  if (n == 0)
    return 0;
  unsigned ctl[4] = {0, 0, 0, 0};
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  CUDA_CHECK(cudaMemcpy(ctl, job.d_ctl.p, sizeof(ctl), cudaMemcpyDeviceToHost));
  if (ctl[1] != 0 || (!job.ticket && ctl[3] != (unsigned)n))
    throw EqlbError(EQLB_ERR_CUDA, "device colouring made no progress (EQLB_HOST_COLOURING=1 selects the host algorithm)");
  eqlb_d2h(colour.data(), job.d_col.p, (size_t)n * sizeof(int));
  return (int)ctl[2];
}

// ---------------------------------------------------------------------------------------------------
// Launch order on the device: counting sort of the active patches by (segment, lane class), vertex order
// within a class (stable radix sort), and the result slabs of the host pipeline.  Raw key of a patch:
// ((chunk * 64 + colour) << 2) | lane class  (< 4096: at most 16 chunks, 64 colours), 0xFFFF for patches that
// are not launched (not owned / grouped).  The host turns the 4096-bin histogram into segment offsets, merges
// small lane classes and sends back the rank of every raw key in launch order.
// ---------------------------------------------------------------------------------------------------
namespace
{
constexpr int ORDER_BINS = 4096;

__global__ void __launch_bounds__(256)
order_rawkey_kernel(int n, const int32_t* __restrict__ node_cell_off, const int32_t* __restrict__ node_cell,
                    const int32_t* __restrict__ node_fct_off, const int* __restrict__ colour, int nchunk, int ncell, int nrhs,
                    uint16_t* __restrict__ rawkey, unsigned* __restrict__ hist)
{
  __shared__ unsigned s_hist[ORDER_BINS];
  for (int i = threadIdx.x; i < ORDER_BINS; i += blockDim.x)
    s_hist[i] = 0u;
  __syncthreads();
  const long per = ((long)n + gridDim.x - 1) / gridDim.x;
  const long z0 = per * blockIdx.x, z1 = min((long)n, z0 + per);
  for (long z = z0 + threadIdx.x; z < z1; z += blockDim.x)
  {
    const int col = colour[z];
    unsigned key = 0xFFFFu;
    if (col >= 0)
    {
      const int c0 = node_cell_off[z], c1 = node_cell_off[z + 1];
      int cmax = 0;
      for (int i = c0; i < c1; ++i)
        cmax = max(cmax, node_cell[i]);
      const int chunk = (int)((long)cmax * nchunk / max(ncell, 1));
      const int nf = node_fct_off[z + 1] - node_fct_off[z], nc = c1 - c0;
      const int cls = (nf > 16 || !(nrhs == 1 || nf == nc)) ? 3 : (nf <= 4 ? 0 : (nf <= 8 ? 1 : 2));
      key = (unsigned)(((chunk * 64 + col) << 2) | cls);
      atomicAdd(&s_hist[key], 1u);
    }
    rawkey[z] = (uint16_t)key;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ORDER_BINS; i += blockDim.x)
    if (s_hist[i])
      atomicAdd(&hist[i], s_hist[i]);
}

__global__ void order_sortkey_kernel(int n, const uint16_t* __restrict__ rawkey, const uint16_t* __restrict__ rank,
                                     uint16_t* __restrict__ sortkey, int32_t* __restrict__ iota)
{
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= n)
    return;
  const uint16_t r = rawkey[z];
  sortkey[z] = (r == 0xFFFFu) ? (uint16_t)0xFFFFu : rank[r];
  iota[z] = z;
}

// last stage that adds into the cells / facets of every result slab (slab s of `cnt` items: [cnt s / nchunk, cnt (s+1) / nchunk))
__global__ void __launch_bounds__(256)
slab_stage_kernel(int ncell, int nfct, const int32_t* __restrict__ cell_node, const int32_t* __restrict__ fct_node,
                  const uint16_t* __restrict__ rawkey, int nchunk, int* __restrict__ cfin, int* __restrict__ ffin)
{
  __shared__ int s_c[16], s_f[16];
  if (threadIdx.x < 16)
    s_c[threadIdx.x] = s_f[threadIdx.x] = 0;
  __syncthreads();
  auto stage = [&](int v)
  {
    const unsigned r = rawkey[v];
    return r == 0xFFFFu ? -1 : (int)(r >> 8);
  };
  auto slab = [&](long i, long cnt)
  {
    int s0 = (int)(i * nchunk / max(cnt, 1L));
    while (s0 + 1 < nchunk && cnt * (s0 + 1) / nchunk <= i)
      ++s0;
    while (s0 > 0 && cnt * s0 / nchunk > i)
      --s0;
    return s0;
  };
  const long stride = (long)gridDim.x * blockDim.x, t0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long c = t0; c < ncell; c += stride)
  {
    const int st = max(stage(cell_node[3 * c]), max(stage(cell_node[3 * c + 1]), stage(cell_node[3 * c + 2])));
    if (st > 0)
      atomicMax(&s_c[slab(c, ncell)], st);
  }
  for (long f = t0; f < nfct; f += stride)
  {
    const int st = max(stage(fct_node[2 * f]), stage(fct_node[2 * f + 1]));
    if (st > 0)
      atomicMax(&s_f[slab(f, nfct)], st);
  }
  __syncthreads();
  if (threadIdx.x < nchunk)
  {
    if (s_c[threadIdx.x] > 0)
      atomicMax(&cfin[threadIdx.x], s_c[threadIdx.x]);
    if (s_f[threadIdx.x] > 0)
      atomicMax(&ffin[threadIdx.x], s_f[threadIdx.x]);
  }
}
} // namespace

// step 1: raw keys + histogram (host vector of ORDER_BINS counts); the raw keys stay in h->d_rawkey
void device_order_histogram(eqlb_handle* h, const int* d_colour, int nchunk, std::vector<uint32_t>& hist)
{
  const int n = h->nnode;
  hist.assign(ORDER_BINS, 0u);
  if (n == 0)
    return;
  if (nchunk > 16)
    throw EqlbError(EQLB_ERR_STATE, "device launch order: more than 16 chunks");
  h->d_rawkey.alloc(n);
  DevBuf<unsigned> d_hist;
  d_hist.alloc(ORDER_BINS);
  d_hist.zero(h->stream);
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
  const int grid = std::max(1, std::min((n + 255) / 256, nsm * 4));
  order_rawkey_kernel<<<grid, 256, 0, h->stream>>>(n, h->d_node_cell_off.p, h->d_node_cell.p, h->d_node_fct_off.p, d_colour, nchunk,
                                                   h->ncell, h->nrhs, h->d_rawkey.p, d_hist.p);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpyAsync(hist.data(), d_hist.p, ORDER_BINS * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

// step 2: rank[raw key] = position of its (segment, merged class) in launch order -> h->d_order[offset ...)
void device_order_sort(eqlb_handle* h, const std::vector<uint16_t>& rank, int offset, int count)
{
  const int n = h->nnode;
  h->d_order.alloc((size_t)std::max(h->nactive, 1));
  if (n == 0 || count == 0)
    return;
  DevBuf<uint16_t> d_rank, d_key_in, d_key_out;
  DevBuf<int32_t> d_val_in, d_val_out;
  d_rank.alloc(ORDER_BINS);
  CUDA_CHECK(cudaMemcpyAsync(d_rank.p, rank.data(), ORDER_BINS * sizeof(uint16_t), cudaMemcpyHostToDevice, h->stream));
  d_key_in.alloc(n);
  d_key_out.alloc(n);
  d_val_in.alloc(n);
  d_val_out.alloc(n);
  order_sortkey_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(n, h->d_rawkey.p, d_rank.p, d_key_in.p, d_val_in.p);
  CUDA_CHECK(cudaGetLastError());
  size_t tmp_bytes = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_key_in.p, d_key_out.p, d_val_in.p, d_val_out.p, n, 0, 16, h->stream));
  DevBuf<uint8_t> d_tmp;
  d_tmp.alloc(std::max<size_t>(tmp_bytes, 1));
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_key_in.p, d_key_out.p, d_val_in.p, d_val_out.p, n, 0, 16, h->stream));
  // the launched patches come first (skipped ones carry the key 0xFFFF)
  CUDA_CHECK(cudaMemcpyAsync(h->d_order.p + offset, d_val_out.p, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

// result slabs of the host pipeline from the raw keys (needs the facet connectivity on the device)
void device_slab_stages(eqlb_handle* h, int nchunk, std::vector<int>& cfin, std::vector<int>& ffin)
{
  cfin.assign(nchunk, 0);
  ffin.assign(nchunk, 0);
  if (h->nnode == 0 || nchunk > 16)
    return;
  DevBuf<int> d_fin;
  d_fin.alloc(32);
  d_fin.zero(h->stream);
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
  slab_stage_kernel<<<nsm * 4, 256, 0, h->stream>>>(h->ncell, h->nfct, h->d_cell_node.p, h->d_fct_node.p, h->d_rawkey.p, nchunk,
                                                   d_fin.p, d_fin.p + 16);
  CUDA_CHECK(cudaGetLastError());
  int fin[32];
  CUDA_CHECK(cudaMemcpyAsync(fin, d_fin.p, sizeof(fin), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  for (int s2 = 0; s2 < nchunk; ++s2)
  {
    cfin[s2] = fin[s2];
    ffin[s2] = fin[16 + s2];
  }
}
