// TEST INFRASTRUCTURE - NOT PRODUCT CODE (see oracle_common.hpp).
// Restatement of se/solve_patch_semiexplt.hpp, se/assembly.hpp,
// se/fluxmin_kernel.hpp, se/PatchData.hpp, se/reconstruction.hpp and the
// patch-BC part of base/BoundaryData.cpp.  Serial node loop, quadrature-loop
// cell kernels with per-call Piola mapping, dense per-patch LLT, += scatter.
#include "oracle_se.hpp"

namespace oracle
{

// ---------------------------------------------------------------------------
// se::PatchData (scratch)  se/PatchData.hpp:37-223
// ---------------------------------------------------------------------------
struct PatchData
{
  const eqlb_tables* t;
  const int k, nrt, nqf, nrhs;
  const bool symconstr;
  int ncells = 0;
  int dim_hdivz = 0, dim_constr = 0;
  bool meanvalue_condition_required = false;
  std::vector<double> J, K, detJ, prefactors, Mm, coeff_flux, coeff_stress, jumpG_Eam1, c_ta_div, cj_ta_ea;
  std::vector<uint8_t> reversed;
  std::vector<double> coefficients_f, coefficients_G_Ta, coefficients_G_Tap1;
  std::vector<int8_t> boundary_markers;
  // equation system
  int ld = 0; // leading dimension of A
  std::vector<double> A, A_rec, L, u_sigma, Te;
  // constrained system
  std::vector<double> B, Ainv_t_B, C, u_c, Be, Ce, Le;
  int ldB = 0, ldC = 0;
  int dim_hdivz_per_cell;

  PatchData(const Patch& patch, bool symconstr_)
      : t(patch.pb.t), k(patch.k), nrt(patch.nrt), nqf(patch.pb.t->nqf), nrhs(patch.nrhs), symconstr(symconstr_)
  {
    const int ncm = patch.ncells_max;
    J.assign(4 * ncm, 0);
    K.assign(4 * ncm, 0);
    detJ.assign(ncm, 0);
    prefactors.assign(2 * ncm, 0);
    reversed.assign(2 * ncm, 0);
    Mm.assign((size_t)ncm * (k + 1) * 2 * nqf, 0);
    coefficients_f.assign(t->ndg, 0);
    coefficients_G_Ta.assign(2 * t->ndg, 0);
    coefficients_G_Tap1.assign(2 * t->ndg, 0);
    coeff_flux.assign((size_t)nrhs * ncm * nrt, 0);
    jumpG_Eam1.assign((size_t)nqf * 4, 0);
    c_ta_div.assign(std::max(t->ndiv, 2), 0);
    cj_ta_ea.assign(std::max(k - 1, 1), 0);
    dim_hdivz_per_cell = 2 * k + t->nadd - 1;
    const int nfm = ncm + 1;
    int hzmax = 1 + (k - 1) * nfm + t->nadd * ncm;
    if (patch.groupsize_max != 1)
      hzmax += patch.groupsize_max * (k - 1);
    ld = hzmax;
    A.assign((size_t)hzmax * hzmax, 0);
    const int dpc = dim_hdivz_per_cell + k - 1;
    Te.assign((size_t)(dpc + 1) * dpc, 0);
    if (symconstr)
    {
      const int npnts_max = nfm + 1;
      boundary_markers.assign(2 * hzmax, 0);
      coeff_stress.assign((size_t)ncm * 2 * nrt, 0);
      L.assign(2 * hzmax + npnts_max + 1, 0);
      u_sigma.assign(2 * hzmax, 0);
      A_rec.assign((size_t)hzmax * hzmax, 0);
      ldB = 2 * npnts_max;
      B.assign((size_t)hzmax * ldB, 0);
      Ainv_t_B.assign((size_t)hzmax * npnts_max, 0);
      ldC = npnts_max + 1;
      C.assign((size_t)ldC * ldC, 0);
      u_c.assign(ldC, 0);
      Be.assign((size_t)dpc * 2 * 3, 0);
      Ce.assign(3, 0);
      Le.assign(2 * dpc + 3, 0);
    }
    else
    {
      boundary_markers.assign(hzmax, 0);
      L.assign(hzmax, 0);
      u_sigma.assign(hzmax, 0);
    }
  }

  // se/PatchData.hpp:168-223
  void reinitialisation(const std::vector<int8_t>& type_patch, int nc)
  {
    ncells = nc;
    const int add = t->nadd;
    if (type_patch[0] == internal)
    {
      dim_hdivz = 1 + (k - 1) * nc + add * nc;
      dim_constr = nc + 1;
      meanvalue_condition_required = true;
    }
    else
    {
      dim_hdivz = 1 + (k - 1) * (nc + 1) + add * nc;
      dim_constr = nc + 2;
      meanvalue_condition_required = true;
      int count = 0;
      for (int i = 0; i < 2 && i < (int)type_patch.size(); ++i)
        if (type_patch[i] == bound_essnt_primal || type_patch[i] == bound_mixed)
          ++count;
      if (count > 0)
        meanvalue_condition_required = false;
    }
    std::fill_n(reversed.begin(), 2 * nc, 0);
    std::fill_n(coeff_flux.begin(), (size_t)nrhs * nc * nrt, 0.0);
    std::fill(jumpG_Eam1.begin(), jumpG_Eam1.end(), 0.0);
  }

  double* coefficients_flux(int i) { return coeff_flux.data() + (size_t)i * ncells * nrt; }
  double& M_mapped(int c, int i, int d, int n) { return Mm[(((size_t)c * (k + 1) + i) * 2 + d) * nqf + n]; }
  double& GtHat_Eam1(int n, int s, int d) { return jumpG_Eam1[(n * 2 + s) * 2 + d]; }
  double* coefficients_stress(int i, int a) { return coeff_stress.data() + ((size_t)(a - 1) * 2 + i) * nrt; }
};

// ---------------------------------------------------------------------------
// Boundary data: base/BoundaryData.cpp:686-745 + interpolate_flux :171-250
// ---------------------------------------------------------------------------
static void calculate_patch_bc(Problem& pb, int rhs_i, int32_t fct, int8_t hat_id, const double* J, double detJ,
                               const double* K)
{
  if (pb.bfct_type(rhs_i, fct) != essnt_dual)
    return;
  const eqlb_tables* t = pb.t;
  const int k = t->k, nrt = t->nrt, nqf = t->nqf;
  const int32_t cell = pb.mv.fct_to_cell(fct)[0];
  const int8_t fct_loc = pb.local_fct_id[fct];
  const int offs_dofs = cell * nrt + fct_loc * k;
  std::vector<double>& bvals = pb.boundary_values[rhs_i];
  const double* x_bfunc = pb.bflux[rhs_i];

  std::vector<double> b(k);
  int ndofs_zero = 0;
  for (int i = 0; i < k; ++i)
  {
    b[i] = x_bfunc ? x_bfunc[offs_dofs + i] : 0.0;
    if (std::fabs(b[i]) < 1e-7)
      ++ndofs_zero;
  }
  if (ndofs_zero >= k)
    return;

  // interpolate_flux: push forward basis, evaluate flux x hat at the facet
  // interpolation points, pull back, apply M
  const int offs_ipnt = fct_loc * nqf;
  const int offs_dof = fct_loc * k;
  std::vector<double> flux(2 * nqf), mflux(2 * nqf);
  for (int ip = 0; ip < nqf; ++ip)
  {
    const double* phi = t->rt_f + (size_t)(offs_ipnt + ip) * nrt * 2;
    double acc0 = 0, acc1 = 0;
    for (int i = 0; i < k; ++i)
    {
      const double r0 = phi[(offs_dof + i) * 2], r1 = phi[(offs_dof + i) * 2 + 1];
      const double m0 = (J[0] * r0 + J[1] * r1) / detJ;
      const double m1 = (J[2] * r0 + J[3] * r1) / detJ;
      acc0 += b[i] * m0;
      acc1 += b[i] * m1;
    }
    const double hat = t->hat_f[(offs_ipnt + ip) * 3 + hat_id];
    flux[2 * ip] = acc0 * hat;
    flux[2 * ip + 1] = acc1 * hat;
  }
  // pull back: detJ * K * v  (_pull_back_flux(.., K, 1/detJ, J))
  for (int ip = 0; ip < nqf; ++ip)
  {
    mflux[2 * ip] = detJ * (K[0] * flux[2 * ip] + K[1] * flux[2 * ip + 1]);
    mflux[2 * ip + 1] = detJ * (K[2] * flux[2 * ip] + K[3] * flux[2 * ip + 1]);
  }
  for (int i = 0; i < k; ++i)
  {
    double dof = 0;
    for (int j = 0; j < nqf; ++j)
      for (int d = 0; d < 2; ++d)
        dof += t->M[(((size_t)fct_loc * k + i) * 2 + d) * nqf + j] * mflux[2 * j + d];
    bvals[offs_dofs + i] = dof;
  }
}

// ---------------------------------------------------------------------------
// se/assembly.hpp:46-98
// ---------------------------------------------------------------------------
static void set_boundary_markers(int8_t* bm, int nbm, const std::vector<int8_t>& types, const std::vector<bool>& revs,
                                 int ncells, int ndofs_hdivz, int k)
{
  std::fill(bm, bm + nbm, 0);
  const int offset_En = ncells * (k - 1);
  for (size_t i = 0; i < types.size(); ++i)
  {
    if (types[i] != bound_essnt_primal)
    {
      int offset_i = (int)i * ndofs_hdivz;
      bm[offset_i] = 1;
      for (int j = 1; j < k; ++j)
      {
        if (types[i] == bound_essnt_dual)
        {
          bm[offset_i + j] = 1;
          bm[offset_i + j + offset_En] = 1;
        }
        else
        {
          if (revs[i])
            bm[offset_i + offset_En + j] = 1;
          else
            bm[offset_i + j] = 1;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// se/fluxmin_kernel.hpp:60-190 (+ se/KernelData.cpp:201-221)
// Te: (ndpc+1) x ndpc row-major with leading dimension ldTe
// ---------------------------------------------------------------------------
static void minimisation_kernel(const eqlb_tables* t, bool asmbl_systmtrx, double* Te, int ldTe, const double* coefficients,
                                const Patch& patch, int a, uint8_t fct_eam1_reversed, double detJ, const double* J,
                                std::vector<double>& phi)
{
  const int k = t->k, nrt = t->nrt, nq = t->nq;
  const int nz = 2 * k + t->nadd - 1;
  const int index_load = nz;
  // Piola map of ALL basis functions at ALL quadrature points
  phi.resize((size_t)nq * nrt * 2);
  const double inv_detJ = 1.0 / detJ;
  for (int iq = 0; iq < nq; ++iq)
    for (int j = 0; j < nrt; ++j)
    {
      const double r0 = t->rt_q[((size_t)iq * nrt + j) * 2], r1 = t->rt_q[((size_t)iq * nrt + j) * 2 + 1];
      phi[((size_t)iq * nrt + j) * 2] = inv_detJ * J[0] * r0 + inv_detJ * J[1] * r1;
      phi[((size_t)iq * nrt + j) * 2 + 1] = inv_detJ * J[2] * r0 + inv_detJ * J[3] * r1;
    }
  auto PHI = [&](int iq, int i, int d) -> double& { return phi[((size_t)iq * nrt + i) * 2 + d]; };
  auto info = [&](int pl, int i) { return patch.dofmap(pl, a, i); };

  std::vector<double> gphi(2 * k, 0.0);
  const int ld0_Eam1 = info(0, 0), ld0_Ea = info(0, k);
  const int p_Eam1 = info(3, 0), p_Ea = info(3, k);

  for (int iq = 0; iq < nq; ++iq)
  {
    double sig0 = 0, sig1 = 0;
    for (int i = 0; i < nrt; ++i)
    {
      sig0 += coefficients[i] * PHI(iq, i, 0);
      sig1 += coefficients[i] * PHI(iq, i, 1);
    }
    if (fct_eam1_reversed)
    {
      for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j)
        {
          int ldj = info(0, j);
          gphi[2 * i] += t->trafo[i * k + j] * PHI(iq, ldj, 0);
          gphi[2 * i + 1] += t->trafo[i * k + j] * PHI(iq, ldj, 1);
        }
      for (int i = 0; i < k; ++i)
      {
        int ldi = info(0, i);
        PHI(iq, ldi, 0) = gphi[2 * i];
        PHI(iq, ldi, 1) = gphi[2 * i + 1];
      }
      std::fill(gphi.begin(), gphi.end(), 0.0);
    }
    PHI(iq, ld0_Ea, 0) = p_Ea * (p_Eam1 * PHI(iq, ld0_Eam1, 0) + p_Ea * PHI(iq, ld0_Ea, 0));
    PHI(iq, ld0_Ea, 1) = p_Ea * (p_Eam1 * PHI(iq, ld0_Eam1, 1) + p_Ea * PHI(iq, ld0_Ea, 1));

    const double dvol = t->qwts[iq] * std::fabs(detJ);
    for (int i = 0; i < nz; ++i)
    {
      const int ip1 = i + 1;
      const double alpha = info(3, ip1) * dvol;
      const double phi_i0 = PHI(iq, info(0, ip1), 0) * alpha;
      const double phi_i1 = PHI(iq, info(0, ip1), 1) * alpha;
      Te[index_load * ldTe + i] -= phi_i0 * sig0 + phi_i1 * sig1;
      if (asmbl_systmtrx)
        for (int j = i; j < nz; ++j)
        {
          const int jp1 = j + 1;
          const double phi_j0 = PHI(iq, info(0, jp1), 0) * info(3, jp1);
          const double phi_j1 = PHI(iq, info(0, jp1), 1) * info(3, jp1);
          Te[i * ldTe + j] += phi_i0 * phi_j0 + phi_i1 * phi_j1;
        }
    }
  }
  if (asmbl_systmtrx)
    for (int i = 1; i < nz; ++i)
      for (int j = 0; j < i; ++j)
        Te[i * ldTe + j] = Te[j * ldTe + i];
}

// se/assembly.hpp:119-274
static void assemble_fluxminimiser(bool asmbl_systmtrx, const Patch& patch, PatchData& pd, int i_rhs, bool requires_flux_bc,
                                   std::vector<double>& phi_scratch)
{
  const eqlb_tables* t = pd.t;
  const int k = t->k, nrt = t->nrt;
  const int ncells = pd.ncells, ld = pd.ld;
  const int nz = pd.dim_hdivz_per_cell;
  const int ldTe = nz + k - 1;
  double* A = pd.A.data();
  double* L = pd.L.data();
  const int8_t* bm = pd.boundary_markers.data();
  if (asmbl_systmtrx)
    std::fill(pd.A.begin(), pd.A.end(), 0.0);
  std::fill(pd.L.begin(), pd.L.end(), 0.0);
  const int index_load = nz;
  for (int a = 1; a < ncells + 1; ++a)
  {
    const int id_a = a - 1;
    const double* coeffs = pd.coefficients_flux(i_rhs) + (size_t)id_a * nrt;
    std::fill(pd.Te.begin(), pd.Te.end(), 0.0);
    minimisation_kernel(t, asmbl_systmtrx, pd.Te.data(), ldTe, coeffs, patch, a, pd.reversed[2 * id_a], pd.detJ[id_a],
                        &pd.J[4 * id_a], phi_scratch);
    const double* Te = pd.Te.data();
    if (k == 1)
    {
      if (requires_flux_bc)
      {
        L[0] = 0;
        if (asmbl_systmtrx)
          A[0] = 1;
      }
      else
      {
        L[0] += Te[1 * ldTe + 0];
        if (asmbl_systmtrx)
          A[0] += Te[0];
      }
    }
    else if (requires_flux_bc)
    {
      for (int i = 0; i < nz; ++i)
      {
        const int dof_i = patch.dofmap(2, a, i + 1);
        const int8_t bi = bm[dof_i];
        if (bi)
          L[dof_i] = 0;
        else
          L[dof_i] += Te[index_load * ldTe + i];
        if (asmbl_systmtrx)
        {
          if (bi)
            A[dof_i * ld + dof_i] = 1;
          else
            for (int j = 0; j < nz; ++j)
            {
              const int dof_j = patch.dofmap(2, a, j + 1);
              if (bm[dof_j])
                A[dof_i * ld + dof_j] = 0;
              else
                A[dof_i * ld + dof_j] += Te[i * ldTe + j];
            }
        }
      }
    }
    else
    {
      for (int i = 0; i < nz; ++i)
      {
        const int dof_i = patch.dofmap(2, a, i + 1);
        L[dof_i] += Te[index_load * ldTe + i];
        if (asmbl_systmtrx)
          for (int j = 0; j < nz; ++j)
            A[dof_i * ld + patch.dofmap(2, a, j + 1)] += Te[i * ldTe + j];
      }
    }
  }
}

// ---------------------------------------------------------------------------
// se/solve_patch_semiexplt.hpp:64-111
// ---------------------------------------------------------------------------
static void calculate_jump(double GtHat_Ea[2][2], const eqlb_tables* t, int nq, int iq_Ta, uint8_t Ea_reversed,
                           const int32_t* dofs_G_Ea, const double* G_Tap1Ea, int fl_Tap1Ea, int hat_Tap1, const double* G_TaEa,
                           int fl_TaEa, int hat_Ta)
{
  const int nf = t->ndg_fct, ndg = t->ndg, nqf = t->nqf;
  const int iq_Tap1 = Ea_reversed ? nq - iq_Ta - 1 : iq_Ta;
  GtHat_Ea[0][0] = GtHat_Ea[0][1] = GtHat_Ea[1][0] = GtHat_Ea[1][1] = 0.0;
  const double* shp_Ta = t->dg_f + (size_t)(fl_TaEa * nqf + iq_Ta) * ndg;
  const double* shp_Tap1 = t->dg_f + (size_t)(fl_Tap1Ea * nqf + iq_Tap1) * ndg;
  for (int i = 0; i < nf; ++i)
  {
    const int id_Ta = dofs_G_Ea[i + nf], id_Tap1 = dofs_G_Ea[i];
    GtHat_Ea[0][0] += G_TaEa[2 * id_Ta] * shp_Ta[id_Ta];
    GtHat_Ea[0][1] += G_TaEa[2 * id_Ta + 1] * shp_Ta[id_Ta];
    GtHat_Ea[1][0] += G_Tap1Ea[2 * id_Tap1] * shp_Tap1[id_Tap1];
    GtHat_Ea[1][1] += G_Tap1Ea[2 * id_Tap1 + 1] * shp_Tap1[id_Tap1];
  }
  const double h_Ta = t->hat_f[(fl_TaEa * nqf + iq_Ta) * 3 + hat_Ta];
  const double h_Tap1 = t->hat_f[(fl_Tap1Ea * nqf + iq_Tap1) * 3 + hat_Tap1];
  GtHat_Ea[0][0] *= h_Ta;
  GtHat_Ea[0][1] *= h_Ta;
  GtHat_Ea[1][0] *= h_Tap1;
  GtHat_Ea[1][1] *= h_Tap1;
}

static void copy_cell_G(const Problem& pb, const double* xg, int32_t cell, double* out)
{
  const int ndg = pb.t->ndg;
  const int32_t* dofs = pb.mv.m->dg_dofmap + (size_t)cell * ndg;
  for (int j = 0; j < ndg; ++j)
  {
    out[2 * j] = xg[2 * dofs[j]];
    out[2 * j + 1] = xg[2 * dofs[j] + 1];
  }
}

// ---------------------------------------------------------------------------
// se/solve_patch_semiexplt.hpp:212-1163
// ---------------------------------------------------------------------------
static void equilibrate_flux_semiexplt(Problem& pb, Patch& patch, PatchData& pd, const double* const* G,
                                       const double* const* F, double* const* sigma, std::vector<double>& phi_scratch)
{
  const eqlb_tables* t = pb.t;
  const eqlb_mesh* m = pb.mv.m;
  const int k = t->k, nrt = t->nrt, ndg = t->ndg, nqf = t->nqf, nq = t->nq;
  const int ndofs_projflux_fct = t->ndg_fct;
  const int ncells = patch.ncells;
  const bool fct_out[3] = {false, true, false}; // basix::cell::facet_orientations (base/KernelData.cpp:61)
  const int id_flux_order = (k == 1) ? 1 : (k == 2 ? 2 : 3);
  auto Mref = [&](int f, int j, int d, int n) { return t->M[(((size_t)f * k + j) * 2 + d) * nqf + n]; };

  double c_ta_ea = 0, c_ta_eam1 = 0, c_tam1_eam1 = 0, c_t1_e0 = 0;
  const int ndofs_hdivz = patch.ndof_min_flux;
  const int ndofs_hdivz_per_cell = 2 * k + t->nadd;

  /* Pre-evaluate repeatedly used cell data  (:297-424) */
  for (int a = 1; a < ncells + 1; ++a)
  {
    const int id_a = a - 1;
    const int32_t c = patch.cells[a];
    const int32_t* xd = m->cell_node + 3 * c;
    double* J = &pd.J[4 * id_a];
    double* K = &pd.K[4 * id_a];
    const double detJ = compute_jacobian(J, K, m->x + 3 * xd[0], m->x + 3 * xd[1], m->x + 3 * xd[2]);
    pd.detJ[id_a] = detJ;

    int8_t fl_eam1, fl_ea;
    patch.fctid_local(a, fl_eam1, fl_ea);
    const bool nout_eam1 = fct_out[fl_eam1], nout_ea = fct_out[fl_ea];

    if (patch.type[0] != internal && (a == 1 || a == ncells))
    {
      if (a == 1)
      {
        const int32_t c_ap1 = patch.cells[a + 1];
        const int8_t fl_tap1_ea = patch.fctid_local(a, a + 1);
        if (m->fct_perms[c * 3 + fl_ea] != m->fct_perms[c_ap1 * 3 + fl_tap1_ea])
          pd.reversed[2 * id_a + 1] = 1;
      }
      else
      {
        const int32_t c_am1 = patch.cells[a - 1];
        const int8_t fl_tam1_eam1 = patch.fctid_local(a - 1, a - 1);
        if (m->fct_perms[c_am1 * 3 + fl_tam1_eam1] != m->fct_perms[c * 3 + fl_eam1])
          pd.reversed[2 * id_a] = 1;
      }
    }
    else
    {
      const int32_t c_am1 = patch.cells[a - 1], c_ap1 = patch.cells[a + 1];
      const int8_t fl_tam1_eam1 = patch.fctid_local(a - 1, a - 1);
      const int8_t fl_tap1_ea = patch.fctid_local(a, a + 1);
      if (m->fct_perms[c_am1 * 3 + fl_tam1_eam1] != m->fct_perms[c * 3 + fl_eam1])
        pd.reversed[2 * id_a] = 1;
      if (m->fct_perms[c * 3 + fl_ea] != m->fct_perms[c_ap1 * 3 + fl_tap1_ea])
        pd.reversed[2 * id_a + 1] = 1;
    }

    const double sgn_detJ = detJ / std::fabs(detJ);
    pd.prefactors[2 * id_a] = nout_eam1 ? sgn_detJ : -sgn_detJ;
    pd.prefactors[2 * id_a + 1] = nout_ea ? sgn_detJ : -sgn_detJ;

    for (int i = 0; i < k + 1; ++i)
      for (int j = 0; j < nqf; ++j)
      {
        const int8_t fctid = (i == 0) ? fl_eam1 : fl_ea;
        const int ii = (i < 2) ? 0 : i - 1;
        pd.M_mapped(id_a, i, 0, j) = detJ * (Mref(fctid, ii, 0, j) * K[0] + Mref(fctid, ii, 1, j) * K[2]);
        pd.M_mapped(id_a, i, 1, j) = detJ * (Mref(fctid, ii, 0, j) * K[1] + Mref(fctid, ii, 1, j) * K[3]);
        if (ii > 0 && pd.reversed[2 * id_a + 1] && a != ncells)
        {
          pd.M_mapped(id_a, i, 0, j) -= pd.M_mapped(id_a, 1, 0, j);
          pd.M_mapped(id_a, i, 1, j) -= pd.M_mapped(id_a, 1, 1, j);
        }
      }
  }

  patch.set_assembly_informations(fct_out, pd.reversed.data(), pd.detJ.data());
  const int offs_ffEa = patch.offs[1], offs_fcdiv = patch.offs[4];
  auto dofmap_flux = [&](int pl, int a, int i) { return patch.dofmap(pl, a, i); };
  auto prefactor_dof = [&](int id, int j) { return pd.prefactors[2 * id + j]; };
  auto reversed_fct = [&](int id, int j) { return pd.reversed[2 * id + j]; };

  double* coefficients_G_Ta = pd.coefficients_G_Ta.data();
  double* coefficients_G_Tap1 = pd.coefficients_G_Tap1.data();
  double* coefficients_f = pd.coefficients_f.data();
  double* c_ta_div = pd.c_ta_div.data();
  double* cj_ta_ea = pd.cj_ta_ea.data();
  std::vector<double> shp_rhs((size_t)3 * nq * ndg);

  for (int i_rhs = 0; i_rhs < pb.nrhs; ++i_rhs)
  {
    const int8_t type_patch = patch.type[i_rhs];
    const bool reversion_required = patch.reversion_required(i_rhs);
    double* coefficients_flux = pd.coefficients_flux(i_rhs);
    auto CF = [&](int id, int i) -> double& { return coefficients_flux[(size_t)id * nrt + i]; };
    const double* x_flux_proj = G[i_rhs];
    const double* x_rhs_proj = F[i_rhs];
    const std::vector<double>& boundary_values = pb.boundary_values[i_rhs];

    /* Step 1 */
    copy_cell_G(pb, x_flux_proj, patch.cells[1], coefficients_G_Tap1);
    if (i_rhs > 0)
    {
      c_tam1_eam1 = 0.0;
      c_t1_e0 = 0.0;
      std::fill(pd.jumpG_Eam1.begin(), pd.jumpG_Eam1.end(), 0.0);
    }

    for (int a = 1; a < ncells + 1; ++a)
    {
      const int id_a = a - 1;
      const int32_t c_a = patch.cells[a];
      const int8_t node_i_Ta = patch.inodes_local[a];
      const int8_t node_i_Tap1 = patch.inodes_local[a + 1];
      const bool fct_on_boundary = patch.is_on_boundary() && (a == 1 || a == ncells);
      bool fct_has_bc = false;
      if (fct_on_boundary)
      {
        if (type_patch == bound_essnt_dual)
          fct_has_bc = true;
        else if (type_patch == bound_mixed)
        {
          if (a == 1)
            fct_has_bc = patch.requires_flux_bcs(i_rhs, 0);
          else if (a == ncells)
            fct_has_bc = patch.requires_flux_bcs(i_rhs, ncells);
        }
      }

      int8_t fl_TaEam1, fl_TaEa;
      patch.fctid_local(a, fl_TaEam1, fl_TaEa);
      // on the last cell of a boundary patch the reference still evaluates
      // fctid_local(a, a+1); it reads _fcts_local[2a+1] (== lfct of E_n in T_n)
      const int8_t fl_Tap1Ea = patch.fcts_local[2 * a + 1];

      const double detJ = pd.detJ[id_a];
      const double sign_detJ = (detJ > 0.0) ? 1.0 : -1.0;

      std::swap(coefficients_G_Ta, coefficients_G_Tap1);
      copy_cell_G(pb, x_flux_proj, patch.cells[a + 1], coefficients_G_Tap1);
      {
        const int32_t* dofs = m->dg_dofmap + (size_t)c_a * ndg;
        for (int j = 0; j < ndg; ++j)
          coefficients_f[j] = x_rhs_proj[dofs[j]];
      }

      auto shp_fct = [&](int fl, int n, int i) { return t->dg_f[(size_t)(fl * nqf + n) * ndg + i]; };
      auto hat_fct = [&](int fl, int n, int v) { return t->hat_f[(fl * nqf + n) * 3 + v]; };

      const int32_t* pflux_ldofs_E0 = nullptr;
      if (a == 1 && (fct_has_bc || type_patch == bound_mixed))
        pflux_ldofs_E0 = patch.dofs_projflux_fct(0);

      c_ta_eam1 = -c_tam1_eam1;
      std::fill(pd.cj_ta_ea.begin(), pd.cj_ta_ea.end(), 0.0);

      if (fct_has_bc)
      {
        int32_t bfct_global;
        int offs_bdofs;
        if (a == 1)
        {
          offs_bdofs = 0;
          bfct_global = patch.fcts[0];
        }
        else
        {
          offs_bdofs = k;
          bfct_global = patch.fcts[a];
        }
        calculate_patch_bc(pb, i_rhs, bfct_global, node_i_Ta, &pd.J[0], detJ, &pd.K[0]);
        if (a == 1)
          c_ta_eam1 += prefactor_dof(id_a, 0) * boundary_values[dofmap_flux(1, a, offs_bdofs)];
        if (id_flux_order > 1)
          for (int j = 1; j < k; ++j)
            CF(id_a, dofmap_flux(0, a, offs_bdofs + j)) += boundary_values[dofmap_flux(1, a, offs_bdofs + j)];
        if (reversion_required)
          c_t1_e0 -= prefactor_dof(id_a, 1) * boundary_values[dofmap_flux(1, a, offs_bdofs)];
      }

      double surfint_c_ta_eam1 = 0.0;
      double GtHat_Ea[2][2] = {{0, 0}, {0, 0}};
      double jGtHat[2];

      for (int n = 0; n < nqf; ++n)
      {
        if (fct_on_boundary)
        {
          if (a == 1)
          {
            if (fct_has_bc || type_patch == bound_mixed)
            {
              for (int i = 0; i < ndofs_projflux_fct; ++i)
              {
                const int id_Ta = pflux_ldofs_E0[i + ndofs_projflux_fct];
                pd.GtHat_Eam1(n, 0, 0) -= coefficients_G_Ta[2 * id_Ta] * shp_fct(fl_TaEam1, n, id_Ta);
                pd.GtHat_Eam1(n, 0, 1) -= coefficients_G_Ta[2 * id_Ta + 1] * shp_fct(fl_TaEam1, n, id_Ta);
              }
              pd.GtHat_Eam1(n, 0, 0) *= hat_fct(fl_TaEam1, n, node_i_Ta);
              pd.GtHat_Eam1(n, 0, 1) *= hat_fct(fl_TaEam1, n, node_i_Ta);

              if (id_flux_order > 1)
              {
                const double* Kc = &pd.K[4 * id_a];
                double g0, g1;
                if (type_patch == bound_mixed && !fct_has_bc)
                {
                  g0 = -pd.GtHat_Eam1(n, 0, 0);
                  g1 = -pd.GtHat_Eam1(n, 0, 1);
                }
                else
                {
                  g0 = pd.GtHat_Eam1(n, 0, 0);
                  g1 = pd.GtHat_Eam1(n, 0, 1);
                }
                const double m0 = detJ * (Kc[0] * g0 + Kc[1] * g1);
                const double m1 = detJ * (Kc[2] * g0 + Kc[3] * g1);
                for (int j = 1; j < k; ++j)
                  CF(id_a, dofmap_flux(0, a, j)) += Mref(fl_TaEam1, j, 0, n) * m0 + Mref(fl_TaEam1, j, 1, n) * m1;
              }
              if (!fct_has_bc)
              {
                pd.GtHat_Eam1(n, 0, 0) = 0.0;
                pd.GtHat_Eam1(n, 0, 1) = 0.0;
              }
            }
            calculate_jump(GtHat_Ea, t, nqf, n, reversed_fct(id_a, 1), patch.dofs_projflux_fct(a), coefficients_G_Tap1,
                           fl_Tap1Ea, node_i_Tap1, coefficients_G_Ta, fl_TaEa, node_i_Ta);
          }
          else
          {
            const int32_t* dofs_G_Ea = patch.dofs_projflux_fct(a);
            const double pfctr = fct_has_bc ? -1.0 : 1.0;
            GtHat_Ea[0][0] = GtHat_Ea[0][1] = GtHat_Ea[1][0] = GtHat_Ea[1][1] = 0.0;
            for (int i = 0; i < ndofs_projflux_fct; ++i)
            {
              const int id_Ta = dofs_G_Ea[i + ndofs_projflux_fct];
              const double sshp = pfctr * shp_fct(fl_TaEa, n, id_Ta);
              GtHat_Ea[1][0] += coefficients_G_Ta[2 * id_Ta] * sshp;
              GtHat_Ea[1][1] += coefficients_G_Ta[2 * id_Ta + 1] * sshp;
            }
            GtHat_Ea[1][0] *= hat_fct(fl_TaEa, n, node_i_Ta);
            GtHat_Ea[1][1] *= hat_fct(fl_TaEa, n, node_i_Ta);
            if (reversion_required)
            {
              const double* Kc = &pd.K[4 * id_a];
              const double m0 = detJ * (Kc[0] * GtHat_Ea[1][0] + Kc[1] * GtHat_Ea[1][1]);
              const double m1 = detJ * (Kc[2] * GtHat_Ea[1][0] + Kc[3] * GtHat_Ea[1][1]);
              const double aux = Mref(fl_TaEa, 0, 0, n) * m0 + Mref(fl_TaEa, 0, 1, n) * m1;
              c_t1_e0 -= prefactor_dof(id_a, 1) * aux;
            }
          }
        }
        else
        {
          calculate_jump(GtHat_Ea, t, nqf, n, reversed_fct(id_a, 1), patch.dofs_projflux_fct(a), coefficients_G_Tap1,
                         fl_Tap1Ea, node_i_Tap1, coefficients_G_Ta, fl_TaEa, node_i_Ta);
        }

        jGtHat[0] = pd.GtHat_Eam1(n, 1, 0) - pd.GtHat_Eam1(n, 0, 0);
        jGtHat[1] = pd.GtHat_Eam1(n, 1, 1) - pd.GtHat_Eam1(n, 0, 1);
        surfint_c_ta_eam1 -= pd.M_mapped(id_a, 0, 0, n) * jGtHat[0] + pd.M_mapped(id_a, 0, 1, n) * jGtHat[1];

        if (id_flux_order > 1)
        {
          jGtHat[0] = GtHat_Ea[1][0] - GtHat_Ea[0][0];
          jGtHat[1] = GtHat_Ea[1][1] - GtHat_Ea[0][1];
          for (int j = 2; j < k + 1; ++j)
            cj_ta_ea[j - 2] += pd.M_mapped(id_a, j, 0, n) * jGtHat[0] + pd.M_mapped(id_a, j, 1, n) * jGtHat[1];
        }

        pd.GtHat_Eam1(n, 0, 0) = GtHat_Ea[0][0];
        pd.GtHat_Eam1(n, 0, 1) = GtHat_Ea[0][1];
        pd.GtHat_Eam1(n, 1, 0) = GtHat_Ea[1][0];
        pd.GtHat_Eam1(n, 1, 1) = GtHat_Ea[1][1];
      }

      if (reversed_fct(id_a, 1))
        for (int i = 0; i < nqf / 2; ++i)
        {
          const int ri = nqf - 1 - i;
          std::swap(pd.GtHat_Eam1(i, 0, 0), pd.GtHat_Eam1(ri, 0, 0));
        }

      c_ta_eam1 += prefactor_dof(id_a, 0) * surfint_c_ta_eam1;
      c_t1_e0 -= prefactor_dof(id_a, 0) * surfint_c_ta_eam1;

      /* DOFs from cell integrals */
      if (id_flux_order == 1)
      {
        const double vol_int = coefficients_f[0] * (std::fabs(detJ) / 6);
        c_ta_ea = vol_int - c_ta_eam1;
        c_t1_e0 += vol_int;
      }
      else
      {
        const double* Kc = &pd.K[4 * id_a];
        // se::KernelData::shapefunctions_cell_rhs(K)  se/KernelData.cpp:225-247
        for (int n = 0; n < nq; ++n)
          for (int i = 0; i < ndg; ++i)
          {
            const double dx = t->dg_q[((size_t)1 * nq + n) * ndg + i], dy = t->dg_q[((size_t)2 * nq + n) * ndg + i];
            shp_rhs[((size_t)0 * nq + n) * ndg + i] = t->dg_q[(size_t)n * ndg + i];
            shp_rhs[((size_t)1 * nq + n) * ndg + i] = Kc[0] * dx + Kc[2] * dy;
            shp_rhs[((size_t)2 * nq + n) * ndg + i] = Kc[1] * dx + Kc[3] * dy;
          }
        c_ta_ea = -c_ta_eam1;
        std::fill(pd.c_ta_div.begin(), pd.c_ta_div.end(), 0.0);
        for (int n = 0; n < nq; ++n)
        {
          double f = 0.0, div_g = 0.0;
          for (int i = 0; i < ndg; ++i)
          {
            f += coefficients_f[i] * shp_rhs[((size_t)0 * nq + n) * ndg + i];
            div_g += coefficients_G_Ta[2 * i] * shp_rhs[((size_t)1 * nq + n) * ndg + i]
                     + coefficients_G_Ta[2 * i + 1] * shp_rhs[((size_t)2 * nq + n) * ndg + i];
          }
          const double aux = (f - div_g) * t->hat_q[n * 3 + node_i_Ta] * t->qwts[n] * detJ;
          const double vol_int = aux * sign_detJ;
          c_ta_ea += vol_int;
          c_t1_e0 += vol_int;
          const double qx = t->qpts[2 * n], qy = t->qpts[2 * n + 1];
          if (id_flux_order == 2)
          {
            c_ta_div[0] += aux * qy;
            c_ta_div[1] += aux * qx;
          }
          else
          {
            int count = 0;
            for (int l = 0; l < k; ++l)
              for (int mm = 0; mm < k - l; ++mm)
                if (l + mm > 0)
                {
                  c_ta_div[count] += aux * std::pow(qx, l) * std::pow(qy, mm);
                  ++count;
                }
          }
        }
      }

      if (id_flux_order > 1)
        if (reversed_fct(id_a, 1) && a != ncells)
          for (int i = 1; i < k; ++i)
            cj_ta_ea[i - 1] += prefactor_dof(id_a + 1, 0) * c_ta_ea;

      CF(id_a, dofmap_flux(0, a, 0)) += prefactor_dof(id_a, 0) * c_ta_eam1;
      CF(id_a, dofmap_flux(0, a, offs_ffEa)) += prefactor_dof(id_a, 1) * c_ta_ea;
      if (id_flux_order > 1)
      {
        for (int i = 1; i < k; ++i)
          CF(id_a, dofmap_flux(0, a, offs_ffEa + i)) += cj_ta_ea[i - 1];
        for (int i = 0; i < t->ndiv; ++i)
          CF(id_a, dofmap_flux(0, a, offs_fcdiv + i)) += c_ta_div[i];
      }
      c_tam1_eam1 = c_ta_ea;
    }

    if (reversion_required)
      for (int a = 1; a < ncells + 1; ++a)
      {
        const int id_a = a - 1;
        CF(id_a, dofmap_flux(0, a, 0)) += prefactor_dof(id_a, 0) * c_t1_e0;
        CF(id_a, dofmap_flux(0, a, offs_ffEa)) -= prefactor_dof(id_a, 1) * c_t1_e0;
        if (id_flux_order > 1 && reversed_fct(id_a, 1) && a != ncells)
          for (int i = 1; i < k; ++i)
            CF(id_a, dofmap_flux(0, a, offs_ffEa + i)) -= prefactor_dof(id_a + 1, 0) * c_t1_e0;
      }

    /* Step 2 */
    if (type_patch == bound_essnt_dual || type_patch == bound_mixed)
      set_boundary_markers(pd.boundary_markers.data(), pd.dim_hdivz, {type_patch}, {reversion_required}, ncells,
                           ndofs_hdivz, k);

    bool assemble_entire_system = false;
    if (i_rhs == 0)
      assemble_entire_system = true;
    else if (patch.is_on_boundary())
      if (patch.type[i_rhs] != patch.type[i_rhs - 1] || patch.type[i_rhs] == bound_mixed)
        assemble_entire_system = true;

    const int hz = pd.dim_hdivz;
    assemble_fluxminimiser(assemble_entire_system, patch, pd, i_rhs, patch.requires_flux_bcs(i_rhs), phi_scratch);
    if (assemble_entire_system && id_flux_order > 1)
      llt_factor(pd.A.data(), hz, pd.ld);

    double* u_sigma = pd.u_sigma.data();
    if (id_flux_order == 1)
      u_sigma[0] = pd.L[0] / pd.A[0];
    else
    {
      for (int i = 0; i < hz; ++i)
        u_sigma[i] = pd.L[i];
      llt_solve(pd.A.data(), hz, pd.ld, u_sigma);
    }

    /* scatter (:1094-1161) */
    double* x_flux_dhdiv = sigma[i_rhs];
    const double* doftrafo = t->trafo;
    for (int a = 1; a < ncells + 1; ++a)
    {
      const int id_a = a - 1;
      const int32_t gd0 = patch.cells[a] * nrt;
      if (id_flux_order == 1)
      {
        if (reversed_fct(id_a, 0))
          CF(id_a, dofmap_flux(0, a, 0)) += doftrafo[0] * dofmap_flux(3, a, 0) * u_sigma[dofmap_flux(2, a, 0)];
        else
          CF(id_a, dofmap_flux(0, a, 0)) += dofmap_flux(3, a, 0) * u_sigma[dofmap_flux(2, a, 0)];
        CF(id_a, dofmap_flux(0, a, 1)) += dofmap_flux(3, a, 1) * u_sigma[dofmap_flux(2, a, 1)];
      }
      else
      {
        int start_i = 0;
        if (reversed_fct(id_a, 0))
        {
          for (int i = 0; i < k; ++i)
          {
            double local_value = 0.0;
            for (int j = 0; j < k; ++j)
            {
              const double pf_j = doftrafo[j * k + i] * dofmap_flux(3, a, j);
              local_value += pf_j * u_sigma[dofmap_flux(2, a, j)];
            }
            CF(id_a, dofmap_flux(0, a, i)) += local_value;
          }
          start_i = k;
        }
        for (int i = start_i; i < ndofs_hdivz_per_cell; ++i)
          CF(id_a, dofmap_flux(0, a, i)) += dofmap_flux(3, a, i) * u_sigma[dofmap_flux(2, a, i)];
      }
      for (int i = 0; i < nrt; ++i)
        x_flux_dhdiv[gd0 + i] += CF(id_a, i);
    }
  }
}

} // namespace oracle

#include "oracle_se_stress.inc"
#include "oracle_se_api.inc"
