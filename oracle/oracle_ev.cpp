// TEST INFRASTRUCTURE - NOT PRODUCT CODE (see oracle_common.hpp).
// Restatement of the constrained-minimisation (Ern-Vohralik) equilibration:
// ev/Patch.cpp (ordering + sub-DOFmaps), ev/assembly.hpp (assemble_tangents,
// apply_lifting), ev/solve_patch.hpp (dense KKT, partial-pivot LU, += scatter),
// ev/reconstruction.hpp (serial node loop), forms of FluxEqlbEV.py:116-133.
//
// Third-party pieces absent from the reference tree and restated here from their
// definition: the FFCx cell kernels of the three fixed forms (quadrature of the
// Piola-mapped basis), the DOLFINx DOF transformation (base transformation R of a
// reflected facet = reversed-facet matrix of se/KernelData.cpp:46-64), the mixed
// RT_k x DG_(k-1) dofmap.  The conforming RT space uses the hierarchic functionals
// of e_raviart_thomas.py with facets oriented low -> high global vertex; global
// numbering [facet dofs: fct*k+j][cell dofs: nfct*k + cell*(k^2-k) + i]; the mixed
// space appends the DG dofs (nflux + cell*ndg + q).
#include "oracle_common.hpp"

namespace oracle
{

struct EvProblem
{
  MeshView mv;
  const eqlb_tables* t;
  int nrhs;
  const int8_t* facet_type;    // [nrhs][nfct]
  const double* const* bflux;  // DRT layout boundary functions (cell-local moments)
  int nflux;                   // size of the conforming flux space
  int8_t bfct_type(int rhs, int fct) const { return facet_type[(size_t)rhs * mv.m->nfct + fct]; }
  // mixed / flux dofmaps
  int ndof_elmt() const { return t->nrt + t->ndg; }
  int gdof(int cell, int ldof) const
  {
    const int k = t->k, nrt = t->nrt;
    if (ldof < 3 * k)
      return mv.m->cell_fct[3 * cell + ldof / k] * k + ldof % k;
    if (ldof < nrt)
      return mv.m->nfct * k + cell * (nrt - 3 * k) + (ldof - 3 * k);
    return nflux + cell * t->ndg + (ldof - nrt);
  }
};

// ---------------------------------------------------------------------------
// ev::OrientedPatch + ev::Patch  (ev/Patch.cpp:83-309, 482-676)
// ---------------------------------------------------------------------------
struct EvPatch
{
  const EvProblem& pb;
  const MeshView& mv;
  const int k, nrt, ndg;
  int ncells_max = 0;
  int nodei = -1, ncells = 0, nfcts = 0;
  std::vector<int8_t> type;
  std::vector<int32_t> cells, fcts, fcts_sorted;
  std::vector<int8_t> inodes_local;
  int ndof_elmt, ndof_flux_fct, ndof_flux_cell, ndof_flux, ndof_cons, ndof_elmt_nz, ndof_flux_nz;
  int ndof_patch_nz = 0, ndof_fluxhdiv = 0;
  std::vector<int32_t> dofsnz_elmt, dofsnz_patch, dofsnz_global, offset_dofmap, list_patch_fluxhdiv, list_global_fluxhdiv;

  explicit EvPatch(const EvProblem& pb_) : pb(pb_), mv(pb_.mv), k(pb_.t->k), nrt(pb_.t->nrt), ndg(pb_.t->ndg)
  {
    type.assign(pb.nrhs, internal);
    for (int i = 0; i < mv.m->nnode; ++i)
      ncells_max = std::max(ncells_max, mv.node_to_cell(i).size());
    cells.assign(ncells_max, 0);
    fcts.assign(ncells_max + 1, 0);
    fcts_sorted.assign(ncells_max + 1, 0);
    inodes_local.assign(ncells_max, 0);
    ndof_elmt = nrt + ndg;
    ndof_flux_fct = k;
    ndof_flux_cell = nrt - 3 * k;
    ndof_flux = nrt;
    ndof_cons = ndg;
    ndof_elmt_nz = ndof_elmt - k;
    ndof_flux_nz = ndof_flux - k;
    const int len = ncells_max * ndof_elmt_nz;
    dofsnz_elmt.assign(len, 0);
    dofsnz_patch.assign(len, 0);
    dofsnz_global.assign(len, 0);
    offset_dofmap.assign(ncells_max + 1, 0);
    const int lenf = ncells_max * ndof_flux_cell + (ncells_max + 1) * k;
    list_patch_fluxhdiv.assign(lenf, 0);
    list_global_fluxhdiv.assign(lenf, 0);
  }

  bool is_on_boundary() const { return type[0] != internal; }
  bool requires_flux_bcs(int i) const { return type[i] == bound_essnt_dual || type[i] == bound_mixed; }

  int8_t get_fctid_local(int32_t fct, Links fc) const
  {
    int8_t l = 0;
    while (l < 3 && fc[l] != fct)
      ++l;
    return l;
  }
  int8_t nodei_local(int32_t cell) const
  {
    Links nc = mv.cell_to_node(cell);
    int8_t l = 0;
    while (nc[l] != nodei)
      ++l;
    return l;
  }
  // ev/Patch.cpp:360-437
  int32_t next_facet_triangle(Links fc, int8_t lf) const
  {
    int a = (lf == 0) ? 1 : 0, b = (lf == 2) ? 1 : 2;
    int32_t e0 = std::min(fc[a], fc[b]), e1 = std::max(fc[a], fc[b]);
    if (e0 < fcts_sorted[0])
      return e1;
    if (e1 > fcts_sorted[nfcts - 1])
      return e0;
    if (std::count(fcts_sorted.begin(), fcts_sorted.begin() + nfcts, e0))
      return e0;
    return e1;
  }

  // ev/Patch.cpp:83-220
  int32_t initialize_patch(int node_i)
  {
    nodei = node_i;
    Links pc = mv.node_to_cell(node_i), pf = mv.node_to_fct(node_i);
    ncells = pc.size();
    nfcts = pf.size();
    std::copy(pf.begin(), pf.end(), fcts_sorted.begin());
    std::sort(fcts_sorted.begin(), fcts_sorted.begin() + nfcts);
    std::fill(type.begin(), type.end(), internal);
    int32_t fct_first = pf[0];
    if (nfcts > ncells)
    {
      int32_t fct_ef[2] = {-1, -1}, fct_ep[2] = {-1, -1};
      for (int32_t id : pf)
      {
        if (pb.bfct_type(0, id) == essnt_primal)
          (fct_ep[0] < 0 ? fct_ep[0] : fct_ep[1]) = id;
        else if (pb.bfct_type(0, id) == essnt_dual)
          (fct_ef[0] < 0 ? fct_ef[0] : fct_ef[1]) = id;
      }
      if (fct_ef[0] < 0)
      {
        type[0] = bound_essnt_primal;
        fct_first = fct_ep[0];
      }
      else
      {
        type[0] = (fct_ep[0] < 0) ? bound_essnt_dual : bound_mixed;
        fct_first = fct_ef[0];
      }
      for (int i = 1; i < pb.nrhs; ++i)
      {
        int32_t f0, fn;
        if (type[0] == bound_essnt_primal)
        {
          f0 = fct_ep[0];
          fn = fct_ep[1];
        }
        else if (type[0] == bound_essnt_dual)
        {
          f0 = fct_ef[0];
          fn = fct_ef[1];
        }
        else
        {
          f0 = fct_ef[0];
          fn = fct_ep[0];
        }
        if (pb.bfct_type(i, f0) == pb.bfct_type(i, fn))
          type[i] = (pb.bfct_type(i, f0) == essnt_primal) ? bound_essnt_primal : bound_essnt_dual;
        else
          type[i] = bound_mixed;
      }
    }
    return fct_first;
  }

  // ev/Patch.cpp:222-309
  void fcti_to_celli(int c_fct, int32_t fct_i, int32_t cell_in, int8_t& lf_ci, int8_t& lf_cim1, int32_t& fct_next)
  {
    Links cf = mv.fct_to_cell(fct_i);
    int32_t cell_i, cell_im1;
    if (type[0] != internal && c_fct == 0)
    {
      cell_i = cf[0];
      cell_im1 = cell_i;
      lf_ci = get_fctid_local(fct_i, mv.cell_to_fct(cell_i));
      lf_cim1 = lf_ci;
    }
    else
    {
      if (cf[0] == cell_in)
      {
        cell_i = cf[1];
        cell_im1 = cf[0];
      }
      else
      {
        cell_i = cf[0];
        cell_im1 = cf[1];
      }
      lf_ci = get_fctid_local(fct_i, mv.cell_to_fct(cell_i));
      lf_cim1 = get_fctid_local(fct_i, mv.cell_to_fct(cell_im1));
    }
    const int8_t inode = nodei_local(cell_i);
    fct_next = next_facet_triangle(mv.cell_to_fct(cell_i), lf_ci);
    if (type[0] != internal)
    {
      cells[c_fct] = cell_i;
      inodes_local[c_fct] = inode;
    }
    else if (c_fct < nfcts - 1)
    {
      cells[c_fct + 1] = cell_i;
      cells[c_fct] = cell_im1;
      inodes_local[c_fct + 1] = inode;
    }
    else
    {
      cells[0] = cell_i;
      inodes_local[0] = inode;
    }
  }

  // ev/Patch.cpp:482-676
  void create_subdofmap(int node_i)
  {
    int32_t fct_i = initialize_patch(node_i);
    const bool bnd = is_on_boundary();
    const int ndof_cell = ndof_flux_cell + ndof_cons;
    const int ndof_fct = 3 * k;
    ndof_patch_nz = nfcts * k + ncells * ndof_cell;
    ndof_fluxhdiv = nfcts * k + ncells * ndof_flux_cell;
    int32_t cell_i = -1, dof_patch = 0, offs_l = 0;
    for (int ii = 0; ii < ncells; ++ii)
    {
      int8_t lf_ci, lf_cim1;
      int32_t fct_next;
      fcti_to_celli(ii, fct_i, cell_i, lf_ci, lf_cim1, fct_next);
      int32_t offs_p = (ii + 1) * ndof_elmt_nz;
      offset_dofmap[ii + 1] = offs_p;
      int32_t offs_f;
      if (bnd)
      {
        cell_i = cells[ii];
        offs_f = (ii == 0) ? k : offset_dofmap[ii - 1] + k;
        offs_p = offset_dofmap[ii];
      }
      else if (ii < nfcts - 1)
      {
        cell_i = cells[ii + 1];
        offs_f = offset_dofmap[ii] + k;
      }
      else
      {
        cell_i = cells[0];
        offs_f = offset_dofmap[ii] + k;
        offs_p = 0;
      }
      for (int jj = 0; jj < k; ++jj)
      {
        const int ldof = lf_ci * k + jj;
        const int g = pb.gdof(cell_i, ldof);
        dofsnz_elmt[offs_p] = ldof;
        dofsnz_elmt[offs_f + jj] = lf_cim1 * k + jj;
        dofsnz_patch[offs_p] = dof_patch;
        dofsnz_patch[offs_f + jj] = dof_patch;
        dofsnz_global[offs_p] = g;
        dofsnz_global[offs_f + jj] = g;
        list_patch_fluxhdiv[offs_l] = dof_patch;
        list_global_fluxhdiv[offs_l] = g;
        ++dof_patch;
        ++offs_p;
        ++offs_l;
      }
      offs_p += k;
      for (int jj = 0; jj < ndof_flux_cell; ++jj)
      {
        const int ldof = ndof_fct + jj;
        dofsnz_elmt[offs_p] = ldof;
        dofsnz_patch[offs_p] = dof_patch;
        dofsnz_global[offs_p] = pb.gdof(cell_i, ldof);
        list_patch_fluxhdiv[offs_l] = dof_patch;
        list_global_fluxhdiv[offs_l] = pb.gdof(cell_i, ldof);
        ++dof_patch;
        ++offs_p;
        ++offs_l;
      }
      for (int jj = ndof_flux_cell; jj < ndof_cell; ++jj)
      {
        const int ldof = ndof_fct + jj;
        dofsnz_elmt[offs_p] = ldof;
        dofsnz_patch[offs_p] = dof_patch;
        dofsnz_global[offs_p] = pb.gdof(cell_i, ldof);
        ++dof_patch;
        ++offs_p;
      }
      fcts[ii] = fct_i;
      fct_i = fct_next;
    }
    if (bnd)
    {
      const int8_t lf = get_fctid_local(fct_i, mv.cell_to_fct(cell_i));
      int32_t offs_p = (ncells - 1) * ndof_elmt_nz + k;
      for (int jj = 0; jj < k; ++jj)
      {
        const int ldof = lf * k + jj;
        dofsnz_elmt[offs_p] = ldof;
        dofsnz_patch[offs_p] = dof_patch;
        dofsnz_global[offs_p] = pb.gdof(cell_i, ldof);
        list_patch_fluxhdiv[offs_l] = dof_patch;
        list_global_fluxhdiv[offs_l] = pb.gdof(cell_i, ldof);
        ++dof_patch;
        ++offs_p;
        ++offs_l;
        fcts[nfcts - 1] = fct_i;
      }
    }
  }
};

// ---------------------------------------------------------------------------
// Cell tensors of the fixed forms (FluxEqlbEV.py:116-133), role of the FFCx
// kernels + DOLFINx DOF transformations (ev/assembly.hpp:170-198)
//   a     = (sig, v) - (r, div v) + (div sig, q)
//   l_pen = (1, q)
//   l     = (hat G, v) + (hat f + grad(hat).G, q)
// in the GLOBAL basis (facets oriented low -> high vertex).
// ---------------------------------------------------------------------------
struct EvCell
{
  const EvProblem& pb;
  const int k, nrt, ndg, nd;
  std::vector<double> phi, dphi, Tm;  // Piola-mapped basis [nq][nrt][2], divergence [nq][nrt], transformation
  std::vector<double> divref;         // reference divergence at q-points [nq][nrt]
  explicit EvCell(const EvProblem& p) : pb(p), k(p.t->k), nrt(p.t->nrt), ndg(p.t->ndg), nd(p.t->nrt + p.t->ndg)
  {
    const eqlb_tables* t = pb.t;
    phi.resize((size_t)t->nq * nrt * 2);
    // reference divergence from the hierarchic functionals: div phi_i is the unique
    // P_{k-1} polynomial with moments int div phi_i x^l y^m = delta (cell dofs) and
    // int div phi_i = sum of zero-order facet fluxes.  Recover it at the quadrature
    // points from its DG_{k-1} expansion via the dg_q table: solve the small moment
    // problem per basis function.
    const int nq = t->nq, np = k * (k + 1) / 2;
    divref.assign((size_t)nq * nrt, 0.0);
    // monomial basis of P_{k-1}: x^l y^m in the order (0,0), div_lm...
    std::vector<int> ll(np), mm(np);
    ll[0] = mm[0] = 0;
    for (int i = 0; i < t->ndiv; ++i)
    {
      ll[1 + i] = t->div_lm[2 * i];
      mm[1 + i] = t->div_lm[2 * i + 1];
    }
    // Gram matrix of monomials Gm[a][b] = int x^(la+lb) y^(ma+mb)
    auto fact = [](int n)
    {
      double f = 1;
      for (int i = 2; i <= n; ++i)
        f *= i;
      return f;
    };
    std::vector<double> Gm((size_t)np * np);
    for (int a = 0; a < np; ++a)
      for (int b = 0; b < np; ++b)
      {
        const int l = ll[a] + ll[b], m = mm[a] + mm[b];
        Gm[a * np + b] = fact(l) * fact(m) / fact(l + m + 2);
      }
    std::vector<int> piv(np);
    lu_factor(Gm.data(), np, np, piv.data());
    for (int i = 0; i < nrt; ++i)
    {
      // moments of div phi_i against the monomials
      std::vector<double> mom(np, 0.0);
      if (i < 3 * k)
      {
        // facet function: int div = int_boundary phi.n_out; only the zero order
        // function has a net flux: +-1 depending on the outward flag of the facet
        if (i % k == 0)
          mom[0] = (i / k == 1) ? 1.0 : -1.0;
      }
      else if (i < 3 * k + t->ndiv)
        mom[1 + (i - 3 * k)] = 1.0;
      lu_solve(Gm.data(), np, np, piv.data(), mom.data());
      for (int q = 0; q < nq; ++q)
      {
        double s = 0;
        for (int a = 0; a < np; ++a)
          s += mom[a] * std::pow(t->qpts[2 * q], ll[a]) * std::pow(t->qpts[2 * q + 1], mm[a]);
        divref[(size_t)q * nrt + i] = s;
      }
    }
  }

  // dof transformation matrix T (nrt x nrt): phi_global = T phi_local
  void transformation(int cell, std::vector<double>& T) const
  {
    T.assign((size_t)nrt * nrt, 0.0);
    for (int i = 0; i < nrt; ++i)
      T[i * nrt + i] = 1.0;
    for (int f = 0; f < 3; ++f)
      if (pb.mv.m->fct_perms[3 * cell + f])
        for (int i = 0; i < k; ++i)
          for (int j = 0; j < k; ++j)
            T[(f * k + i) * nrt + f * k + j] = pb.t->trafo[i * k + j];
  }

  // Ae [nd][nd], Pe [ndg]
  void tangent(int cell, double* Ae, double* Pe)
  {
    const eqlb_tables* t = pb.t;
    const eqlb_mesh* m = pb.mv.m;
    const int nq = t->nq;
    const int32_t* xd = m->cell_node + 3 * cell;
    double J[4], K[4];
    const double detJ = compute_jacobian(J, K, m->x + 3 * xd[0], m->x + 3 * xd[1], m->x + 3 * xd[2]);
    std::vector<double> Aloc((size_t)nd * nd, 0.0);
    std::fill(Pe, Pe + ndg, 0.0);
    for (int q = 0; q < nq; ++q)
    {
      const double dvol = t->qwts[q] * std::fabs(detJ);
      for (int i = 0; i < nrt; ++i)
      {
        const double r0 = t->rt_q[((size_t)q * nrt + i) * 2], r1 = t->rt_q[((size_t)q * nrt + i) * 2 + 1];
        phi[((size_t)q * nrt + i) * 2] = (J[0] * r0 + J[1] * r1) / detJ;
        phi[((size_t)q * nrt + i) * 2 + 1] = (J[2] * r0 + J[3] * r1) / detJ;
      }
      for (int i = 0; i < nrt; ++i)
      {
        const double pi0 = phi[((size_t)q * nrt + i) * 2], pi1 = phi[((size_t)q * nrt + i) * 2 + 1];
        const double divi = divref[(size_t)q * nrt + i] / detJ;
        for (int j = 0; j < nrt; ++j)
          Aloc[i * nd + j] += dvol * (pi0 * phi[((size_t)q * nrt + j) * 2] + pi1 * phi[((size_t)q * nrt + j) * 2 + 1]);
        for (int c = 0; c < ndg; ++c)
        {
          const double psi = t->dg_q[(size_t)q * ndg + c];
          Aloc[i * nd + nrt + c] -= dvol * psi * divi;   // -(r, div v): row v_i, col r_c
          Aloc[(nrt + c) * nd + i] += dvol * divi * psi;  // (div sig, q): row q_c, col sig_i
        }
      }
      for (int c = 0; c < ndg; ++c)
        Pe[c] += dvol * t->dg_q[(size_t)q * ndg + c];
    }
    // DOF transformation  A_g = T A T^T on the flux block (dof_transform + _to_transpose)
    std::vector<double> T;
    transformation(cell, T);
    std::vector<double> tmp((size_t)nd * nd, 0.0);
    for (int i = 0; i < nd; ++i)
      for (int j = 0; j < nd; ++j)
      {
        double s = 0;
        if (i < nrt)
          for (int c = 0; c < nrt; ++c)
            s += T[i * nrt + c] * Aloc[c * nd + j];
        else
          s = Aloc[i * nd + j];
        tmp[i * nd + j] = s;
      }
    for (int i = 0; i < nd; ++i)
      for (int j = 0; j < nd; ++j)
      {
        double s = 0;
        if (j < nrt)
          for (int c = 0; c < nrt; ++c)
            s += tmp[i * nd + c] * T[j * nrt + c];
        else
          s = tmp[i * nd + j];
        Ae[i * nd + j] = s;
      }
  }

  // Le [nd] for hat function at local node `inode`
  void load(int cell, int inode, const double* Gx, const double* Fx, double* Le)
  {
    const eqlb_tables* t = pb.t;
    const eqlb_mesh* m = pb.mv.m;
    const int nq = t->nq;
    const int32_t* xd = m->cell_node + 3 * cell;
    double J[4], K[4];
    const double detJ = compute_jacobian(J, K, m->x + 3 * xd[0], m->x + 3 * xd[1], m->x + 3 * xd[2]);
    const int32_t* dofs = m->dg_dofmap + (size_t)cell * ndg;
    const double ghat_ref[3][2] = {{-1, -1}, {1, 0}, {0, 1}};
    const double gh0 = K[0] * ghat_ref[inode][0] + K[2] * ghat_ref[inode][1];
    const double gh1 = K[1] * ghat_ref[inode][0] + K[3] * ghat_ref[inode][1];
    std::vector<double> Lloc(nd, 0.0);
    for (int q = 0; q < nq; ++q)
    {
      const double dvol = t->qwts[q] * std::fabs(detJ);
      double g0 = 0, g1 = 0, f = 0;
      for (int c = 0; c < ndg; ++c)
      {
        const double psi = t->dg_q[(size_t)q * ndg + c];
        g0 += Gx[2 * dofs[c]] * psi;
        g1 += Gx[2 * dofs[c] + 1] * psi;
        f += Fx[dofs[c]] * psi;
      }
      const double hat = t->hat_q[q * 3 + inode];
      for (int i = 0; i < nrt; ++i)
      {
        const double r0 = t->rt_q[((size_t)q * nrt + i) * 2], r1 = t->rt_q[((size_t)q * nrt + i) * 2 + 1];
        const double p0 = (J[0] * r0 + J[1] * r1) / detJ, p1 = (J[2] * r0 + J[3] * r1) / detJ;
        Lloc[i] += dvol * hat * (g0 * p0 + g1 * p1);
      }
      const double dq = hat * f + gh0 * g0 + gh1 * g1;
      for (int c = 0; c < ndg; ++c)
        Lloc[nrt + c] += dvol * dq * t->dg_q[(size_t)q * ndg + c];
    }
    std::vector<double> T;
    transformation(cell, T);
    for (int i = 0; i < nd; ++i)
    {
      double s = 0;
      if (i < nrt)
        for (int c = 0; c < nrt; ++c)
          s += T[i * nrt + c] * Lloc[c];
      else
        s = Lloc[i];
      Le[i] = s;
    }
  }
};

// ---------------------------------------------------------------------------
// patch boundary values in the global basis (BoundaryData.cpp:636-684, 686-745)
// ---------------------------------------------------------------------------
static void ev_patch_bc(const EvProblem& pb, int rhs, int32_t fct, int8_t hat_id, std::vector<double>& bvalues)
{
  if (pb.bfct_type(rhs, fct) != essnt_dual)
    return;
  const eqlb_tables* t = pb.t;
  const int k = t->k, nrt = t->nrt;
  const int32_t cell = pb.mv.fct_to_cell(fct)[0];
  const int lf = (pb.mv.m->cell_fct[3 * cell] == fct) ? 0 : ((pb.mv.m->cell_fct[3 * cell + 1] == fct) ? 1 : 2);
  const double* xb = pb.bflux[rhs];
  std::vector<double> b(k), bv(k, 0.0);
  int nzero = 0;
  for (int i = 0; i < k; ++i)
  {
    b[i] = xb ? xb[(size_t)cell * nrt + lf * k + i] : 0.0;
    if (std::fabs(b[i]) < 1e-7)
      ++nzero;
  }
  if (nzero < k)
    for (int j = 0; j < k; ++j)
      for (int i = 0; i < k; ++i)
        bv[j] += t->bc_mat[((lf * 3 + hat_id) * k + j) * k + i] * b[i];
  // local functional values -> global dofs: c_g = R^T c_loc (R involution)
  const bool refl = pb.mv.m->fct_perms[3 * cell + lf];
  for (int i = 0; i < k; ++i)
  {
    double s = 0;
    if (refl)
      for (int j = 0; j < k; ++j)
        s += t->trafo[j * k + i] * bv[j];
    else
      s = bv[i];
    bvalues[(size_t)fct * k + i] = s;
  }
}

} // namespace oracle

extern "C"
{
extern const char* oracle_last_error();
void oracle_set_error(const char* msg);

// ev::reconstruction (ev/reconstruction.hpp:64-141) + equilibrate_flux_constrmin
// (ev/solve_patch.hpp:58-239).  sigma[rhs] [nfct*k + ncell*(k^2-k)] accumulated.
int oracle_ev_run(const eqlb_mesh* mesh, const eqlb_tables* tables, int nrhs, const int8_t* facet_type,
                  const double* const* bflux, const double* const* G, const double* const* F, double* const* sigma)
{
  using namespace oracle;
  try
  {
    const int k = tables->k, nrt = tables->nrt, ndg = tables->ndg;
    EvProblem pb{MeshView{mesh}, tables, nrhs, facet_type, bflux, mesh->nfct * k + mesh->ncell * (nrt - 3 * k)};
    EvPatch patch(pb);
    EvCell cellk(pb);
    const int nd = nrt + ndg;
    // StorageStiffness (ev/StorageStiffness.hpp): A_e cached once per cell
    std::vector<double> storA((size_t)mesh->ncell * nd * nd), storP((size_t)mesh->ncell * ndg);
    std::vector<uint8_t> evaluated(mesh->ncell, 0);
    // boundary markers (BoundaryData ctor) on flux dofs, patch bc scratch
    std::vector<std::vector<int8_t>> bmarkers(nrhs, std::vector<int8_t>(pb.nflux + (size_t)mesh->ncell * ndg, 0));
    std::vector<std::vector<double>> bvalues(nrhs, std::vector<double>(pb.nflux + (size_t)mesh->ncell * ndg, 0.0));
    for (int r = 0; r < nrhs; ++r)
      for (int f = 0; f < mesh->nfct; ++f)
        if (pb.bfct_type(r, f) == essnt_dual)
          for (int j = 0; j < k; ++j)
            bmarkers[r][(size_t)f * k + j] = 1;

    const int nmax = patch.ncells_max * (nrt - 3 * k + ndg) + (patch.ncells_max + 1) * k + 1;
    std::vector<double> A((size_t)nmax * nmax), L(nmax), u(nmax), Le(nd);
    std::vector<int> piv(nmax);

    for (int node = 0; node < mesh->nnode; ++node)
    {
      if (mesh->node_owned && !mesh->node_owned[node])
        continue;
      patch.create_subdofmap(node);
      const int ncells = patch.ncells;
      const int ndof_patch = patch.ndof_patch_nz, n = ndof_patch + 1;
      const int nz = patch.ndof_elmt_nz;
      if (patch.is_on_boundary())
      {
        const int32_t bf[2] = {patch.fcts[0], patch.fcts[patch.nfcts - 1]};
        const int8_t hn[2] = {patch.inodes_local[0], patch.inodes_local[ncells - 1]};
        for (int i = 0; i < 2; ++i)
          for (int r = 0; r < nrhs; ++r)
            ev_patch_bc(pb, r, bf[i], hn[i], bvalues[r]);
      }
      for (int r = 0; r < nrhs; ++r)
      {
        bool entire = (r == 0);
        if (r > 0 && patch.is_on_boundary())
          if (patch.type[r] != patch.type[r - 1] || patch.type[r] == bound_mixed)
            entire = true;
        if (entire)
          std::fill(A.begin(), A.begin() + (size_t)n * n, 0.0);
        std::fill(L.begin(), L.begin() + n, 0.0);
        const int8_t tp = patch.type[r];
        const std::vector<int8_t>& bm = bmarkers[r];
        const std::vector<double>& bv = bvalues[r];
        for (int index = 0; index < ncells; ++index)
        {
          const int32_t c = patch.cells[index];
          double* Ae = &storA[(size_t)c * nd * nd];
          double* Pe = &storP[(size_t)c * ndg];
          if (entire && !evaluated[c])
          {
            cellk.tangent(c, Ae, Pe);
            evaluated[c] = 1;
          }
          cellk.load(c, patch.inodes_local[index], G[r], F[r], Le.data());
          const int32_t* de = &patch.dofsnz_elmt[patch.offset_dofmap[index]];
          const int32_t* dp = &patch.dofsnz_patch[patch.offset_dofmap[index]];
          const int32_t* dg = &patch.dofsnz_global[patch.offset_dofmap[index]];
          if (patch.requires_flux_bcs(r))
          {
            // apply_lifting (ev/assembly.hpp:53-87)
            for (int a = 0; a < nz; ++a)
              if (bm[dg[a]] == 0)
                for (int l = 0; l < nz; ++l)
                  if (bm[dg[l]] != 0)
                    Le[de[a]] -= Ae[de[a] * nd + de[l]] * bv[dg[l]];
            for (int a = 0; a < nz; ++a)
            {
              if (bm[dg[a]] != 0)
              {
                if (entire)
                  A[(size_t)dp[a] * n + dp[a]] = 1;
                L[dp[a]] = bv[dg[a]];
              }
              else
              {
                L[dp[a]] += Le[de[a]];
                if (entire)
                  for (int l = 0; l < nz; ++l)
                    if (bm[dg[l]] == 0)
                      A[(size_t)dp[a] * n + dp[l]] += Ae[de[a] * nd + de[l]];
              }
            }
          }
          else
          {
            for (int a = 0; a < nz; ++a)
            {
              L[dp[a]] += Le[de[a]];
              if (entire)
                for (int l = 0; l < nz; ++l)
                  A[(size_t)dp[a] * n + dp[l]] += Ae[de[a] * nd + de[l]];
            }
          }
          if (entire)
          {
            if (tp == internal || tp == bound_essnt_dual)
            {
              const int offset = patch.ndof_flux_nz;
              for (int q = 0; q < ndg; ++q)
              {
                A[(size_t)dp[offset + q] * n + (n - 1)] += Pe[q];
                A[(size_t)(n - 1) * n + dp[offset + q]] += Pe[q];
              }
            }
            else
              A[(size_t)(n - 1) * n + (n - 1)] = 1.0;
          }
        }
        if (entire)
          lu_factor(A.data(), n, n, piv.data());
        for (int i = 0; i < n; ++i)
          u[i] = L[i];
        lu_solve(A.data(), n, n, piv.data(), u.data());
        double* x = sigma[r];
        for (int q = 0; q < patch.ndof_fluxhdiv; ++q)
          x[patch.list_global_fluxhdiv[q]] += u[patch.list_patch_fluxhdiv[q]];
      }
    }
  }
  catch (const std::exception& e)
  {
    oracle_set_error(e.what());
    return -1;
  }
  return 0;
}

// EV integer maps of one patch (ev/Patch.cpp:482-676) for parity checks
int oracle_ev_patch_maps(const eqlb_mesh* mesh, const eqlb_tables* tables, int nrhs, const int8_t* facet_type, int node,
                         int32_t* ncells, int32_t* cells, int32_t* fcts, int8_t* inodes_local, int32_t* dofs_elmt,
                         int32_t* dofs_patch, int32_t* dofs_global, int32_t* list_patch, int32_t* list_global)
{
  using namespace oracle;
  try
  {
    const int k = tables->k, nrt = tables->nrt;
    EvProblem pb{MeshView{mesh}, tables, nrhs, facet_type, nullptr, mesh->nfct * k + mesh->ncell * (nrt - 3 * k)};
    EvPatch patch(pb);
    patch.create_subdofmap(node);
    *ncells = patch.ncells;
    for (int i = 0; i < patch.ncells; ++i)
    {
      cells[i] = patch.cells[i];
      inodes_local[i] = patch.inodes_local[i];
    }
    for (int i = 0; i < patch.nfcts; ++i)
      fcts[i] = patch.fcts[i];
    for (int i = 0; i < patch.ncells * patch.ndof_elmt_nz; ++i)
    {
      dofs_elmt[i] = patch.dofsnz_elmt[i];
      dofs_patch[i] = patch.dofsnz_patch[i];
      dofs_global[i] = patch.dofsnz_global[i];
    }
    for (int i = 0; i < patch.ndof_fluxhdiv; ++i)
    {
      list_patch[i] = patch.list_patch_fluxhdiv[i];
      list_global[i] = patch.list_global_fluxhdiv[i];
    }
  }
  catch (const std::exception& e)
  {
    oracle_set_error(e.what());
    return -1;
  }
  return 0;
}
}
