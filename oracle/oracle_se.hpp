// TEST INFRASTRUCTURE - NOT PRODUCT CODE (see oracle_common.hpp).
// Restatement of the semi-explicit equilibration: se/Patch.{hpp,cpp},
// se/PatchData.hpp, se/solve_patch_semiexplt.hpp, se/assembly.hpp,
// se/fluxmin_kernel.hpp, se/KernelData.cpp, base/BoundaryData.cpp:636-780.
#pragma once

#include "oracle_common.hpp"

namespace oracle
{

// ---------------------------------------------------------------------------
// Problem data (se::ProblemData + the part of base::BoundaryData the hot path
// reads, base/BoundaryData.cpp:636-780)
// ---------------------------------------------------------------------------
struct Problem
{
  MeshView mv;
  const eqlb_tables* t;
  int nrhs;
  const int8_t* facet_type;      // [nrhs][nfct]
  const double* const* bflux;    // [nrhs] -> [ncell*nrt] boundary functions (may be null)
  const int8_t* local_fct_id;    // [nfct]
  const int8_t* pnt_on_bndr;     // [nnode] (stress) or null
  bool stress;
  std::vector<std::vector<double>> boundary_values; // patch-BC scratch, global sized

  int8_t bfct_type(int rhs, int fct) const { return facet_type[(size_t)rhs * mv.m->nfct + fct]; }
};

// ---------------------------------------------------------------------------
// se::OrientedPatch + se::Patch
// ---------------------------------------------------------------------------
struct Patch
{
  const Problem& pb;
  const MeshView& mv;
  const int k;        // _ndof_flux_fct (== Basix degree)
  const int nrt;      // _ndof_flux
  const int nadd;     // _ndof_flux_add_cell
  const int ndiv;     // _ndof_flux_div_cell
  const int ndg_fct;  // _ndof_fluxdg_fct
  const bool symconstr;
  int ncells_max = 0, groupsize_max = 1;

  int nodei = -1, ncells = 0, nfcts = 0, nrhs;
  std::vector<int8_t> type;
  std::vector<int32_t> cells, fcts, fcts_sorted;
  std::vector<int8_t> inodes_local, fcts_local;

  // DOF map (se/Patch.hpp:449-461): 4 x (ncells+2) x ndpc
  int ndpc;
  int offs[5];
  std::vector<int32_t> ddofmap;
  std::vector<int32_t> list_fctdofs_fluxdg; // (ncells+1) x 2*ndg_fct
  int ndof_min_flux = 0;

  Patch(const Problem& pb_, bool symconstr_, int ncells_min, int ncells_crit)
      : pb(pb_), mv(pb_.mv), k(pb_.t->k), nrt(pb_.t->nrt), nadd(pb_.t->nadd), ndiv(pb_.t->ndiv),
        ndg_fct(pb_.t->ndg_fct), symconstr(symconstr_), nrhs(pb_.nrhs)
  {
    type.assign(nrhs, internal);
    set_max_patch_size(ncells_min, ncells_crit);
    const int sp2 = ncells_max + 2;
    cells.assign(sp2, 0);
    fcts.assign(sp2, 0);
    fcts_sorted.assign(sp2, 0);
    fcts_local.assign(2 * (ncells_max + 1), 0);
    inodes_local.assign(sp2, 0);
    const int ndof_flux_nz = nrt - k; // se/Patch.hpp:443
    ndpc = symconstr ? ndof_flux_nz + 3 : ndof_flux_nz;
    offs[0] = 0;
    offs[1] = k;
    offs[2] = 2 * k;
    offs[3] = offs[2] + nadd;
    offs[4] = symconstr ? offs[3] + 3 : offs[3];
    ddofmap.assign((size_t)4 * sp2 * ndpc, 0);
    list_fctdofs_fluxdg.assign((size_t)2 * (ncells_max + 1) * ndg_fct, 0);
  }

  int& dofmap(int pl, int a, int i) { return ddofmap[((size_t)pl * (ncells + 2) + a) * ndpc + i]; }
  int dofmap(int pl, int a, int i) const { return ddofmap[((size_t)pl * (ncells + 2) + a) * ndpc + i]; }
  int32_t& dofs_fluxdg(int a, int i) { return list_fctdofs_fluxdg[(size_t)a * 2 * ndg_fct + i]; }
  const int32_t* dofs_projflux_fct(int fct_i) const { return &list_fctdofs_fluxdg[(size_t)2 * ndg_fct * fct_i]; }

  bool is_internal() const { return type[0] == internal; }
  bool is_on_boundary() const { return type[0] != internal; }
  int ncells_of(int node) const { return mv.node_to_cell(node).size(); }

  // se/Patch.cpp:337-404
  void set_max_patch_size(int ncells_min, int ncells_crit)
  {
    const int nnodes = mv.m->nnode;
    ncells_max = 0;
    groupsize_max = 1;
    for (int i = 0; i < nnodes; ++i)
    {
      int nc = mv.node_to_cell(i).size();
      const bool owned = !(mv.m->node_owned && !mv.m->node_owned[i]);
      if (owned && nc == ncells_min)
        throw std::runtime_error("Patch around node " + std::to_string(i) + " has only "
                                 + std::to_string(ncells_min) + " cells.");
      ncells_max = std::max(ncells_max, nc);
      if (ncells_crit != 1)
      {
        int gs = (int)group_boundary_patches(i, ncells_crit).size() - 1;
        groupsize_max = std::max(groupsize_max, gs);
      }
    }
  }

  // se/Patch.cpp:60-104
  std::vector<int32_t> group_boundary_patches(int node_i, int ncells_crit) const
  {
    std::vector<int32_t> grouped;
    const int8_t* pob = pb.pnt_on_bndr;
    if (pob && pob[node_i] && ncells_of(node_i) == ncells_crit)
    {
      int inner = adjacent_internal_patch(node_i);
      grouped.push_back(inner);
      for (int32_t cell : mv.node_to_cell(inner))
        for (int32_t pnt : mv.cell_to_node(cell))
          if (pob[pnt] && std::find(grouped.begin(), grouped.end(), pnt) == grouped.end()
              && ncells_of(pnt) == ncells_crit)
            grouped.push_back(pnt);
    }
    return grouped;
  }

  // se/Patch.cpp:761-784
  int32_t adjacent_internal_patch(int node_i) const
  {
    int32_t inner = -1;
    for (int32_t fct : mv.node_to_fct(node_i))
      if (pb.bfct_type(0, fct) == f_internal)
      {
        Links nf = mv.fct_to_node(fct);
        inner = (nf[0] == node_i) ? nf[1] : nf[0];
        break;
      }
    return inner;
  }

  bool requires_flux_bcs(int index) const { return type[index] == bound_essnt_dual || type[index] == bound_mixed; }
  bool requires_flux_bcs(int index, int fct_id) const { return pb.bfct_type(index, fcts[fct_id]) == essnt_dual; }

  // se/Patch.cpp:106-128
  bool reversion_required(int index) const
  {
    bool rev = false;
    if (index > 0 && requires_flux_bcs(index))
    {
      if (type[index] != type[index - 1] || type[index] == bound_mixed)
        if (pb.bfct_type(index, fcts[0]) != essnt_dual)
          rev = true;
    }
    return rev;
  }

  int8_t get_fctid_local(int32_t fct, Links fc) const
  {
    int8_t l = 0;
    while (l < 3 && fc[l] != fct)
      ++l;
    return l;
  }
  int8_t node_local(int32_t cell, int32_t node) const
  {
    Links nc = mv.cell_to_node(cell);
    int8_t l = 0;
    while (nc[l] != node)
      ++l;
    return l;
  }

  // se/Patch.cpp:682-759
  int32_t next_facet(Links fc, int8_t lf) const
  {
    int32_t e0, e1;
    int a = (lf == 0) ? 1 : 0, b = (lf == 2) ? 1 : 2;
    if (fc[a] < fc[b])
    {
      e0 = fc[a];
      e1 = fc[b];
    }
    else
    {
      e0 = fc[b];
      e1 = fc[a];
    }
    if (e0 < fcts_sorted[0])
      return e1;
    if (e1 > fcts_sorted[nfcts - 1])
      return e0;
    if (std::count(fcts_sorted.begin(), fcts_sorted.begin() + nfcts, e0))
      return e0;
    return e1;
  }

  // se/Patch.cpp:406-635
  void initialize_patch(int node_i)
  {
    nodei = node_i;
    Links pc = mv.node_to_cell(node_i);
    Links pf = mv.node_to_fct(node_i);
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
        {
          if (fct_ep[0] < 0)
            fct_ep[0] = id;
          else
            fct_ep[1] = id;
        }
        else if (pb.bfct_type(0, id) == essnt_dual)
        {
          if (fct_ef[0] < 0)
            fct_ef[0] = id;
          else
            fct_ef[1] = id;
        }
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
      for (int i = 1; i < nrhs; ++i)
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

    if (is_internal())
      fcts[1] = fct_first;
    else
      fcts[0] = fct_first;

    int lloop = ncells + 1;
    if (type[0] == internal)
    {
      cells[1] = mv.fct_to_cell(fcts[1])[1];
    }
    else
    {
      cells[1] = mv.fct_to_cell(fcts[0])[0];
      int8_t lf = get_fctid_local(fct_first, mv.cell_to_fct(cells[1]));
      fcts_local[0] = lf;
      fcts_local[1] = lf;
      fcts[1] = next_facet(mv.cell_to_fct(cells[1]), lf);
      lloop = ncells;
    }

    for (int a = 1; a < lloop; ++a)
    {
      int32_t fct_a = fcts[a], cell_a = cells[a];
      Links cf = mv.fct_to_cell(fct_a);
      int32_t cell_ap1 = (cf[0] == cell_a) ? cf[1] : cf[0];
      cells[a + 1] = cell_ap1;
      Links fc_ap1 = mv.cell_to_fct(cell_ap1);
      int8_t lf_ap1 = get_fctid_local(fct_a, fc_ap1);
      fcts_local[2 * a] = get_fctid_local(fct_a, mv.cell_to_fct(cell_a));
      fcts_local[2 * a + 1] = lf_ap1;
      inodes_local[a] = node_local(cell_a, nodei);
      fcts[a + 1] = next_facet(fc_ap1, lf_ap1);
    }

    if (is_on_boundary())
    {
      inodes_local[ncells] = node_local(cells[ncells], nodei);
      int8_t lf = get_fctid_local(fcts[ncells], mv.cell_to_fct(cells[ncells]));
      fcts_local[2 * ncells] = lf;
      fcts_local[2 * ncells + 1] = lf;
    }
    else
    {
      cells[0] = cells[ncells];
      cells[ncells + 1] = cells[1];
      inodes_local[0] = inodes_local[ncells];
      inodes_local[ncells + 1] = inodes_local[1];
      fcts[0] = fcts[nfcts];
      fcts_local[0] = fcts_local[2 * nfcts];
      fcts_local[1] = fcts_local[2 * nfcts + 1];
    }
  }

  // se/Patch.hpp:1001-1007
  void fctid_local(int a, int8_t& fl_eam1, int8_t& fl_ea) const
  {
    fl_eam1 = fcts_local[2 * a - 1];
    fl_ea = fcts_local[2 * a];
  }
  // se/Patch.hpp:921-996 (valid arguments only)
  int8_t fctid_local(int fct_i, int cell_i) const
  {
    int offst;
    if (type[0] == internal && (fct_i == 0 || fct_i == ncells))
      offst = (cell_i == 1 || cell_i == ncells + 1) ? 1 : 0;
    else if (fct_i == 0)
      offst = 0;
    else
      offst = (cell_i == fct_i) ? 0 : 1;
    return fcts_local[2 * fct_i + offst];
  }

  // se/Patch.hpp:468-619
  void flux_dofmap_cell(int a)
  {
    const int32_t cell = cells[a];
    int8_t fl_Eam1, fl_Ea;
    fctid_local(a, fl_Eam1, fl_Ea);
    const int32_t gdof = cell * nrt;
    if (k == 1)
    {
      dofmap(0, a, 0) = fl_Eam1;
      dofmap(0, a, 1) = fl_Ea;
      dofmap(1, a, 0) = gdof + fl_Eam1;
      dofmap(1, a, 1) = gdof + fl_Ea;
      dofmap(2, a, 0) = 0;
      dofmap(2, a, 1) = 0;
    }
    else
    {
      int o = offs[1];
      int pdof_Eam1 = (a - 1) * (k - 1), pdof_Ea;
      if (is_internal() && a == ncells)
        pdof_Ea = 0;
      else
        pdof_Ea = pdof_Eam1 + k - 1;
      for (int ii = 0; ii < k; ++ii)
      {
        int l_Eam1 = fl_Eam1 * k + ii, l_Ea = fl_Ea * k + ii;
        dofmap(0, a, ii) = l_Eam1;
        dofmap(0, a, o) = l_Ea;
        dofmap(1, a, ii) = gdof + l_Eam1;
        dofmap(1, a, o) = gdof + l_Ea;
        if (ii == 0)
        {
          dofmap(2, a, ii) = 0;
          dofmap(2, a, o) = 0;
        }
        else
        {
          dofmap(2, a, ii) = pdof_Eam1 + ii;
          dofmap(2, a, o) = pdof_Ea + ii;
        }
        ++o;
      }
      if (k > 2)
      {
        o = offs[2];
        int ldof = 3 * k + ndiv;
        int pdof = nfcts * (k - 1) + 1 + (a - 1) * nadd;
        for (int ii = 0; ii < nadd; ++ii)
        {
          dofmap(0, a, o) = ldof;
          dofmap(1, a, o) = gdof + ldof;
          dofmap(2, a, o) = pdof + ii;
          dofmap(3, a, o) = 1;
          ++o;
          ++ldof;
        }
      }
      o = offs[4];
      int ldof = 3 * k;
      for (int ii = 0; ii < ndiv; ++ii)
      {
        dofmap(0, a, o) = ldof;
        dofmap(1, a, o) = gdof + ldof;
        dofmap(3, a, o) = 0;
        ++o;
        ++ldof;
      }
    }
    if (symconstr)
    {
      if (is_internal())
        fctdofs_constraint_space(a);
      else
      {
        if (a == 1)
        {
          fctdofs_constraint_space_bnd(0, a);
          fctdofs_constraint_space(1);
        }
        if (a == ncells)
          fctdofs_constraint_space_bnd(a, a);
        else
          fctdofs_constraint_space(a);
      }
    }
  }

  // se/Patch.hpp:621-671
  void fctdofs_constraint_space(int a)
  {
    Links nof = mv.fct_to_node(fcts[a]);
    const int pc1 = a, pc2 = (a == ncells) ? 1 : a + 1;
    const int32_t c1 = cells[pc1], c2 = cells[pc2];
    const int o = offs[3];
    for (int32_t node : nof)
    {
      if (node == nodei)
      {
        dofmap(0, pc1, o) = inodes_local[pc1];
        dofmap(0, pc2, o) = inodes_local[pc2];
        dofmap(2, pc1, o) = 0;
        dofmap(2, pc2, o) = 0;
        dofmap(3, pc1, o) = 1;
        dofmap(3, pc2, o) = 1;
      }
      else
      {
        dofmap(0, pc1, o + 1) = node_local(c1, node);
        dofmap(0, pc2, o + 2) = node_local(c2, node);
        dofmap(2, pc1, o + 1) = a;
        dofmap(2, pc2, o + 2) = a;
        dofmap(3, pc1, o + 1) = 1;
        dofmap(3, pc2, o + 2) = 1;
      }
    }
  }

  // se/Patch.hpp:673-708
  void fctdofs_constraint_space_bnd(int pfct, int pcell)
  {
    Links nof = mv.fct_to_node(fcts[pfct]);
    const int32_t cell = cells[pcell];
    int o, pdof;
    if (pfct == 0)
    {
      o = offs[3] + 2;
      pdof = nfcts - 1;
    }
    else
    {
      o = offs[3] + 1;
      pdof = nfcts;
    }
    int32_t node = (nof[0] == nodei) ? nof[1] : nof[0];
    dofmap(0, pcell, o) = node_local(cell, node);
    dofmap(2, pcell, o) = pdof;
    dofmap(3, pcell, o) = 1;
  }

  // se/Patch.hpp:792-898
  void create_subdofmap(int node_i)
  {
    initialize_patch(node_i);
    ndof_min_flux = 1 + (k - 1) * nfcts + nadd * ncells;
    // NOTE: the reference keeps stale entries of the previous patch in _ddofmap;
    // they are never read. Zero-filling makes the map comparable bit by bit.
    std::fill(ddofmap.begin(), ddofmap.end(), 0);
    std::fill(list_fctdofs_fluxdg.begin(), list_fctdofs_fluxdg.end(), 0);
    const int32_t* closure = pb.t->fct_closure;
    for (int a = 1; a < ncells + 1; ++a)
    {
      flux_dofmap_cell(a);
      if (k == 1 || pb.t->p == 0)
      {
        dofs_fluxdg(a, 0) = 0;
        dofs_fluxdg(a, 1) = 0;
      }
      else
      {
        int8_t lf_eam1, lf_ea;
        fctid_local(a, lf_eam1, lf_ea);
        for (int i = 0; i < ndg_fct; ++i)
        {
          dofs_fluxdg(a - 1, i) = closure[lf_eam1 * ndg_fct + i];
          dofs_fluxdg(a, ndg_fct + i) = closure[lf_ea * ndg_fct + i];
        }
      }
    }
    if (is_internal())
    {
      for (int ii = 0; ii < ndpc; ++ii)
        for (int pl = 0; pl < 3; ++pl)
        {
          dofmap(pl, 0, ii) = dofmap(pl, ncells, ii);
          dofmap(pl, ncells + 1, ii) = dofmap(pl, 1, ii);
        }
      for (int ii = 0; ii < ndg_fct; ++ii)
      {
        dofs_fluxdg(ncells, ii) = dofs_fluxdg(0, ii);
        dofs_fluxdg(0, ndg_fct + ii) = dofs_fluxdg(ncells, ndg_fct + ii);
      }
    }
    else
    {
      for (int ii = 0; ii < ndg_fct; ++ii)
      {
        dofs_fluxdg(0, ndg_fct + ii) = dofs_fluxdg(0, ii);
        dofs_fluxdg(ncells, ii) = dofs_fluxdg(ncells, ndg_fct + ii);
      }
    }
  }

  // se/Patch.hpp:710-789
  void set_assembly_informations(const bool fct_out[3], const uint8_t* reversed /*[ncells][2]*/,
                                 const double* detJ)
  {
    for (int a = 1; a < ncells + 1; ++a)
    {
      const int id_a = a - 1;
      int8_t fl_eam1, fl_ea;
      fctid_local(a, fl_eam1, fl_ea);
      int p_eam1, p_ea;
      if (detJ[id_a] < 0)
      {
        if (reversed[2 * id_a])
          p_eam1 = dofmap(3, a - 1, k);
        else
          p_eam1 = fct_out[fl_eam1] ? -1 : 1;
        p_ea = fct_out[fl_ea] ? 1 : -1;
      }
      else
      {
        if (reversed[2 * id_a])
          p_eam1 = dofmap(3, a - 1, k);
        else
          p_eam1 = fct_out[fl_eam1] ? 1 : -1;
        p_ea = fct_out[fl_ea] ? -1 : 1;
      }
      for (int i = 0; i < k; ++i)
      {
        dofmap(3, a, i) = p_eam1;
        dofmap(3, a, k + i) = p_ea;
      }
    }
    if (is_internal())
    {
      if (reversed[0])
        for (int i = 0; i < k; ++i)
          dofmap(3, 1, i) = dofmap(3, ncells, k + i);
      for (int ii = 0; ii < ndpc; ++ii)
      {
        dofmap(3, 0, ii) = dofmap(3, ncells, ii);
        dofmap(3, ncells + 1, ii) = dofmap(3, 1, ii);
      }
    }
  }

  double estimate_squared_korn_constant() const;
};

} // namespace oracle
