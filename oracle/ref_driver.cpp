// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
// C entry points of `oracle/_ref/libeqlb_ref.so`: the reference's OWN sources
// (/root/reference/cpp/dolfinx_eqlb: se/Patch.cpp, se/KernelData.cpp, base/KernelData.cpp,
// base/BoundaryData.cpp, ev/Patch.cpp and the SE header templates se/reconstruction.hpp ->
// solve_patch_semiexplt.hpp, assembly.hpp, fluxmin_kernel.hpp, stressmin_kernel.hpp,
// solve_patch_weaksym.hpp, PatchData.hpp) compiled UNCHANGED from where they lie against the
// stand-in headers in oracle/ref_shim (DOLFINx / Basix / Eigen are not in this image).
// This file only marshals the plain arrays of the C ABI (`eqlb_mesh`) into the stand-in
// objects and calls the reference's entry points - it contains no equilibration logic.
// Only tests/ and bench.py's reference arm may load the library.
#include <dolfinx_eqlb/base/BoundaryData.hpp>
#include <dolfinx_eqlb/base/FluxBC.hpp>
#include <dolfinx_eqlb/ev/Patch.hpp>
#include <dolfinx_eqlb/base/local_solver.hpp>
#include <dolfinx_eqlb/ev/reconstruction.hpp>
#include <dolfinx_eqlb/se/reconstruction.hpp>

#include "../include/eqlb_b200.h"

#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace dolfinx;
namespace eqlb = dolfinx_eqlb;

extern "C"
{
/// the flux element as produced by the reference's `create_hierarchic_rt` (tests/golden/ref_rt_element.npz)
typedef struct ref_element
{
  int32_t k;          // Basix degree of the RT element
  int32_t nrt, nmono; // k(k+2), (k+1)(k+2)/2
  const double* coef; // [nrt][2][nmono]
  int32_t npts, mcols;
  const double* X; // [npts][2]
  const double* M; // [nrt][mcols]
  int32_t p;       // degree of the projected flux / RHS (DG_p)
} ref_element;
}

namespace
{
thread_local std::string g_err;

using AL = graph::AdjacencyList<std::int32_t>;

std::shared_ptr<const AL> csr(const int32_t* off, const int32_t* data, int n)
{
  return std::make_shared<const AL>(std::vector<std::int32_t>(data, data + off[n]), std::vector<std::int32_t>(off, off + n + 1));
}

std::shared_ptr<const mesh::Mesh> make_mesh(const eqlb_mesh* m)
{
  if (m->node_owned)
    for (int i = 0; i < m->nnode; ++i)
      if (!m->node_owned[i])
        throw std::runtime_error("ref: partitioned meshes (node_owned) are not supported by the reference build");
  mesh::Topology t;
  t.set_connectivity(csr(m->node_cell_off, m->node_cell, m->nnode), 0, 2);
  t.set_connectivity(csr(m->node_fct_off, m->node_fct, m->nnode), 0, 1);
  t.set_connectivity(csr(m->fct_cell_off, m->fct_cell, m->nfct), 1, 2);
  t.set_connectivity(std::make_shared<const AL>(AL::regular(m->fct_node, m->nfct, 2)), 1, 0);
  t.set_connectivity(std::make_shared<const AL>(AL::regular(m->cell_fct, m->ncell, 3)), 2, 1);
  t.set_connectivity(std::make_shared<const AL>(AL::regular(m->cell_node, m->ncell, 3)), 2, 0);
  t.set_index_map(0, std::make_shared<const common::IndexMap>(m->nnode));
  t.set_index_map(1, std::make_shared<const common::IndexMap>(m->nfct));
  t.set_index_map(2, std::make_shared<const common::IndexMap>(m->ncell));
  t.set_permutations(std::vector<std::uint8_t>(m->fct_perms, m->fct_perms + 3 * (size_t)m->ncell),
                     m->cell_perm_info ? std::vector<std::uint32_t>(m->cell_perm_info, m->cell_perm_info + m->ncell)
                                       : std::vector<std::uint32_t>(m->ncell, 0));
  mesh::Geometry g(2, AL::regular(m->cell_node, m->ncell, 3), std::vector<double>(m->x, m->x + 3 * (size_t)m->nnode));
  return std::make_shared<const mesh::Mesh>(std::move(t), std::move(g));
}

basix::FiniteElement make_rt(const ref_element* e, bool discontinuous)
{
  const int k = e->k, nrt = e->nrt;
  std::vector<std::vector<std::vector<int>>> ed(3), ec(3);
  ed[0].resize(3);
  ed[1].resize(3);
  ed[2].resize(1);
  if (discontinuous)
  {
    for (int i = 0; i < nrt; ++i)
      ed[2][0].push_back(i);
  }
  else
  {
    for (int f = 0; f < 3; ++f)
      for (int j = 0; j < k; ++j)
        ed[1][f].push_back(f * k + j);
    for (int i = 3 * k; i < nrt; ++i)
      ed[2][0].push_back(i);
  }
  ec = ed;
  return basix::FiniteElement::from_monomials(
      basix::cell::type::triangle, k, k, 2, nrt, std::vector<double>(e->coef, e->coef + (size_t)nrt * 2 * e->nmono), discontinuous,
      basix::element::lagrange_variant::equispaced, basix::maps::type::contravariantPiola,
      std::vector<double>(e->X, e->X + 2 * (size_t)e->npts), {(size_t)e->npts, 2},
      std::vector<double>(e->M, e->M + (size_t)nrt * e->mcols), {(size_t)nrt, (size_t)e->mcols}, ed, ec);
}

/// the function spaces of FluxEqlbSE (`FluxEqlbSE.py:62-81`): DRT_k flux, DG_p^2 projected flux, DG_p RHS
struct Spaces
{
  std::shared_ptr<const mesh::Mesh> mesh;
  std::shared_ptr<fem::FunctionSpace> V_hdiv, V_dg2, V_dg;
  int nrt, ndg;

  Spaces(const eqlb_mesh* m, const ref_element* e)
  {
    mesh = make_mesh(m);
    basix::FiniteElement rt = make_rt(e, true);
    nrt = e->nrt;
    basix::FiniteElement dg
        = basix::element::create_lagrange(basix::cell::type::triangle, e->p, basix::element::lagrange_variant::equispaced, true);
    ndg = dg.dim();
    // DRT: identity layout cell*nrt + i (`se/Patch.hpp:480`)
    std::vector<std::int32_t> ident((size_t)m->ncell * nrt);
    for (size_t i = 0; i < ident.size(); ++i)
      ident[i] = (std::int32_t)i;
    auto dm_rt = std::make_shared<const fem::DofMap>(AL::regular(ident.data(), m->ncell, nrt), m->ncell * nrt, 1,
                                                     fem::ElementDofLayout(rt.entity_dofs()));
    V_hdiv = std::make_shared<fem::FunctionSpace>(mesh, std::make_shared<const fem::FiniteElement>(rt, 1), dm_rt);
    auto dm_dg2 = std::make_shared<const fem::DofMap>(AL::regular(m->dg_dofmap, m->ncell, ndg), m->ncell * ndg, 2,
                                                      fem::ElementDofLayout(dg.entity_dofs()));
    V_dg2 = std::make_shared<fem::FunctionSpace>(mesh, std::make_shared<const fem::FiniteElement>(dg, 2), dm_dg2);
    auto dm_dg = std::make_shared<const fem::DofMap>(AL::regular(m->dg_dofmap, m->ncell, ndg), m->ncell * ndg, 1,
                                                     fem::ElementDofLayout(dg.entity_dofs()));
    V_dg = std::make_shared<fem::FunctionSpace>(mesh, std::make_shared<const fem::FiniteElement>(dg, 1), dm_dg);
  }
};

/// `base::BoundaryData` in the state its constructor leaves it in (`base/BoundaryData.cpp:279-633`),
/// with the flux-BC part (facet types, cell-local facet ids, boundary DOFs, node markers)
/// injected from the arrays of the C ABI instead of being evaluated from FluxBC kernels.
class InjectedBoundaryData : public eqlb::base::BoundaryData<double>
{
public:
  InjectedBoundaryData(std::vector<std::vector<std::shared_ptr<eqlb::base::FluxBC<double>>>>& no_bcs,
                       std::vector<std::shared_ptr<fem::Function<double>>>& bfuncs, std::shared_ptr<const fem::FunctionSpace> V,
                       int qdegree, const std::vector<std::vector<std::int32_t>>& fct_prime, bool stress, const eqlb_mesh* m,
                       int nrhs, const int8_t* facet_type, const int8_t* local_fct_id, const int8_t* node_on_bnd)
      : eqlb::base::BoundaryData<double>(no_bcs, bfuncs, V, true, qdegree, fct_prime, stress)
  {
    const int k = V->element()->basix_element().degree();
    for (int r = 0; r < nrhs; ++r)
    {
      std::span<std::int8_t> ft = this->facet_type(r);
      std::span<std::int8_t> bm = this->boundary_markers(r);
      for (int f = 0; f < m->nfct; ++f)
      {
        ft[f] = facet_type[(size_t)r * m->nfct + f];
        if (ft[f] == eqlb::base::PatchFacetType::essnt_dual)
        {
          const std::int32_t c = m->fct_cell[m->fct_cell_off[f]];
          std::int8_t fl = 0;
          for (int j = 0; j < 3; ++j)
            if (m->cell_fct[3 * c + j] == f)
              fl = j;
          if (local_fct_id && local_fct_id[f] != fl)
            throw std::runtime_error("ref: local_fct_id inconsistent with the mesh");
          this->_local_fct_id[f] = fl;
          std::vector<std::int32_t> bd(k);
          this->boundary_dofs(c, fl, bd);
          for (int i = 0; i < k; ++i)
            bm[bd[i]] = true;
        }
      }
    }
    if (stress && node_on_bnd)
      for (int i = 0; i < m->nnode; ++i)
        this->_pnt_on_esnt_boundary[i] = node_on_bnd[i];
  }
};

struct Problem
{
  Spaces sp;
  int nrhs;
  std::vector<std::vector<double>> zero_bflux;
  std::vector<std::shared_ptr<fem::Function<double>>> bfuncs;
  std::vector<std::vector<std::shared_ptr<eqlb::base::FluxBC<double>>>> no_bcs;
  std::vector<std::vector<std::int32_t>> fct_prime;
  std::shared_ptr<InjectedBoundaryData> bdata;

  Problem(const eqlb_mesh* m, const ref_element* e, int nrhs_, const int8_t* facet_type, const double* const* bflux,
          const int8_t* local_fct_id, const int8_t* node_on_bnd, bool stress)
      : sp(m, e), nrhs(nrhs_), no_bcs(nrhs_), fct_prime(nrhs_)
  {
    const size_t n = (size_t)m->ncell * sp.nrt;
    zero_bflux.resize(nrhs);
    for (int r = 0; r < nrhs; ++r)
    {
      // the boundary function of rhs r (DRT layout); the reference holds it as a fem::Function
      zero_bflux[r].assign(n, 0.0);
      if (bflux && bflux[r])
        std::memcpy(zero_bflux[r].data(), bflux[r], n * sizeof(double));
      bfuncs.push_back(
          std::make_shared<fem::Function<double>>(sp.V_hdiv, std::make_shared<la::Vector<double>>(zero_bflux[r].data(), n)));
      for (int f = 0; f < m->nfct; ++f)
        if (facet_type[(size_t)r * m->nfct + f] == eqlb::base::PatchFacetType::essnt_primal)
          fct_prime[r].push_back(f);
    }
    // quadrature degree of the (unused here) projection branch: `bcs.py` default 2k
    bdata = std::make_shared<InjectedBoundaryData>(no_bcs, bfuncs, sp.V_hdiv, 2 * e->k, fct_prime, stress, m, nrhs, facet_type,
                                                   local_fct_id, node_on_bnd);
  }
};

template <typename F>
int guarded(F&& f)
{
  try
  {
    f();
    return 0;
  }
  catch (const std::exception& ex)
  {
    g_err = ex.what();
    return -1;
  }
}
} // namespace

namespace
{
/// ID = the reference's dispatch `se/reconstruction.hpp:393-406` (1, 2, 3 = general)
template <int ID>
void se_patch_maps_impl(const eqlb_mesh* mesh, const ref_element* elmt, int nrhs, const int8_t* facet_type,
                        const int8_t* node_on_bnd, int stress, int ncmax, int32_t* ncells, int32_t* cells, int32_t* fcts,
                        int8_t* inodes_local, int8_t* fcts_local, int8_t* type, uint8_t* reversed, uint8_t* reversion,
                        int32_t* dofmap, int32_t* projflux_fct, int8_t* bmarkers)
{
        using T = double;
        Problem pb(mesh, elmt, nrhs, facet_type, nullptr, nullptr, node_on_bnd, stress != 0);
        const int k = elmt->k;
        const size_t np = mesh->nnode, nc_all = mesh->ncell;
        // zero data functions
        std::vector<std::vector<double>> zs(nrhs, std::vector<double>(nc_all * pb.sp.nrt, 0.0));
        std::vector<double> zG(nc_all * pb.sp.ndg * 2, 0.0), zF(nc_all * pb.sp.ndg, 0.0);
        std::vector<std::shared_ptr<fem::Function<double>>> flux_hdiv, flux_dg, rhs_dg;
        for (int r = 0; r < nrhs; ++r)
        {
          flux_hdiv.push_back(std::make_shared<fem::Function<double>>(
              pb.sp.V_hdiv, std::make_shared<la::Vector<double>>(zs[r].data(), zs[r].size())));
          flux_dg.push_back(
              std::make_shared<fem::Function<double>>(pb.sp.V_dg2, std::make_shared<la::Vector<double>>(zG.data(), zG.size())));
          rhs_dg.push_back(
              std::make_shared<fem::Function<double>>(pb.sp.V_dg, std::make_shared<la::Vector<double>>(zF.data(), zF.size())));
        }
        std::shared_ptr<eqlb::base::BoundaryData<double>> bd = pb.bdata;
        eqlb::se::ProblemData<T> problem_data(flux_hdiv, flux_dg, rhs_dg, bd);

        // the set-up block of se::reconstruction (se/reconstruction.hpp:78-153)
        auto msh = problem_data.mesh();
        const std::vector<std::uint8_t>& fct_perms = msh->topology().get_facet_permutations();
        const basix::FiniteElement& el_hdiv = problem_data.fspace_flux_hdiv()->element()->basix_element();
        const basix::FiniteElement& el_rhs = problem_data.fspace_flux_dg()->element()->basix_element();
        const int degree_rhs = el_rhs.degree();
        basix::FiniteElement el_rhscg
            = basix::element::create_lagrange(el_rhs.cell_type(), degree_rhs, el_rhs.lagrange_variant(), degree_rhs == 0);
        basix::FiniteElement el_hat = basix::element::create_lagrange(el_rhs.cell_type(), 1, el_rhs.lagrange_variant(), false);
        const int qdeg = (k == 1) ? 2 : 2 * k + 1;
        auto qr = std::make_shared<eqlb::base::QuadratureRule>(msh->topology().cell_type(), qdeg, 2);
        eqlb::se::KernelData<T> kernel_data(msh, qr, el_hdiv, el_rhs, el_hat);
        auto kmin = eqlb::se::generate_flux_minimisation_kernel<T, true>(kernel_data, 2, k - 1);
        auto kmin_l = eqlb::se::generate_flux_minimisation_kernel<T, false>(kernel_data, 2, k - 1);

        const int nadd = (k - 1) * (k - 2) / 2, ndiv = k * (k + 1) / 2 - 1;
        const int ndpc = 2 * k + nadd + ndiv + (stress ? 3 : 0);
        const int hzmax = 1 + (k - 1) * (ncmax + 1) + nadd * ncmax;
        int nf = 0;
        if (cells) std::fill(cells, cells + np * (ncmax + 2), -1);
        if (fcts) std::fill(fcts, fcts + np * (ncmax + 2), -1);
        if (inodes_local) std::fill(inodes_local, inodes_local + np * (ncmax + 2), (int8_t)-1);
        if (fcts_local) std::fill(fcts_local, fcts_local + np * 2 * (ncmax + 1), (int8_t)-1);
        if (reversed) std::fill(reversed, reversed + np * ncmax * 2, (uint8_t)255);
        if (dofmap) std::fill(dofmap, dofmap + np * 4 * (ncmax + 2) * ndpc, -1);
        if (bmarkers) std::fill(bmarkers, bmarkers + np * nrhs * hzmax, (int8_t)-1);

        for (size_t z = 0; z < np; ++z)
        {
          eqlb::se::Patch<T, ID> patch(msh, problem_data.facet_type(), problem_data.node_on_essnt_boundary_stress(),
                                       problem_data.fspace_flux_hdiv(), problem_data.fspace_flux_dg(), el_rhscg, stress != 0, 1,
                                       stress ? 2 : 1);
          eqlb::se::PatchData<T, ID> patch_data(patch, kernel_data.nipoints_facet(), stress != 0);
          if (patch.ncells_max() > ncmax)
            throw std::runtime_error("ncmax too small");
          nf = patch.ndofs_fluxdg_fct();
          if (z == 0 && projflux_fct)
            std::fill(projflux_fct, projflux_fct + np * (ncmax + 1) * 2 * nf, -1);
          patch.create_subdofmap((int)z);
          patch_data.reinitialisation(patch.type(), patch.ncells());
          eqlb::se::equilibrate_flux_semiexplt<T, ID>(msh->geometry(), fct_perms, patch, patch_data, problem_data, kernel_data,
                                                      kmin, kmin_l);
          const int nc = patch.ncells();
          if (ncells) ncells[z] = nc;
          auto pc = patch.cells();
          auto pf = patch.fcts();
          if (cells)
            for (int a = patch.is_internal() ? 0 : 1; a < (int)pc.size(); ++a)
              cells[z * (ncmax + 2) + a] = pc[a];
          if (fcts)
            for (int a = 0; a < (int)pf.size(); ++a)
              fcts[z * (ncmax + 2) + a] = pf[a];
          if (inodes_local)
            for (int a = patch.is_internal() ? 0 : 1; a < (int)pc.size(); ++a)
              inodes_local[z * (ncmax + 2) + a] = patch.inode_local(a);
          if (fcts_local)
          {
            auto fl = patch.eqlb::se::OrientedPatch::fctid_local();
            const int n = 2 * patch.nfcts() + (patch.is_internal() ? 2 : 0);
            for (int a = 0; a < n; ++a)
              fcts_local[z * 2 * (ncmax + 1) + a] = fl[a];
          }
          if (type)
            for (int i = 0; i < nrhs; ++i)
              type[z * nrhs + i] = (int8_t)patch.type(i);
          if (reversion)
            for (int i = 0; i < nrhs; ++i)
              reversion[z * nrhs + i] = patch.reversion_required(i);
          if (reversed)
          {
            auto rv = patch_data.reversed_facets_per_cell();
            for (int a = 0; a < nc; ++a)
              for (int j = 0; j < 2; ++j)
                reversed[z * ncmax * 2 + 2 * a + j] = rv(a, j);
          }
          if (dofmap)
          {
            auto dm = patch.assembly_info_minimisation();
            if ((int)dm.extent(2) != ndpc)
              throw std::runtime_error("ref: unexpected DOF-map width");
            for (int pl = 0; pl < 4; ++pl)
              for (int a = 0; a < nc + 2; ++a)
                for (int i = 0; i < ndpc; ++i)
                  dofmap[((z * 4 + pl) * (ncmax + 2) + a) * ndpc + i] = dm(pl, a, i);
          }
          if (projflux_fct)
            for (int a = 0; a < nc + 1; ++a)
            {
              auto d = patch.dofs_projflux_fct(a);
              for (int i = 0; i < 2 * nf; ++i)
                projflux_fct[(z * (ncmax + 1) + a) * 2 * nf + i] = d[i];
            }
          if (bmarkers)
          {
            const int hz = patch.ndofs_flux_hdiz_zero();
            for (int i = 0; i < nrhs; ++i)
            {
              std::vector<int8_t> bm(hz, 0);
              const auto tp = patch.type(i);
              if (tp == eqlb::base::PatchType::bound_essnt_dual || tp == eqlb::base::PatchType::bound_mixed)
                eqlb::se::set_boundary_markers(std::span<std::int8_t>(bm), {tp}, {patch.reversion_required(i)}, nc, hz, patch.ndofs_flux_fct());
              for (int j = 0; j < hz; ++j)
                bmarkers[(z * nrhs + i) * hzmax + j] = bm[j];
            }
          }
        }
}
} // namespace

extern "C"
{
const char* ref_last_error() { return g_err.c_str(); }

/// test hook: complement of basix::cell::facet_orientations (see ref_shim/basix/finite-element.h)
void ref_set_flip_orientations(int flip) { basix::cell::flip_orientations_flag() = flip != 0; }

/// `reconstruct_fluxes_semiexplt[_with_kornconst]` (wrappers.cpp:97-137) -> se::reconstruction
/// (se/reconstruction.hpp:317-407): sigma accumulated in place, arguments as oracle_se_run.
int ref_se_run(const eqlb_mesh* mesh, const ref_element* elmt, int nrhs, const int8_t* facet_type, const double* const* bflux,
               const int8_t* local_fct_id, const int8_t* node_on_bnd, int stress, const double* const* G, const double* const* F,
               double* const* sigma, double* korn)
{
  return guarded(
      [&]
      {
        Problem pb(mesh, elmt, nrhs, facet_type, bflux, local_fct_id, node_on_bnd, stress != 0);
        const size_t nc = mesh->ncell;
        std::vector<std::shared_ptr<fem::Function<double>>> flux_hdiv, flux_dg, rhs_dg;
        for (int r = 0; r < nrhs; ++r)
        {
          flux_hdiv.push_back(std::make_shared<fem::Function<double>>(
              pb.sp.V_hdiv, std::make_shared<la::Vector<double>>(sigma[r], nc * pb.sp.nrt)));
          flux_dg.push_back(std::make_shared<fem::Function<double>>(
              pb.sp.V_dg2, std::make_shared<la::Vector<double>>(const_cast<double*>(G[r]), nc * pb.sp.ndg * 2)));
          rhs_dg.push_back(std::make_shared<fem::Function<double>>(
              pb.sp.V_dg, std::make_shared<la::Vector<double>>(const_cast<double*>(F[r]), nc * pb.sp.ndg)));
        }
        std::shared_ptr<fem::Function<double>> kc;
        if (korn)
          kc = std::make_shared<fem::Function<double>>(pb.sp.V_dg, std::make_shared<la::Vector<double>>(korn, nc));
        std::shared_ptr<eqlb::base::BoundaryData<double>> bd = pb.bdata;
        eqlb::se::reconstruction<double>(flux_hdiv, flux_dg, rhs_dg, bd, stress != 0, kc);
      });
}

/// Integer maps of every patch, produced by the reference's se::Patch (`se/Patch.cpp:406-635`
/// initialize_patch, `se/Patch.hpp:792-898` create_subdofmap, `:710-789`
/// set_assembly_informations) and by one pass of `equilibrate_flux_semiexplt` over zero data
/// (reversed-facet flags `se/solve_patch_semiexplt.hpp:325-389`, boundary markers
/// `se/assembly.hpp:46-98`).  Output layout = oracle_se_patch_maps / eqlb_get_patch_maps.
/// A fresh Patch object is built per node so that no entries of a previous patch survive.
int ref_se_patch_maps(const eqlb_mesh* mesh, const ref_element* elmt, int nrhs, const int8_t* facet_type,
                      const int8_t* node_on_bnd, int stress, int ncmax, int32_t* ncells, int32_t* cells, int32_t* fcts,
                      int8_t* inodes_local, int8_t* fcts_local, int8_t* type, uint8_t* reversed, uint8_t* reversion,
                      int32_t* dofmap, int32_t* projflux_fct, int8_t* bmarkers)
{
  return guarded(
      [&]
      {
        auto run = [&](auto id)
        {
          se_patch_maps_impl<decltype(id)::value>(mesh, elmt, nrhs, facet_type, node_on_bnd, stress, ncmax, ncells, cells, fcts, inodes_local,
                                                  fcts_local, type, reversed, reversion, dofmap, projflux_fct, bmarkers);
        };
        if (elmt->k == 1)
          run(std::integral_constant<int, 1>{});
        else if (elmt->k == 2)
          run(std::integral_constant<int, 2>{});
        else
          run(std::integral_constant<int, 3>{});
      });
}

/// EV patch ordering and sub-DOF maps of one patch from the reference's ev::Patch
/// (`ev/Patch.cpp:83-309` initialize_patch / fcti_to_celli, `:482-676` create_subdofmap).
/// Mixed RT_k x DG_(k-1) dofmap in the numbering of include/eqlb_b200.h (eqlb_get_ev_dofmaps):
/// facet*k+j | nfct*k + cell*(k*k-k) + i | nflux + cell*ndg + q.  Output as oracle_ev_patch_maps.
int ref_ev_patch_maps(const eqlb_mesh* mesh, const ref_element* elmt, int nrhs, const int8_t* facet_type, int node,
                      int32_t* ncells, int32_t* cells, int32_t* fcts, int8_t* inodes_local, int32_t* dofs_elmt,
                      int32_t* dofs_patch, int32_t* dofs_global, int32_t* list_patch, int32_t* list_global)
{
  return guarded(
      [&]
      {
        auto msh = make_mesh(mesh);
        const int k = elmt->k, nrt = elmt->nrt;
        basix::FiniteElement rt = make_rt(elmt, false);
        basix::FiniteElement dg
            = basix::element::create_lagrange(basix::cell::type::triangle, elmt->p, basix::element::lagrange_variant::equispaced, true);
        const int ndg = dg.dim(), nel = nrt + ndg;
        const std::int32_t nflux = mesh->nfct * k + mesh->ncell * (nrt - 3 * k);
        std::vector<std::int32_t> mixed((size_t)mesh->ncell * nel), flux((size_t)mesh->ncell * nrt);
        for (std::int32_t c = 0; c < mesh->ncell; ++c)
          for (int l = 0; l < nel; ++l)
          {
            std::int32_t g;
            if (l < 3 * k)
              g = mesh->cell_fct[3 * c + l / k] * k + l % k;
            else if (l < nrt)
              g = mesh->nfct * k + c * (nrt - 3 * k) + (l - 3 * k);
            else
              g = nflux + c * ndg + (l - nrt);
            mixed[(size_t)c * nel + l] = g;
            if (l < nrt)
              flux[(size_t)c * nrt + l] = g;
          }
        auto el_flux = std::make_shared<const fem::FiniteElement>(rt, 1);
        auto el_mixed = std::make_shared<fem::FiniteElement>(rt, 1);
        el_mixed->set_space_dimension(nel);
        auto V_flux = std::make_shared<fem::FunctionSpace>(
            msh, el_flux,
            std::make_shared<const fem::DofMap>(AL::regular(flux.data(), mesh->ncell, nrt), nflux, 1,
                                                fem::ElementDofLayout(rt.entity_dofs())));
        auto V = std::make_shared<fem::FunctionSpace>(
            msh, el_mixed,
            std::make_shared<const fem::DofMap>(AL::regular(mixed.data(), mesh->ncell, nel), nflux + mesh->ncell * ndg, 1,
                                                fem::ElementDofLayout(rt.entity_dofs())));
        eqlb::base::mdspan_t<const std::int8_t, 2> ft(facet_type, (std::size_t)nrhs, (std::size_t)mesh->nfct);
        eqlb::ev::Patch patch(mesh->nnode, msh, ft, V, V_flux, rt);
        patch.create_subdofmap(node);
        const int nc = patch.ncells();
        *ncells = nc;
        for (int i = 0; i < nc; ++i)
        {
          cells[i] = patch.cell(i);
          inodes_local[i] = patch.inode_local(i);
        }
        for (int i = 0; i < patch.nfcts(); ++i)
          fcts[i] = patch.fct(i);
        const int nz = patch.ndofs_elmt_nz();
        for (int c = 0; c < nc; ++c)
        {
          auto de = patch.dofs_elmt(c);
          auto dp = patch.dofs_patch(c);
          auto dgl = patch.dofs_global(c);
          if ((int)de.size() != nz)
            throw std::runtime_error("ref: unexpected EV sub-DOFmap width");
          for (int i = 0; i < nz; ++i)
          {
            dofs_elmt[c * nz + i] = de[i];
            dofs_patch[c * nz + i] = dp[i];
            dofs_global[c * nz + i] = dgl[i];
          }
        }
        auto lp = patch.dofs_fluxhdiv_patch();
        auto lg = patch.dofs_fluxhdiv_global();
        for (std::size_t i = 0; i < lp.size(); ++i)
        {
          list_patch[i] = lp[i];
          list_global[i] = lg[i];
        }
      });
}
}

namespace
{
/// The three fixed forms of FluxEqlbEV (`python/dolfinx_eqlb/eqlb/FluxEqlbEV.py:116-133`) as cell kernels
/// with the FFCx `tabulate_tensor` signature.  FFCx itself is third party and absent; these are the
/// integrals as written in the UFL forms, evaluated by quadrature of the Piola-mapped basis on the
/// (affine) cell, WITHOUT DOF transformations (the reference applies those itself, `ev/assembly.hpp:185-198`):
///   a      = (sig, v) - (r, div v) + (div sig, q)                  A[test i][trial j], mixed RT_k x DG_(k-1)
///   l_pen  = (1, q)
///   l      = (hat G, v) + (hat f + grad(hat) . G, q)               coefficients packed as [G (ndg x 2) | f (ndg) | hat (3)]
struct EvForms
{
  int nrt, ndg, nd, nq;
  std::vector<double> w, rt, rtdiv, dg, hat; // rt [q][i][2], rtdiv [q][i], dg [q][c], hat [q][3]

  EvForms(const basix::FiniteElement& el_rt, const basix::FiniteElement& el_dg, int k)
  {
    nrt = el_rt.dim();
    ndg = el_dg.dim();
    nd = nrt + ndg;
    auto quad = basix::quadrature::make_quadrature(basix::cell::type::triangle, 2 * k + 1);
    const std::vector<double>& pts = quad[0];
    w = quad[1];
    nq = (int)w.size();
    std::vector<double> t((size_t)3 * nq * nrt * 2);
    el_rt.tabulate(1, pts, {(size_t)nq, 2}, t);
    rt.assign(t.begin(), t.begin() + (size_t)nq * nrt * 2);
    rtdiv.resize((size_t)nq * nrt);
    const size_t plane = (size_t)nq * nrt * 2;
    for (int q = 0; q < nq; ++q)
      for (int i = 0; i < nrt; ++i)
        rtdiv[(size_t)q * nrt + i] = t[plane + ((size_t)q * nrt + i) * 2] + t[2 * plane + ((size_t)q * nrt + i) * 2 + 1];
    dg.resize((size_t)nq * ndg);
    el_dg.tabulate(0, pts, {(size_t)nq, 2}, dg);
    hat.resize((size_t)nq * 3);
    for (int q = 0; q < nq; ++q)
    {
      hat[3 * q] = 1.0 - pts[2 * q] - pts[2 * q + 1];
      hat[3 * q + 1] = pts[2 * q];
      hat[3 * q + 2] = pts[2 * q + 1];
    }
  }

  static double jac(const double* x, double J[4], double K[4])
  {
    J[0] = x[3] - x[0];
    J[1] = x[6] - x[0];
    J[2] = x[4] - x[1];
    J[3] = x[7] - x[1];
    const double det = J[0] * J[3] - J[1] * J[2];
    K[0] = J[3] / det;
    K[1] = -J[1] / det;
    K[2] = -J[2] / det;
    K[3] = J[0] / det;
    return det;
  }

  void kernel_a(double* A, const double* x) const
  {
    double J[4], K[4];
    const double det = jac(x, J, K);
    std::vector<double> phi((size_t)nrt * 2);
    for (int q = 0; q < nq; ++q)
    {
      const double dv = w[q] * std::fabs(det);
      for (int i = 0; i < nrt; ++i)
      {
        const double r0 = rt[((size_t)q * nrt + i) * 2], r1 = rt[((size_t)q * nrt + i) * 2 + 1];
        phi[2 * i] = (J[0] * r0 + J[1] * r1) / det;
        phi[2 * i + 1] = (J[2] * r0 + J[3] * r1) / det;
      }
      for (int i = 0; i < nrt; ++i)
      {
        const double divi = rtdiv[(size_t)q * nrt + i] / det;
        for (int j = 0; j < nrt; ++j)
          A[i * nd + j] += dv * (phi[2 * i] * phi[2 * j] + phi[2 * i + 1] * phi[2 * j + 1]);
        for (int c = 0; c < ndg; ++c)
        {
          const double psi = dg[(size_t)q * ndg + c];
          A[i * nd + nrt + c] -= dv * psi * divi;
          A[(nrt + c) * nd + i] += dv * divi * psi;
        }
      }
    }
  }

  void kernel_lpen(double* P, const double* x) const
  {
    double J[4], K[4];
    const double det = jac(x, J, K);
    for (int q = 0; q < nq; ++q)
      for (int c = 0; c < ndg; ++c)
        P[c] += w[q] * std::fabs(det) * dg[(size_t)q * ndg + c];
  }

  void kernel_l(double* L, const double* cf, const double* x) const
  {
    double J[4], K[4];
    const double det = jac(x, J, K);
    const double* G = cf;
    const double* f = cf + 2 * ndg;
    const double* h = cf + 3 * ndg;
    // grad(hat) = K^T grad_ref(hat)
    const double gr[3][2] = {{-1, -1}, {1, 0}, {0, 1}};
    double gh[2] = {0, 0};
    for (int n = 0; n < 3; ++n)
    {
      gh[0] += h[n] * (K[0] * gr[n][0] + K[2] * gr[n][1]);
      gh[1] += h[n] * (K[1] * gr[n][0] + K[3] * gr[n][1]);
    }
    for (int q = 0; q < nq; ++q)
    {
      const double dv = w[q] * std::fabs(det);
      double hq = 0, Gq[2] = {0, 0}, fq = 0;
      for (int n = 0; n < 3; ++n)
        hq += h[n] * hat[3 * q + n];
      for (int c = 0; c < ndg; ++c)
      {
        const double psi = dg[(size_t)q * ndg + c];
        Gq[0] += G[2 * c] * psi;
        Gq[1] += G[2 * c + 1] * psi;
        fq += f[c] * psi;
      }
      for (int i = 0; i < nrt; ++i)
      {
        const double r0 = rt[((size_t)q * nrt + i) * 2], r1 = rt[((size_t)q * nrt + i) * 2 + 1];
        const double p0 = (J[0] * r0 + J[1] * r1) / det, p1 = (J[2] * r0 + J[3] * r1) / det;
        L[i] += dv * hq * (Gq[0] * p0 + Gq[1] * p1);
      }
      const double s = hq * fq + gh[0] * Gq[0] + gh[1] * Gq[1];
      for (int c = 0; c < ndg; ++c)
        L[nrt + c] += dv * s * dg[(size_t)q * ndg + c];
    }
  }
};

/// DOLFINx' DOF transformations of the conforming hierarchic RT element on reflected facets
/// (`topology().get_facet_permutations()` bit): facet functionals int v.n s^j ds with s -> 1-s and
/// n -> -n give c_glob = R c_loc, R[j][i] = (-1)^(i+1) binom(j, i) (an involution: the matrix of
/// `se/KernelData.cpp:55-64`), hence phi_glob = M phi_loc with M = R^T on the facet block.
struct RtTransform
{
  int k, nrt;
  std::vector<double> R; // [k][k]
  const std::uint8_t* perms;

  RtTransform(int k_, int nrt_, const std::uint8_t* p) : k(k_), nrt(nrt_), R((size_t)k_ * k_, 0.0), perms(p)
  {
    for (int j = 0; j < k; ++j)
    {
      long b = 1;
      for (int i = 0; i <= j; ++i)
      {
        R[j * k + i] = ((i % 2) == 0) ? -(double)b : (double)b;
        b = b * (j - i) / (i + 1);
      }
    }
  }
  /// data (ndofs x bs) <- Mb data, Mb = R^T (transpose = false) or R (transpose = true) on reflected facet blocks
  void pre(const std::span<double>& data, std::int32_t cell, int bs, bool use_R) const
  {
    std::vector<double> tmp((size_t)k * bs);
    for (int f = 0; f < 3; ++f)
      if (perms[3 * cell + f])
      {
        for (int i = 0; i < k; ++i)
          for (int b = 0; b < bs; ++b)
          {
            double s = 0;
            for (int j = 0; j < k; ++j)
              s += (use_R ? R[i * k + j] : R[j * k + i]) * data[(size_t)(f * k + j) * bs + b];
            tmp[(size_t)i * bs + b] = s;
          }
        for (int i = 0; i < k; ++i)
          for (int b = 0; b < bs; ++b)
            data[(size_t)(f * k + i) * bs + b] = tmp[(size_t)i * bs + b];
      }
  }
  /// data (bs x ndofs, row length nd) <- data M^T
  void post_t(const std::span<double>& data, std::int32_t cell, int bs, int nd) const
  {
    std::vector<double> tmp(k);
    for (int f = 0; f < 3; ++f)
      if (perms[3 * cell + f])
        for (int b = 0; b < bs; ++b)
        {
          for (int i = 0; i < k; ++i)
          {
            double s = 0;
            for (int j = 0; j < k; ++j)
              s += data[(size_t)b * nd + f * k + j] * R[j * k + i]; // (data M^T)[i] = sum_j data[j] M[i][j], M = R^T
            tmp[i] = s;
          }
          for (int i = 0; i < k; ++i)
            data[(size_t)b * nd + f * k + i] = tmp[i];
        }
  }
};
} // namespace

extern "C"
{
/// `reconstruct_fluxes_minimisation` (wrappers.cpp:85-95) -> ev::reconstruction (ev/reconstruction.hpp:64-177):
/// the reference's patch loop, ev::Patch, assemble_tangents, apply_lifting, dense partial-pivot LU and the
/// += scatter run unchanged; forms and DOF transformations as described above.  sigma: conforming
/// hierarchic-RT vectors [nfct*k + ncell*(k*k-k)], accumulated; bflux as for ref_se_run (DRT layout,
/// cell-local moments).  Arguments as oracle_ev_run.
int ref_ev_run(const eqlb_mesh* mesh, const ref_element* elmt, int nrhs, const int8_t* facet_type, const double* const* bflux,
               const double* const* G, const double* const* F, double* const* sigma)
{
  return guarded(
      [&]
      {
        auto msh = make_mesh(mesh);
        const int k = elmt->k, nrt = elmt->nrt;
        basix::FiniteElement rt = make_rt(elmt, false);
        basix::FiniteElement dg
            = basix::element::create_lagrange(basix::cell::type::triangle, elmt->p, basix::element::lagrange_variant::equispaced, true);
        basix::FiniteElement p1
            = basix::element::create_lagrange(basix::cell::type::triangle, 1, basix::element::lagrange_variant::equispaced, false);
        const int ndg = dg.dim(), nel = nrt + ndg;
        const std::int32_t nc = mesh->ncell;
        const std::int32_t nflux = mesh->nfct * k + nc * (nrt - 3 * k), nmixed = nflux + nc * ndg;
        std::vector<std::int32_t> mixed((size_t)nc * nel), flux((size_t)nc * nrt);
        for (std::int32_t c = 0; c < nc; ++c)
          for (int l = 0; l < nel; ++l)
          {
            std::int32_t g;
            if (l < 3 * k)
              g = mesh->cell_fct[3 * c + l / k] * k + l % k;
            else if (l < nrt)
              g = mesh->nfct * k + c * (nrt - 3 * k) + (l - 3 * k);
            else
              g = nflux + c * ndg + (l - nrt);
            mixed[(size_t)c * nel + l] = g;
            if (l < nrt)
              flux[(size_t)c * nrt + l] = g;
          }
        auto trafo = std::make_shared<RtTransform>(k, nrt, mesh->fct_perms);
        auto install = [&](fem::FiniteElement& e, int nd)
        {
          e.set_transformations(
              [trafo](const std::span<double>& d, const std::span<const std::uint32_t>&, std::int32_t c, int bs)
              { trafo->pre(d, c, bs, false); },
              [trafo, nd](const std::span<double>& d, const std::span<const std::uint32_t>&, std::int32_t c, int bs)
              { trafo->post_t(d, c, bs, nd); },
              [trafo](const std::span<double>& d, const std::span<const std::uint32_t>&, std::int32_t c, int bs)
              { trafo->pre(d, c, bs, true); });
        };
        auto el_mixed = std::make_shared<fem::FiniteElement>(rt, 1);
        el_mixed->set_space_dimension(nel);
        install(*el_mixed, nel);
        auto el_flux = std::make_shared<fem::FiniteElement>(rt, 1);
        install(*el_flux, nrt);
        fem::ElementDofLayout layout(rt.entity_dofs());
        auto V = std::make_shared<fem::FunctionSpace>(
            msh, el_mixed, std::make_shared<const fem::DofMap>(AL::regular(mixed.data(), nc, nel), nmixed, 1, layout));
        // V.sub(0): the flux sub-space keeps the parent's numbering and index map (`FluxEqlbEV.py:158`)
        auto V0 = std::make_shared<fem::FunctionSpace>(
            msh, el_flux, std::make_shared<const fem::DofMap>(AL::regular(flux.data(), nc, nrt), nmixed, 1, layout));
        V->add_sub(V0);
        auto V_flux = std::make_shared<fem::FunctionSpace>(
            msh, el_flux, std::make_shared<const fem::DofMap>(AL::regular(flux.data(), nc, nrt), nflux, 1, layout));
        auto V_dg2 = std::make_shared<fem::FunctionSpace>(
            msh, std::make_shared<const fem::FiniteElement>(dg, 2),
            std::make_shared<const fem::DofMap>(AL::regular(mesh->dg_dofmap, nc, ndg), nc * ndg, 2, fem::ElementDofLayout(dg.entity_dofs())));
        auto V_dg = std::make_shared<fem::FunctionSpace>(
            msh, std::make_shared<const fem::FiniteElement>(dg, 1),
            std::make_shared<const fem::DofMap>(AL::regular(mesh->dg_dofmap, nc, ndg), nc * ndg, 1, fem::ElementDofLayout(dg.entity_dofs())));
        auto V_hat = std::make_shared<fem::FunctionSpace>(
            msh, std::make_shared<const fem::FiniteElement>(p1, 1),
            std::make_shared<const fem::DofMap>(AL::regular(mesh->cell_node, nc, 3), mesh->nnode, 1, fem::ElementDofLayout(p1.entity_dofs())));

        // boundary data: facet types + boundary function in the global orientation (c_glob = R c_loc)
        std::vector<std::vector<double>> bvals(nrhs, std::vector<double>(nmixed, 0.0));
        std::vector<std::shared_ptr<fem::Function<double>>> bfuncs;
        std::vector<std::vector<std::shared_ptr<eqlb::base::FluxBC<double>>>> no_bcs(nrhs);
        std::vector<std::vector<std::int32_t>> fct_prime(nrhs);
        for (int r = 0; r < nrhs; ++r)
        {
          for (int f = 0; f < mesh->nfct; ++f)
          {
            const int8_t ft = facet_type[(size_t)r * mesh->nfct + f];
            if (ft == eqlb::base::PatchFacetType::essnt_primal)
              fct_prime[r].push_back(f);
            if (ft == eqlb::base::PatchFacetType::essnt_dual && bflux && bflux[r])
            {
              const std::int32_t c = mesh->fct_cell[mesh->fct_cell_off[f]];
              int fl = 0;
              for (int j = 0; j < 3; ++j)
                if (mesh->cell_fct[3 * c + j] == f)
                  fl = j;
              std::vector<double> cl(nrt, 0.0);
              for (int j = 0; j < k; ++j)
                cl[fl * k + j] = bflux[r][(size_t)c * nrt + fl * k + j];
              trafo->pre(std::span<double>(cl), c, 1, true);
              for (int j = 0; j < k; ++j)
                bvals[r][f * k + j] = cl[fl * k + j];
            }
          }
          bfuncs.push_back(std::make_shared<fem::Function<double>>(V, std::make_shared<la::Vector<double>>(bvals[r].data(), nmixed)));
        }
        std::shared_ptr<eqlb::base::BoundaryData<double>> bd = std::make_shared<InjectedBoundaryData>(
            no_bcs, bfuncs, V0, std::max(2 * (k - 1), 0), fct_prime, false, mesh, nrhs, facet_type, nullptr, nullptr);

        // forms
        auto forms = std::make_shared<EvForms>(rt, dg, k);
        using kern_t = fem::Form<double>::kernel_t;
        kern_t ka = [forms](double* A, const double*, const double*, const double* x, const int*, const std::uint8_t*) { forms->kernel_a(A, x); };
        kern_t kp = [forms](double* P, const double*, const double*, const double* x, const int*, const std::uint8_t*) { forms->kernel_lpen(P, x); };
        kern_t kl = [forms](double* L, const double* w, const double*, const double* x, const int*, const std::uint8_t*) { forms->kernel_l(L, w, x); };
        fem::Form<double> a({V, V}, ka, {}, {}, msh);
        fem::Form<double> l_pen({V_dg}, kp, {}, {}, msh);
        std::vector<double> hat0(mesh->nnode, 0.0);
        auto hat = std::make_shared<fem::Function<double>>(V_hat, std::make_shared<la::Vector<double>>(hat0.data(), hat0.size()));
        hat->name = "hat";
        std::vector<std::shared_ptr<const fem::Form<double>>> l;
        std::vector<std::shared_ptr<fem::Function<double>>> flux_hdiv;
        for (int r = 0; r < nrhs; ++r)
        {
          auto Gf = std::make_shared<fem::Function<double>>(V_dg2, std::make_shared<la::Vector<double>>(const_cast<double*>(G[r]), (size_t)nc * ndg * 2));
          auto Ff = std::make_shared<fem::Function<double>>(V_dg, std::make_shared<la::Vector<double>>(const_cast<double*>(F[r]), (size_t)nc * ndg));
          std::vector<std::shared_ptr<const fem::Function<double>>> cf{Gf, Ff, hat};
          l.push_back(std::make_shared<const fem::Form<double>>(std::vector<std::shared_ptr<const fem::FunctionSpace>>{V}, kl, cf,
                                                                std::vector<std::shared_ptr<const fem::Constant<double>>>{}, msh));
          flux_hdiv.push_back(std::make_shared<fem::Function<double>>(V_flux, std::make_shared<la::Vector<double>>(sigma[r], nflux)));
        }
        eqlb::ev::reconstruction<double>(a, l_pen, l, flux_hdiv, bd);
      });
}
}

extern "C"
{
/// The reference's `base::BoundaryData` CONSTRUCTOR (`base/BoundaryData.cpp:279-633`) with FluxBC objects
/// (`base/FluxBC.hpp`) whose boundary kernel evaluates a polynomial normal traction: on facet fcts[i] of
/// rhs r the prescribed outward normal flux is g(s) = sum_j coeffs[i*ncoef+j] s^j, s = facet parameter of the
/// adjacent cell.  Interpolation branch (`:580-597`) when qdegree_proj < 0, projection branch (`:511-578`)
/// otherwise (the kernel is then evaluated at the facet quadrature points).
///   nbc[r] facets per rhs, fcts[r], coeffs[r]; prime[r] (nprime[r]) Dirichlet facets of the primal problem.
/// Outputs (any may be null): facet_type [nrhs][nfct], bvals[r] [ncell*nrt] (DRT boundary function),
/// node_on_bnd [nnode] (stress).
int ref_boundary_data(const eqlb_mesh* mesh, const ref_element* elmt, int nrhs, int stress, int qdegree_proj, const int32_t* nprime,
                      const int32_t* const* prime, const int32_t* nbc, const int32_t* const* fcts, int ncoef,
                      const double* const* coeffs, int8_t* facet_type, double* const* bvals, int8_t* node_on_bnd)
{
  return guarded(
      [&]
      {
        Spaces sp(mesh, elmt);
        const int k = elmt->k;
        const size_t n = (size_t)mesh->ncell * sp.nrt;
        // coefficient function of the traction: per cell [local facet][ncoef] (DG-like space, identity dofmap)
        const int nloc = 3 * ncoef;
        std::vector<std::int32_t> ident((size_t)mesh->ncell * nloc);
        for (size_t i = 0; i < ident.size(); ++i)
          ident[i] = (std::int32_t)i;
        basix::FiniteElement dg0 = basix::element::create_lagrange(basix::cell::type::triangle, 0,
                                                                    basix::element::lagrange_variant::equispaced, true);
        auto el_tr = std::make_shared<fem::FiniteElement>(dg0, 1);
        el_tr->set_space_dimension(nloc);
        auto V_tr = std::make_shared<fem::FunctionSpace>(
            sp.mesh, el_tr,
            std::make_shared<const fem::DofMap>(AL::regular(ident.data(), mesh->ncell, nloc), mesh->ncell * nloc, 1,
                                                fem::ElementDofLayout(dg0.entity_dofs())));
        // evaluation points per facet: interpolation points of the flux element or the facet quadrature points
        const bool project = qdegree_proj >= 0;
        const int qdeg = project ? qdegree_proj : 2 * (k - 1);
        std::vector<double> s_pts;  // facet parameter of the evaluation points
        if (project)
        {
          auto q = basix::quadrature::make_quadrature(basix::cell::type::interval, qdeg);
          s_pts = q[0];
        }
        else
        {
          // facet 1 interpolation points are (0, s): read s from the element's point list
          int nip = 0;
          while (elmt->X[2 * nip] > 0.0)
            ++nip;
          for (int i = 0; i < nip; ++i)
            s_pts.push_back(elmt->X[2 * (nip + i) + 1]);
        }
        const int nev = (int)s_pts.size();
        std::function<void(double*, const double*, const double*, const double*, const int*, const std::uint8_t*)> kern
            = [s_pts, nev, ncoef](double* v, const double* w, const double*, const double*, const int*, const std::uint8_t*)
        {
          for (int f = 0; f < 3; ++f)
            for (int l = 0; l < nev; ++l)
            {
              double g = 0.0, sp_ = 1.0;
              for (int j = 0; j < ncoef; ++j)
              {
                g += w[f * ncoef + j] * sp_;
                sp_ *= s_pts[l];
              }
              v[f * nev + l] = g;
            }
        };
        std::vector<std::vector<double>> tr(nrhs), bv(nrhs);
        std::vector<std::vector<std::shared_ptr<eqlb::base::FluxBC<double>>>> bcs(nrhs);
        std::vector<std::shared_ptr<fem::Function<double>>> bfuncs;
        std::vector<std::vector<std::int32_t>> fct_prime(nrhs);
        for (int r = 0; r < nrhs; ++r)
        {
          tr[r].assign((size_t)mesh->ncell * nloc, 0.0);
          for (int i = 0; i < nbc[r]; ++i)
          {
            const std::int32_t f = fcts[r][i];
            const std::int32_t c = mesh->fct_cell[mesh->fct_cell_off[f]];
            int fl = 0;
            for (int j = 0; j < 3; ++j)
              if (mesh->cell_fct[3 * c + j] == f)
                fl = j;
            for (int j = 0; j < ncoef; ++j)
              tr[r][(size_t)c * nloc + fl * ncoef + j] = coeffs[r][(size_t)i * ncoef + j];
          }
          auto trf = std::make_shared<const fem::Function<double>>(V_tr, std::make_shared<la::Vector<double>>(tr[r].data(), tr[r].size()));
          if (nbc[r] > 0)
          {
            std::vector<std::int32_t> fl(fcts[r], fcts[r] + nbc[r]);
            std::vector<std::shared_ptr<const fem::Function<double>>> cf{trf};
            if (project)
              bcs[r].push_back(std::make_shared<eqlb::base::FluxBC<double>>(sp.V_hdiv, fl, kern, nev, qdeg, cf, std::vector<int>{0},
                                                                             std::vector<std::shared_ptr<const fem::Constant<double>>>{}));
            else
              bcs[r].push_back(std::make_shared<eqlb::base::FluxBC<double>>(sp.V_hdiv, fl, kern, nev, cf, std::vector<int>{0},
                                                                             std::vector<std::shared_ptr<const fem::Constant<double>>>{}));
          }
          bv[r].assign(n, 0.0);
          bfuncs.push_back(std::make_shared<fem::Function<double>>(sp.V_hdiv, std::make_shared<la::Vector<double>>(bv[r].data(), n)));
          fct_prime[r].assign(prime[r], prime[r] + nprime[r]);
        }
        eqlb::base::BoundaryData<double> bd(bcs, bfuncs, sp.V_hdiv, true, qdeg, fct_prime, stress != 0);
        if (facet_type)
        {
          auto ft = bd.facet_type();
          for (int r = 0; r < nrhs; ++r)
            for (int f = 0; f < mesh->nfct; ++f)
              facet_type[(size_t)r * mesh->nfct + f] = ft(r, f);
        }
        if (bvals)
          for (int r = 0; r < nrhs; ++r)
            if (bvals[r])
              std::memcpy(bvals[r], bv[r].data(), n * sizeof(double));
        if (node_on_bnd && stress)
        {
          auto nb = bd.node_on_essnt_boundary_stress();
          for (int i = 0; i < mesh->nnode; ++i)
            node_on_bnd[i] = nb[i];
        }
      });
}
}

extern "C"
{
/// The reference's cell-wise solver `base::local_solver_{lu,cholesky,cg}` (`base/local_solver.hpp:38-240`, the
/// engine of `lsolver/projection.py:17-77`) on the fixed projection forms: a = (u, v) on DG_p, l_i = (data_i, v)
/// with data_i given at the cell quadrature points, qvals[i] [ncell][nq].  The element loop, the solver and the
/// scatter are the reference's; the two FFCx cell kernels are restated with the tables of the C ABI (`qwts`, `dg_q`).
/// solver: 0 = PartialPivLU, 1 = LLT, 2 = ConjugateGradient.  out[i] [ncell*ndg] (DOLFINx DG layout).
int ref_local_solver(const eqlb_mesh* mesh, const eqlb_tables* T, int nfun, int solver, const double* const* qvals, double* const* out)
{
  return guarded(
      [&]
      {
        auto msh = make_mesh(mesh);
        const int nc = mesh->ncell, ndg = T->ndg, nq = T->nq;
        basix::FiniteElement dg
            = basix::element::create_lagrange(basix::cell::type::triangle, T->p, basix::element::lagrange_variant::equispaced, true);
        if (dg.dim() != ndg)
          throw std::runtime_error("ref_local_solver: DG dimension mismatch");
        std::vector<std::int32_t> ident_dg((size_t)nc * ndg), ident_q((size_t)nc * nq);
        for (size_t i = 0; i < ident_dg.size(); ++i)
          ident_dg[i] = (std::int32_t)i;
        for (size_t i = 0; i < ident_q.size(); ++i)
          ident_q[i] = (std::int32_t)i;
        auto V_dg = std::make_shared<fem::FunctionSpace>(
            msh, std::make_shared<const fem::FiniteElement>(dg, 1),
            std::make_shared<const fem::DofMap>(AL::regular(ident_dg.data(), nc, ndg), nc * ndg, 1, fem::ElementDofLayout(dg.entity_dofs())));
        // "quadrature element": one value per quadrature point and cell
        auto el_q = std::make_shared<fem::FiniteElement>(dg, 1);
        el_q->set_space_dimension(nq);
        auto V_q = std::make_shared<fem::FunctionSpace>(
            msh, el_q, std::make_shared<const fem::DofMap>(AL::regular(ident_q.data(), nc, nq), nc * nq, 1, fem::ElementDofLayout(dg.entity_dofs())));
        auto detJ = [](const double* x) { return (x[3] - x[0]) * (x[7] - x[1]) - (x[6] - x[0]) * (x[4] - x[1]); };
        using kern_t = fem::Form<double>::kernel_t;
        kern_t ka = [T, ndg, nq, detJ](double* A, const double*, const double*, const double* x, const int*, const std::uint8_t*)
        {
          const double d = std::fabs(detJ(x));
          for (int q = 0; q < nq; ++q)
            for (int i = 0; i < ndg; ++i)
              for (int j = 0; j < ndg; ++j)
                A[i * ndg + j] += d * T->qwts[q] * T->dg_q[(size_t)q * ndg + i] * T->dg_q[(size_t)q * ndg + j];
        };
        kern_t kl = [T, ndg, nq, detJ](double* L, const double* w, const double*, const double* x, const int*, const std::uint8_t*)
        {
          const double d = std::fabs(detJ(x));
          for (int q = 0; q < nq; ++q)
            for (int i = 0; i < ndg; ++i)
              L[i] += d * T->qwts[q] * w[q] * T->dg_q[(size_t)q * ndg + i];
        };
        fem::Form<double> a({V_dg, V_dg}, ka, {}, {}, msh);
        std::vector<std::shared_ptr<const fem::Form<double>>> l;
        std::vector<std::shared_ptr<fem::Function<double>>> sol;
        for (int i = 0; i < nfun; ++i)
        {
          auto data = std::make_shared<fem::Function<double>>(
              V_q, std::make_shared<la::Vector<double>>(const_cast<double*>(qvals[i]), (size_t)nc * nq));
          std::vector<std::shared_ptr<const fem::Function<double>>> cf{data};
          l.push_back(std::make_shared<const fem::Form<double>>(std::vector<std::shared_ptr<const fem::FunctionSpace>>{V_dg}, kl, cf,
                                                                std::vector<std::shared_ptr<const fem::Constant<double>>>{}, msh));
          sol.push_back(std::make_shared<fem::Function<double>>(V_dg, std::make_shared<la::Vector<double>>(out[i], (size_t)nc * ndg)));
        }
        if (solver == 0)
          eqlb::base::local_solver_lu<double>(sol, a, l);
        else if (solver == 1)
          eqlb::base::local_solver_cholesky<double>(sol, a, l);
        else
          eqlb::base::local_solver_cg<double>(sol, a, l);
      });
}
}
