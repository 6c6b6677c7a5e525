// pybind11 module `dolfinx_eqlb_b200.cpp`: the compiled host layer of the drop-in boundary.
//
// Successor of the reference's `python/dolfinx_eqlb/wrappers.cpp` (module `dolfinx_eqlb.cpp`): the same
// Python-visible names and argument order,
//     reconstruct_fluxes_minimisation(a, l_pen, l, flux_hdiv, boundary_data)                 wrappers.cpp:85-95
//     reconstruct_fluxes_semiexplt(flux_hdiv, flux_dg, rhs_dg, boundary_data, reconstruct_stress)      :97-115
//     reconstruct_fluxes_semiexplt_with_kornconst(..., cells_kornconst)                               :117-137
//     local_solver_lu / local_solver_cholesky / local_solver_cg(solution, a, l)                         :54-79
//     FluxBC(function_space, facets, pointer_boundary_kernel, nevals_per_fct[, quadrature_degree],
//            coefficients, position_of_coefficients, constants)                                      :144-232
//     BoundaryData(list_of_bcs, list_of_boundary_fluxes, V_flux_hdiv, rtflux_is_custom,
//                  quadrature_degree, list_bfcts_prime, reconstruct_stress)                           :235-256
// host code in C++ that calls the CUDA library through the C ABI (`include/eqlb_b200.h`) - nothing else.
// DOLFINx is not available in this image, so the DOLFINx classes the reference's signatures take are
// replaced by thin array holders (`Mesh`, `FunctionSpace`, `Function`, `Form`) carrying exactly what the
// reference extracts from them; in a DOLFINx deployment the same functions take the DOLFINx
// objects and fill these holders.
//
// A handle cache keyed by (mesh, degrees, number of RHS, flags) keeps the device-resident problem
// (`eqlb_create`: mesh upload, Jacobians, colouring) alive across calls; the boundary data of a
// `BoundaryData` object are built on the device once (`eqlb_set_bcs_poly`) and rebuilt only when the
// object changes.  Errors surface as RuntimeError with the reference's texts.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/eqlb_b200.h"

namespace py = pybind11;

namespace
{
using darray = py::array_t<double, py::array::c_style | py::array::forcecast>;
using iarray = py::array_t<std::int32_t, py::array::c_style | py::array::forcecast>;

void check(int rc)
{
  if (rc != EQLB_OK)
    throw std::runtime_error(eqlb_last_error());
}

struct HandleEntry
{
  eqlb_handle* h = nullptr;
  const void* bd = nullptr;  // BoundaryData the handle's BCs were built from
  std::uint64_t bd_version = 0;
  ~HandleEntry()
  {
    if (h)
      eqlb_destroy(h);
  }
};

/// what the reference reads from dolfinx::mesh::Mesh (`se/Patch.cpp:23-28`, `se/reconstruction.hpp:83-93`)
struct Mesh
{
  darray x;
  iarray cell_node, cell_fct, fct_node, fct_cell_off, fct_cell, node_cell_off, node_cell, node_fct_off, node_fct;
  py::array_t<std::uint8_t, py::array::c_style | py::array::forcecast> fct_perms;
  py::array_t<std::uint32_t, py::array::c_style | py::array::forcecast> cell_perm_info;
  int nnode, ncell, nfct;
  std::map<std::tuple<int, int, int, unsigned>, std::shared_ptr<HandleEntry>> cache;

  Mesh(darray x_, iarray cn, iarray cf, iarray fn, iarray fco, iarray fc, iarray nco, iarray nc, iarray nfo, iarray nf,
       py::array_t<std::uint8_t, py::array::c_style | py::array::forcecast> perms,
       py::array_t<std::uint32_t, py::array::c_style | py::array::forcecast> info)
      : x(x_), cell_node(cn), cell_fct(cf), fct_node(fn), fct_cell_off(fco), fct_cell(fc), node_cell_off(nco), node_cell(nc),
        node_fct_off(nfo), node_fct(nf), fct_perms(perms), cell_perm_info(info)
  {
    nnode = (int)x.shape(0);
    ncell = (int)cell_node.shape(0);
    nfct = (int)fct_node.shape(0);
    if (x.ndim() != 2 || x.shape(1) != 3 || cell_node.ndim() != 2 || cell_node.shape(1) != 3)
      throw std::runtime_error("Mesh: x must be [nnode][3], cell_node [ncell][3]");
  }
  int num_cached_handles() const { return (int)cache.size(); }
  void clear_cache() { cache.clear(); }
};

/// family: "DRT" (discontinuous hierarchic RT, FluxEqlbSE), "RT" (conforming, FluxEqlbEV), "DG", "P"
struct FunctionSpace
{
  std::shared_ptr<Mesh> mesh;
  std::string family;
  int degree, bs;
  FunctionSpace(std::shared_ptr<Mesh> m, std::string fam, int deg, int bs_) : mesh(std::move(m)), family(std::move(fam)), degree(deg), bs(bs_)
  {
    if (family != "DRT" && family != "RT" && family != "DG" && family != "P")
      throw std::runtime_error("FunctionSpace: family must be DRT, RT, DG or P");
  }
  int ndofs_cell() const
  {
    if (family == "DG" || family == "P")
      return (degree + 1) * (degree + 2) / 2;
    return degree * (degree + 2);
  }
  std::int64_t size() const
  {
    const int k = degree;
    if (family == "DRT")
      return (std::int64_t)mesh->ncell * k * (k + 2);
    if (family == "RT")
      return (std::int64_t)mesh->nfct * k + (std::int64_t)mesh->ncell * (k * k - k);
    if (family == "DG")
      return (std::int64_t)mesh->ncell * ndofs_cell() * bs;
    return (std::int64_t)mesh->nnode * bs;  // P1 only
  }
};

struct Function
{
  std::shared_ptr<FunctionSpace> V;
  darray x;
  std::string name = "u";
  int host_calls = 0;   // equilibration calls that moved this vector over PCIe
  bool pinned = false;  // page-locked by note_host_call
  // A first call runs from pageable memory (the library stages the copies through its pinned pool); when a vector
  // shows up in a second call (time loops) it is page-locked once so that the stage pipeline can take over.
  void note_host_call()
  {
    if (pinned || x.nbytes() < (py::ssize_t)(4 << 20))
      return;
    if (++host_calls >= 2 && eqlb_pin_host(x.mutable_data(), (size_t)x.nbytes()) == EQLB_OK)
      pinned = true;
  }
  ~Function()
  {
    if (pinned)
      eqlb_unpin_host(x.mutable_data());
  }
  Function(const Function&) = delete;
  Function& operator=(const Function&) = delete;
  Function(std::shared_ptr<FunctionSpace> V_, py::object arr) : V(std::move(V_))
  {
    if (arr.is_none())
      x = darray((py::ssize_t)V->size());
    else
    {
      // the caller's array is used IN PLACE (the reference accumulates into `function.x.array`)
      if (!py::isinstance<py::array>(arr))
        throw std::runtime_error("Function: array expected");
      py::array a = py::reinterpret_borrow<py::array>(arr);
      if (a.dtype().kind() != 'f' || a.itemsize() != 8 || !(a.flags() & py::array::c_style))
        throw std::runtime_error("Function: contiguous float64 array required (it is updated in place)");
      x = py::reinterpret_borrow<darray>(arr);
    }
    if ((std::int64_t)x.size() != V->size())
      throw std::runtime_error("Function: array size does not match the function space");
    if (arr.is_none())
      std::memset(x.mutable_data(), 0, sizeof(double) * x.size());
  }
};

using bc_kernel_t = void (*)(double*, const double*, const double*, const double*, const int*, const std::uint8_t*);

/// `base::FluxBC` (`base/FluxBC.hpp`): either a compiled boundary kernel (the reference's signature; evaluated
/// on the host like the reference does) or polynomial coefficients per facet (evaluated on the device)
struct FluxBC
{
  std::shared_ptr<FunctionSpace> V;
  iarray facets;
  bc_kernel_t kernel = nullptr;
  int nevals = 0, quadrature_degree = 0;
  bool projection = false;
  std::vector<std::shared_ptr<Function>> coefficients;
  std::vector<int> positions;
  std::vector<darray> constants;
  darray poly;  // [nfct][ncoef] (polynomial form)
  bool is_poly = false;
};

struct BoundaryData
{
  std::vector<std::vector<std::shared_ptr<FluxBC>>> bcs;
  std::vector<std::shared_ptr<Function>> bfuncs;
  std::shared_ptr<FunctionSpace> V;
  bool custom, stress;
  int qdegree;
  std::vector<iarray> prime;
  std::uint64_t version = 1;
  int num_rhs() const { return (int)bcs.size(); }
  bool all_poly() const
  {
    for (auto& l : bcs)
      for (auto& b : l)
        if (!b->is_poly)
          return false;
    return true;
  }
};

/// fixed forms of the hot path (the reference takes dolfinx::fem::Form objects built from these UFL forms)
struct Form
{
  std::string kind;  // "ev_a", "ev_lpen", "ev_l", "mass", "projection_rhs"
  std::shared_ptr<FunctionSpace> V;
  std::shared_ptr<Function> flux_dg, rhs_dg;
  darray qvals;
};

// ---- tables: built by dolfinx_eqlb_b200.tables.make_tables (exact rational construction), packed here ----
struct Tables
{
  py::object obj;
  std::vector<py::array> keep;
  eqlb_tables t{};
  Tables(int k, int p)
  {
    py::object mod = py::module_::import("dolfinx_eqlb_b200.tables");
    obj = mod.attr("make_tables")(k, p);
    auto geti = [&](const char* n) { return obj.attr(n).cast<int>(); };
    t.k = geti("k");
    t.p = geti("p");
    t.nrt = geti("nrt");
    t.ndg = geti("ndg");
    t.ndg_fct = geti("ndg_fct");
    t.nq = geti("nq");
    t.nqf = geti("nqf");
    t.ndiv = geti("ndiv");
    t.nadd = geti("nadd");
    t.npk = geti("npk");
    auto getd = [&](const char* n) -> const double*
    {
      darray a = darray::ensure(obj.attr(n));
      keep.push_back(a);
      return a.data();
    };
    auto getiarr = [&](const char* n) -> const std::int32_t*
    {
      iarray a = iarray::ensure(obj.attr(n));
      if (a.size() == 0)
        a = iarray(2);
      keep.push_back(a);
      return a.data();
    };
    t.qpts = getd("qpts");
    t.qwts = getd("qwts");
    t.fpts_s = getd("fpts_s");
    t.fwts = getd("fwts");
    t.M = getd("M");
    t.rt_q = getd("rt_q");
    t.rt_f = getd("rt_f");
    t.dg_q = getd("dg_q");
    t.dg_f = getd("dg_f");
    t.hat_q = getd("hat_q");
    t.hat_f = getd("hat_f");
    t.trafo = getd("trafo");
    t.fct_closure = getiarr("fct_closure");
    t.div_lm = getiarr("div_lm");
    t.rt_mass = getd("rt_mass");
    t.fct_mom = getd("fct_mom");
    t.cell_mom_f = getd("cell_mom_f");
    t.cell_mom_g = getd("cell_mom_g");
    t.bc_mat = getd("bc_mat");
    t.rt_p1 = getd("rt_p1");
    t.dg_mono = getd("dg_mono");
    t.hat_dg_rt = getd("hat_dg_rt");
    t.mono_int = getd("mono_int");
    t.rt_basix_fct = getd("rt_basix_fct");
    t.rt_basix_int = getd("rt_basix_int");
    t.pk_grad_dg = getd("pk_grad_dg");
    t.pk_to_dg = getd("pk_to_dg");
    t.pk_q = getd("pk_q");
    t.pk_gq = getd("pk_gq");
    t.pk_hq = getd("pk_hq");
    t.rt_div_q = getd("rt_div_q");
  }
};

std::shared_ptr<HandleEntry> get_handle(Mesh& m, int k, int p, int nrhs, unsigned flags)
{
  auto key = std::make_tuple(k, p, nrhs, flags);
  auto it = m.cache.find(key);
  if (it != m.cache.end())
    return it->second;
  Tables T(k, p);
  eqlb_mesh em{};
  em.nnode = m.nnode;
  em.ncell = m.ncell;
  em.nfct = m.nfct;
  em.x = m.x.data();
  em.cell_node = m.cell_node.data();
  em.cell_fct = m.cell_fct.data();
  em.fct_node = m.fct_node.data();
  em.fct_cell_off = m.fct_cell_off.data();
  em.fct_cell = m.fct_cell.data();
  em.node_cell_off = m.node_cell_off.data();
  em.node_cell = m.node_cell.data();
  em.node_fct_off = m.node_fct_off.data();
  em.node_fct = m.node_fct.data();
  em.fct_perms = m.fct_perms.data();
  em.cell_perm_info = m.cell_perm_info.data();
  em.dg_dofmap = nullptr;  // the holders' DG spaces use the DOLFINx layout cell * ndg + local (a DOLFINx build passes the dofmap)
  em.node_owned = nullptr;
  auto e = std::make_shared<HandleEntry>();
  check(eqlb_create(&em, &T.t, nrhs, flags, &e->h));
  m.cache[key] = e;
  return e;
}

/// Boundary data of a handle: device path for polynomial FluxBCs, else host evaluation of the compiled
/// boundary kernels (interpolation branch of `base/BoundaryData.cpp:580-597`) + eqlb_set_bcs
void ensure_bcs(HandleEntry& e, BoundaryData& bd, Mesh& m, int k)
{
  if (e.bd == &bd && e.bd_version == bd.version)
    return;
  const int nrhs = bd.num_rhs();
  std::vector<std::int32_t> nprime(nrhs), nbc(nrhs);
  std::vector<const std::int32_t*> prime(nrhs);
  for (int r = 0; r < nrhs; ++r)
  {
    nprime[r] = (std::int32_t)bd.prime[r].size();
    prime[r] = bd.prime[r].data();
    nbc[r] = (std::int32_t)bd.bcs[r].size();
  }
  // compiled kernels are evaluated on the host into polynomial-free facet moments: turn every FluxBC into
  // per-facet monomial coefficients of its interpolant (degree k-1), then use the device path for all
  std::vector<std::vector<eqlb_fluxbc>> rows(nrhs);
  std::vector<std::vector<double>> owned;
  for (int r = 0; r < nrhs; ++r)
    for (auto& b : bd.bcs[r])
    {
      eqlb_fluxbc fb{};
      fb.nfct = (std::int32_t)b->facets.size();
      fb.facets = b->facets.data();
      if (b->is_poly)
      {
        fb.ncoef = (std::int32_t)b->poly.shape(1);
        fb.coeffs = b->poly.data();
      }
      else
      {
        // interpolation points = the k+1 (k == 1: 1) Gauss points of the hierarchic RT element on each facet
        // (`e_raviart_thomas.py:63-71`); the degree-(nev-1) Lagrange interpolant of the kernel values through
        // them reproduces exactly the moments int g s^j the reference's interpolation operator computes
        if (b->projection)
          throw std::runtime_error("FluxBC: facet-local projection of compiled kernels is not supported on the device path; "
                                   "give the traction as a polynomial (DG_(k-1)) FluxBC");
        const int nev = b->nevals;
        const int nip = (k == 1) ? 1 : k + 1;
        if (nev != nip)
          throw std::runtime_error("BoundaryData: FunctionSpace of FluxBC does not match!");
        // Gauss-Legendre nodes on [0, 1]
        std::vector<double> s(nev);
        {
          // Newton on Legendre P_n
          for (int i = 0; i < nev; ++i)
          {
            double t = std::cos(M_PI * (i + 0.75) / (nev + 0.5));
            for (int it = 0; it < 100; ++it)
            {
              double p0 = 1.0, p1 = t;
              for (int n = 2; n <= nev; ++n)
              {
                const double p2 = ((2.0 * n - 1.0) * t * p1 - (n - 1.0) * p0) / n;
                p0 = p1;
                p1 = p2;
              }
              const double pn = (nev == 0) ? 1.0 : p1, pn1 = (nev == 1) ? 1.0 : p0;
              const double dp = nev * (t * pn - pn1) / (t * t - 1.0);
              const double dt = pn / dp;
              t -= dt;
              if (std::fabs(dt) < 1e-16)
                break;
            }
            s[nev - 1 - i] = 0.5 + 0.5 * t;
          }
        }
        // pack coefficients like FluxBC::extract_coefficients (`base/FluxBC.hpp:146-222`)
        int cstride = 0;
        for (auto& f : b->coefficients)
          cstride += f->V->ndofs_cell() * f->V->bs;
        std::vector<double> cst;
        for (auto& c : b->constants)
          cst.insert(cst.end(), c.data(), c.data() + c.size());
        owned.emplace_back((size_t)fb.nfct * nev, 0.0);
        std::vector<double>& poly = owned.back();
        std::vector<double> w(std::max(cstride, 1)), vals(3 * nev), coords(9);
        // monomial coefficients of the interpolant: solve the nev x nev Vandermonde system per facet
        for (int i = 0; i < fb.nfct; ++i)
        {
          const std::int32_t f = fb.facets[i];
          const std::int32_t c = m.fct_cell.data()[m.fct_cell_off.data()[f]];
          int lf = 0;
          for (int j = 0; j < 3; ++j)
            if (m.cell_fct.data()[3 * c + j] == f)
              lf = j;
          int off = 0;
          for (int pos : b->positions)
          {
            auto& fn = b->coefficients.at(pos);
            const int nd = fn->V->ndofs_cell(), bs = fn->V->bs;
            for (int j = 0; j < nd; ++j)
              for (int q = 0; q < bs; ++q)
                w[off + j * bs + q] = fn->x.data()[((size_t)c * nd + j) * bs + q];  // DG layout
            off += nd * bs;
          }
          for (int j = 0; j < 3; ++j)
            for (int d = 0; d < 3; ++d)
              coords[3 * j + d] = m.x.data()[3 * (size_t)m.cell_node.data()[3 * c + j] + d];
          std::fill(vals.begin(), vals.end(), 0.0);
          b->kernel(vals.data(), w.data(), cst.data(), coords.data(), nullptr, nullptr);
          // Vandermonde solve (tiny): sum_j a_j s_l^j = vals[lf*nev + l]
          std::vector<double> A((size_t)nev * nev), rhs(nev);
          for (int l = 0; l < nev; ++l)
          {
            double pw = 1.0;
            for (int j = 0; j < nev; ++j)
            {
              A[(size_t)l * nev + j] = pw;
              pw *= s[l];
            }
            rhs[l] = vals[lf * nev + l];
          }
          for (int cidx = 0; cidx < nev; ++cidx)
          {
            int piv = cidx;
            for (int rr = cidx + 1; rr < nev; ++rr)
              if (std::fabs(A[(size_t)rr * nev + cidx]) > std::fabs(A[(size_t)piv * nev + cidx]))
                piv = rr;
            for (int j = 0; j < nev; ++j)
              std::swap(A[(size_t)cidx * nev + j], A[(size_t)piv * nev + j]);
            std::swap(rhs[cidx], rhs[piv]);
            for (int rr = cidx + 1; rr < nev; ++rr)
            {
              const double fct = A[(size_t)rr * nev + cidx] / A[(size_t)cidx * nev + cidx];
              for (int j = cidx; j < nev; ++j)
                A[(size_t)rr * nev + j] -= fct * A[(size_t)cidx * nev + j];
              rhs[rr] -= fct * rhs[cidx];
            }
          }
          for (int rr = nev - 1; rr >= 0; --rr)
          {
            double sacc = rhs[rr];
            for (int j = rr + 1; j < nev; ++j)
              sacc -= A[(size_t)rr * nev + j] * poly[(size_t)i * nev + j];
            poly[(size_t)i * nev + rr] = sacc / A[(size_t)rr * nev + rr];
          }
        }
        fb.ncoef = nev;
        fb.coeffs = poly.data();
      }
      rows[r].push_back(fb);
    }
  std::vector<const eqlb_fluxbc*> rp(nrhs);
  for (int r = 0; r < nrhs; ++r)
    rp[r] = rows[r].empty() ? nullptr : rows[r].data();
  check(eqlb_set_bcs_poly(e.h, nprime.data(), prime.data(), nbc.data(), rp.data()));
  // the reference stores the boundary DOFs in the functions handed to BoundaryData
  // (DRT layout cell * nrt + local; the boundary functions of the EV path live in the mixed space and are not filled)
  std::vector<double*> bp(nrhs);
  const std::int64_t ndrt = (std::int64_t)m.ncell * k * (k + 2);
  for (int r = 0; r < nrhs; ++r)
    bp[r] = (bd.bfuncs[r]->x.size() == ndrt) ? bd.bfuncs[r]->x.mutable_data() : nullptr;
  check(eqlb_get_boundary_data(e.h, nullptr, bp.data(), nullptr, nullptr));
  e.bd = &bd;
  e.bd_version = bd.version;
}

int degree_of(const std::shared_ptr<Function>& f) { return f->V->degree; }

void se_impl(std::vector<std::shared_ptr<Function>>& flux_hdiv, std::vector<std::shared_ptr<Function>>& flux_dg,
             std::vector<std::shared_ptr<Function>>& rhs_dg, std::shared_ptr<BoundaryData> bd, bool stress, Function* korn)
{
  // input checks of se/reconstruction.hpp:336-388
  const int n_rhs = (int)rhs_dg.size(), n_hdiv = (int)flux_hdiv.size(), n_dg = (int)flux_dg.size();
  if (n_rhs == 0 || n_rhs != bd->num_rhs() || n_rhs != n_hdiv || n_rhs != n_dg)
    throw std::runtime_error("Equilibration: Input sizes does not match");
  const int k = degree_of(flux_hdiv[0]), p_flux = degree_of(flux_dg[0]), p_rhs = degree_of(rhs_dg[0]);
  if (p_rhs > k - 1 || p_flux > p_rhs)
    throw std::runtime_error("Equilibration: Wrong polynomial degree of the projected RHS");
  if (p_flux != p_rhs)
    throw std::runtime_error("Equilibration: Degrees of projected flux and RHS have to match");
  if (stress)
  {
    if (n_rhs < 2)
      throw std::runtime_error("Stress equilibration: Specify all rows of stress tensor");
    if (n_hdiv < 2)
      throw std::runtime_error("Stress equilibration: RT_k with k>1 required!");
  }
  if (flux_hdiv[0]->V->family != "DRT")
    throw std::runtime_error("Equilibration: the semi-explicit flux lives in the discontinuous RT space");
  Mesh& m = *flux_hdiv[0]->V->mesh;
  auto e = get_handle(m, k, p_rhs, n_rhs, (stress ? EQLB_FLAG_STRESS : 0u) | EQLB_FLAG_HOST_PIPELINE);
  ensure_bcs(*e, *bd, m, k);
  std::vector<const double*> G(n_rhs), F(n_rhs);
  std::vector<double*> S(n_rhs);
  for (int r = 0; r < n_rhs; ++r)
  {
    G[r] = flux_dg[r]->x.data();
    F[r] = rhs_dg[r]->x.data();
    S[r] = flux_hdiv[r]->x.mutable_data();
    flux_dg[r]->note_host_call();
    rhs_dg[r]->note_host_call();
    flux_hdiv[r]->note_host_call();
  }
  py::gil_scoped_release rel;
  check(eqlb_se_run(e->h, G.data(), F.data(), S.data(), korn ? korn->x.mutable_data() : nullptr, EQLB_HOST));
}
} // namespace

PYBIND11_MODULE(cpp, mod)
{
  mod.doc() = "B200-native successor of dolfinx_eqlb.cpp (python/dolfinx_eqlb/wrappers.cpp): C++ host layer over the C ABI "
              "of libeqlb_b200.so";

  py::class_<Mesh, std::shared_ptr<Mesh>>(mod, "Mesh", "Mesh arrays the hot path reads (SURVEY 8b)")
      .def(py::init<darray, iarray, iarray, iarray, iarray, iarray, iarray, iarray, iarray, iarray,
                    py::array_t<std::uint8_t, py::array::c_style | py::array::forcecast>,
                    py::array_t<std::uint32_t, py::array::c_style | py::array::forcecast>>(),
           py::arg("x"), py::arg("cell_node"), py::arg("cell_fct"), py::arg("fct_node"), py::arg("fct_cell_off"),
           py::arg("fct_cell"), py::arg("node_cell_off"), py::arg("node_cell"), py::arg("node_fct_off"), py::arg("node_fct"),
           py::arg("fct_perms"), py::arg("cell_perm_info"))
      .def_readonly("nnode", &Mesh::nnode)
      .def_readonly("ncell", &Mesh::ncell)
      .def_readonly("nfct", &Mesh::nfct)
      .def_property_readonly("num_cached_handles", &Mesh::num_cached_handles)
      .def("clear_cache", &Mesh::clear_cache);

  py::class_<FunctionSpace, std::shared_ptr<FunctionSpace>>(mod, "FunctionSpace")
      .def(py::init<std::shared_ptr<Mesh>, std::string, int, int>(), py::arg("mesh"), py::arg("family"), py::arg("degree"),
           py::arg("block_size") = 1)
      .def_readonly("mesh", &FunctionSpace::mesh)
      .def_readonly("family", &FunctionSpace::family)
      .def_readonly("degree", &FunctionSpace::degree)
      .def_property_readonly("size", &FunctionSpace::size);

  py::class_<Function, std::shared_ptr<Function>>(mod, "Function")
      .def(py::init<std::shared_ptr<FunctionSpace>, py::object>(), py::arg("function_space"), py::arg("array") = py::none())
      .def_readonly("function_space", &Function::V)
      .def_readwrite("name", &Function::name)
      .def_property_readonly("x", [](Function& f) { return f.x; });

  py::class_<Form, std::shared_ptr<Form>>(mod, "Form", "Fixed forms of the hot path (FluxEqlbEV.py:116-133, projection.py:17-77)")
      .def_static("ev_bilinear",
                  [](std::shared_ptr<FunctionSpace> V)
                  {
                    auto f = std::make_shared<Form>();
                    f->kind = "ev_a";
                    f->V = V;
                    return f;
                  })
      .def_static("ev_penalty",
                  [](std::shared_ptr<FunctionSpace> V)
                  {
                    auto f = std::make_shared<Form>();
                    f->kind = "ev_lpen";
                    f->V = V;
                    return f;
                  })
      .def_static("ev_linear",
                  [](std::shared_ptr<FunctionSpace> V, std::shared_ptr<Function> flux_dg, std::shared_ptr<Function> rhs_dg)
                  {
                    auto f = std::make_shared<Form>();
                    f->kind = "ev_l";
                    f->V = V;
                    f->flux_dg = flux_dg;
                    f->rhs_dg = rhs_dg;
                    return f;
                  })
      .def_static("mass",
                  [](std::shared_ptr<FunctionSpace> V)
                  {
                    auto f = std::make_shared<Form>();
                    f->kind = "mass";
                    f->V = V;
                    return f;
                  })
      .def_static("projection_rhs",
                  [](std::shared_ptr<FunctionSpace> V, darray qvals)
                  {
                    auto f = std::make_shared<Form>();
                    f->kind = "projection_rhs";
                    f->V = V;
                    f->qvals = qvals;
                    return f;
                  })
      .def_readonly("kind", &Form::kind);

  py::class_<FluxBC, std::shared_ptr<FluxBC>>(mod, "FluxBC", "FluxBC object (wrappers.cpp:144-232)")
      .def(py::init(
               [](std::shared_ptr<FunctionSpace> V, iarray facets, std::uintptr_t fn_addr, int nevals,
                  std::vector<std::shared_ptr<Function>> coefficients, std::vector<int> positions, std::vector<darray> constants)
               {
                 auto b = std::make_shared<FluxBC>();
                 b->V = V;
                 b->facets = facets;
                 b->kernel = reinterpret_cast<bc_kernel_t>(fn_addr);
                 b->nevals = nevals;
                 b->coefficients = coefficients;
                 b->positions = positions;
                 b->constants = constants;
                 return b;
               }),
           py::arg("function_space"), py::arg("facets"), py::arg("pointer_boundary_kernel"), py::arg("nevals_per_fct"),
           py::arg("coefficients"), py::arg("position_of_coefficients"), py::arg("constants"))
      .def(py::init(
               [](std::shared_ptr<FunctionSpace> V, iarray facets, std::uintptr_t fn_addr, int nevals, int qdeg,
                  std::vector<std::shared_ptr<Function>> coefficients, std::vector<int> positions, std::vector<darray> constants)
               {
                 auto b = std::make_shared<FluxBC>();
                 b->V = V;
                 b->facets = facets;
                 b->kernel = reinterpret_cast<bc_kernel_t>(fn_addr);
                 b->nevals = nevals;
                 b->quadrature_degree = qdeg;
                 b->projection = true;
                 b->coefficients = coefficients;
                 b->positions = positions;
                 b->constants = constants;
                 return b;
               }),
           py::arg("function_space"), py::arg("facets"), py::arg("pointer_boundary_kernel"), py::arg("nevals_per_fct"),
           py::arg("quadrature_degree"), py::arg("coefficients"), py::arg("position_of_coefficients"), py::arg("constants"))
      .def(py::init(
               [](std::shared_ptr<FunctionSpace> V, iarray facets, darray coeffs)
               {
                 auto b = std::make_shared<FluxBC>();
                 b->V = V;
                 b->facets = facets;
                 if (coeffs.ndim() != 2 || coeffs.shape(0) != facets.size())
                   throw std::runtime_error("FluxBC: one coefficient row per facet required");
                 b->poly = coeffs;
                 b->is_poly = true;
                 return b;
               }),
           py::arg("function_space"), py::arg("facets"), py::arg("coefficients"),
           "polynomial traction: outward normal flux sum_j coefficients[i, j] s^j on facet i (evaluated on the device)")
      .def_property_readonly("quadrature_degree", [](const FluxBC& b) { return b.quadrature_degree; });

  py::class_<BoundaryData, std::shared_ptr<BoundaryData>>(mod, "BoundaryData", "BoundaryData object (wrappers.cpp:235-256)")
      .def(py::init(
               [](std::vector<std::vector<std::shared_ptr<FluxBC>>> list_bcs, std::vector<std::shared_ptr<Function>> bfuncs,
                  std::shared_ptr<FunctionSpace> V, bool custom, int qdeg, std::vector<iarray> prime, bool stress)
               {
                 if (list_bcs.size() != bfuncs.size() || list_bcs.size() != prime.size())
                   throw std::runtime_error("Size of input data does not match!");
                 auto b = std::make_shared<BoundaryData>();
                 b->bcs = list_bcs;
                 b->bfuncs = bfuncs;
                 b->V = V;
                 b->custom = custom;
                 b->qdegree = qdeg;
                 b->prime = prime;
                 b->stress = stress;
                 return b;
               }),
           py::arg("list_of_bcs"), py::arg("list_of_boundary_fluxes"), py::arg("V_flux_hdiv"), py::arg("rtflux_is_custom"),
           py::arg("quadrature_degree"), py::arg("list_bfcts_prime"), py::arg("reconstruct_stress"))
      .def_property_readonly("num_rhs", &BoundaryData::num_rhs)
      .def("touch", [](BoundaryData& b) { ++b.version; }, "mark the object as modified (boundary data are rebuilt on the next call)");

  mod.def(
      "reconstruct_fluxes_minimisation",
      [](std::shared_ptr<Form> a, std::shared_ptr<Form> l_pen, std::vector<std::shared_ptr<Form>> l,
         std::vector<std::shared_ptr<Function>> flux_hdiv, std::shared_ptr<BoundaryData> bd)
      {
        // input checks of ev/reconstruction.hpp:158-166
        const int n_rhs = (int)l.size();
        if (n_rhs == 0 || n_rhs != bd->num_rhs() || n_rhs != (int)flux_hdiv.size())
          throw std::runtime_error("Equilibration: Input sizes does not match");
        if (a->kind != "ev_a" || l_pen->kind != "ev_lpen")
          throw std::runtime_error("Equilibration: only the fixed forms of FluxEqlbEV (Form.ev_bilinear / ev_penalty / ev_linear) "
                                   "can be evaluated on the device");
        const int k = degree_of(flux_hdiv[0]);
        if (flux_hdiv[0]->V->family != "RT")
          throw std::runtime_error("Equilibration: the constrained-minimisation flux lives in the conforming RT space");
        Mesh& m = *flux_hdiv[0]->V->mesh;
        std::vector<const double*> G(n_rhs), F(n_rhs);
        std::vector<double*> S(n_rhs);
        for (int r = 0; r < n_rhs; ++r)
        {
          if (l[r]->kind != "ev_l" || !l[r]->flux_dg || !l[r]->rhs_dg)
            throw std::runtime_error("Equilibration: linear forms have to be Form.ev_linear(V, flux_dg, rhs_dg)");
          G[r] = l[r]->flux_dg->x.data();
          F[r] = l[r]->rhs_dg->x.data();
          S[r] = flux_hdiv[r]->x.mutable_data();
          l[r]->flux_dg->note_host_call();
          l[r]->rhs_dg->note_host_call();
          flux_hdiv[r]->note_host_call();
        }
        const int p = degree_of(l[0]->rhs_dg);
        auto e = get_handle(m, k, p, n_rhs, EQLB_FLAG_HOST_PIPELINE);
        ensure_bcs(*e, *bd, m, k);
        py::gil_scoped_release rel;
        check(eqlb_ev_run(e->h, G.data(), F.data(), S.data(), EQLB_HOST));
      },
      py::arg("a"), py::arg("l_pen"), py::arg("l"), py::arg("flux_hdiv"), py::arg("boundary_data"),
      "Local equilibration of H(div) conforming fluxes, solving patch-wise constrained minimisation problems.");

  mod.def(
      "reconstruct_fluxes_semiexplt",
      [](std::vector<std::shared_ptr<Function>> flux_hdiv, std::vector<std::shared_ptr<Function>> flux_dg,
         std::vector<std::shared_ptr<Function>> rhs_dg, std::shared_ptr<BoundaryData> bd, bool stress)
      { se_impl(flux_hdiv, flux_dg, rhs_dg, bd, stress, nullptr); },
      py::arg("flux_hdiv"), py::arg("flux_dg"), py::arg("rhs_dg"), py::arg("boundary_data"), py::arg("reconstruct_stress"),
      "Local equilibration of H(div) conforming fluxes, using an explicit determination of the flues alongside with an "
      "unconstrained minimisation problem on a reduced space.");

  mod.def(
      "reconstruct_fluxes_semiexplt_with_kornconst",
      [](std::vector<std::shared_ptr<Function>> flux_hdiv, std::vector<std::shared_ptr<Function>> flux_dg,
         std::vector<std::shared_ptr<Function>> rhs_dg, std::shared_ptr<BoundaryData> bd, bool stress,
         std::shared_ptr<Function> korn)
      {
        if (!korn || korn->V->family != "DG" || korn->V->degree != 0)
          throw std::runtime_error("Korn constants live in a DG0 function");
        se_impl(flux_hdiv, flux_dg, rhs_dg, bd, stress, korn.get());
      },
      py::arg("flux_hdiv"), py::arg("flux_dg"), py::arg("rhs_dg"), py::arg("boundary_data"), py::arg("reconstruct_stress"),
      py::arg("cells_kornconst"),
      "Local equilibration of H(div) conforming fluxes (semi-explicit) with estimation of the cells Korn constants.");

  auto local_solver = [](std::vector<std::shared_ptr<Function>> solution, std::shared_ptr<Form> a, std::vector<std::shared_ptr<Form>> l)
  {
    // base/local_solver.hpp:38-187 for the fixed forms of lsolver/projection.py:17-77: a = mass form of a DG_p
    // space, l[i] = (f_i, v) given by the values of f_i at the cell quadrature points
    if (a->kind != "mass" || a->V->family != "DG")
      throw std::runtime_error("local_solver: only the mass form of a DG space (Form.mass) can be evaluated on the device");
    if (solution.size() != l.size())
      throw std::runtime_error("Local solver: Input sizes does not match");
    const int p = a->V->degree, k = p + 1, nfun = (int)l.size();
    Mesh& m = *a->V->mesh;
    auto e = get_handle(m, k, p, 1, 0u);
    std::vector<const double*> q(nfun);
    std::vector<double*> out(nfun);
    for (int i = 0; i < nfun; ++i)
    {
      if (l[i]->kind != "projection_rhs")
        throw std::runtime_error("local_solver: linear forms have to be Form.projection_rhs(V, values_at_quadrature_points)");
      q[i] = l[i]->qvals.data();
      out[i] = solution[i]->x.mutable_data();
    }
    py::gil_scoped_release rel;
    check(eqlb_local_project(e->h, nfun, q.data(), out.data(), EQLB_HOST));
  };
  mod.def("local_solver_lu", local_solver, py::arg("solution"), py::arg("a"), py::arg("l"), "Local solver based on the LU decomposition");
  mod.def("local_solver_cholesky", local_solver, py::arg("solution"), py::arg("a"), py::arg("l"),
          "Local solver based on the Cholesky decomposition");
  mod.def("local_solver_cg", local_solver, py::arg("solution"), py::arg("a"), py::arg("l"), "Local solver based on a CG solver");
  mod.def("version", []() { return std::string(eqlb_version()); });
}
