// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
// Stand-in for the parts of Basix 0.6 (third party, absent from /root/reference and from
// this image) that the reference hot path calls, so that the UNCHANGED reference sources
// compile (oracle/Makefile, target _ref).  Elements are polynomial tables:
//   * the hierarchic RT element comes from EXECUTING the reference's own
//     `elmtlib/e_raviart_thomas.py` (tests/golden/make_ref_element.py) - monomial
//     coefficients, interpolation points and interpolation matrix are handed in by the
//     driver (`FiniteElement::from_monomials`);
//   * Lagrange P_p / DG_p: equispaced nodal basis in Basix' DOF order (vertices, edge
//     interiors e0 e1 e2 low->high vertex, interior), identical to GLL-warped for p <= 2;
//   * quadrature: Gauss-Jacobi rules of the requested exactness.  Basix' default triangle
//     rule is a Xiao-Gimbutas table (different points); every integrand on the hot path
//     is a polynomial of at most that degree, so sums agree to rounding.
// Reference-cell conventions as in Basix: triangle (0,0),(1,0),(0,1); facet f opposite
// vertex f with vertices {1,2},{0,2},{0,1}; facet_normals = rotated tangents
// (-1,-1)/sqrt2, (-1,0), (0,1); facet_orientations: see cell::facet_orientations below.
#pragma once

#include "mdspan.hpp"

#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <numeric>
#include <algorithm>
#include <span>
#include <stdexcept>
#include <utility>
#include <vector>

namespace basix
{
namespace cell
{
enum class type
{
  point = 0,
  interval = 1,
  triangle = 2,
  tetrahedron = 3
};

inline int topological_dimension(type t) { return static_cast<int>(t); }

inline int num_sub_entities(type t, int dim)
{
  if (t == type::triangle)
    return dim == 0 ? 3 : (dim == 1 ? 3 : 1);
  if (t == type::interval)
    return dim == 0 ? 2 : 1;
  throw std::runtime_error("basix shim: cell type");
}

inline type sub_entity_type(type t, int dim, int)
{
  if (dim == topological_dimension(t))
    return t;
  return static_cast<type>(dim);
}

inline std::vector<std::vector<std::vector<int>>> topology(type t)
{
  if (t == type::triangle)
    return {{{0}, {1}, {2}}, {{1, 2}, {0, 2}, {0, 1}}, {{0, 1, 2}}};
  if (t == type::interval)
    return {{{0}, {1}}, {{0, 1}}};
  throw std::runtime_error("basix shim: cell type");
}

inline std::pair<std::vector<double>, std::array<std::size_t, 2>> geometry(type t)
{
  if (t == type::triangle)
    return {{0, 0, 1, 0, 0, 1}, {3, 2}};
  if (t == type::interval)
    return {{0, 1}, {2, 1}};
  throw std::runtime_error("basix shim: cell type");
}

inline std::pair<std::vector<double>, std::array<std::size_t, 2>> sub_entity_geometry(type t, int dim, int index)
{
  auto [x, xs] = geometry(t);
  const std::vector<int> verts = topology(t)[dim][index];
  std::vector<double> out;
  for (int v : verts)
    for (std::size_t d = 0; d < xs[1]; ++d)
      out.push_back(x[v * xs[1] + d]);
  return {out, {verts.size(), xs[1]}};
}

// normal = tangent (x1 - x0) rotated by +90 degrees, normalised
inline std::pair<std::vector<double>, std::array<std::size_t, 2>> facet_normals(type t)
{
  if (t != type::triangle)
    throw std::runtime_error("basix shim: facet_normals");
  auto [x, xs] = geometry(t);
  auto facets = topology(t)[1];
  std::vector<double> n(6);
  for (int f = 0; f < 3; ++f)
  {
    const double tx = x[2 * facets[f][1]] - x[2 * facets[f][0]];
    const double ty = x[2 * facets[f][1] + 1] - x[2 * facets[f][0] + 1];
    const double nrm = std::sqrt(tx * tx + ty * ty);
    n[2 * f] = -ty / nrm;
    n[2 * f + 1] = tx / nrm;
  }
  return {n, {3, 2}};
}

// The reference stores this as `_fct_normal_out` ("reference normal is outward",
// `base/KernelData.cpp:61`, used for the DOF prefactors `se/solve_patch_semiexplt.hpp:395`).
// The flux-balance formulas of step 1 are only consistent if `true` means "the facet
// normal above points out of the cell" - for the triangle {false, true, false}.
// EQLB_REF_FLIP_ORIENT (test hook) returns the complement.
inline bool& flip_orientations_flag()
{
  static bool f = false;
  return f;
}

inline std::vector<bool> facet_orientations(type t)
{
  auto [n, ns] = facet_normals(t);
  auto [x, xs] = geometry(t);
  auto facets = topology(t)[1];
  std::vector<bool> out(3);
  for (int f = 0; f < 3; ++f)
  {
    const double dx = x[2 * facets[f][0]] - 1.0 / 3.0, dy = x[2 * facets[f][0] + 1] - 1.0 / 3.0;
    out[f] = (n[2 * f] * dx + n[2 * f + 1] * dy > 0) != flip_orientations_flag();
  }
  return out;
}

inline std::pair<std::vector<double>, std::array<std::size_t, 2>> facet_outward_normals(type t)
{
  auto [n, ns] = facet_normals(t);
  auto [x, xs] = geometry(t);
  auto facets = topology(t)[1];
  for (int f = 0; f < 3; ++f)
  {
    const double dx = x[2 * facets[f][0]] - 1.0 / 3.0, dy = x[2 * facets[f][0] + 1] - 1.0 / 3.0;
    if (n[2 * f] * dx + n[2 * f + 1] * dy < 0)
    {
      n[2 * f] = -n[2 * f];
      n[2 * f + 1] = -n[2 * f + 1];
    }
  }
  return {n, ns};
}
} // namespace cell

namespace element
{
enum class family
{
  custom = 0,
  P = 1,
  RT = 2
};
enum class lagrange_variant
{
  unset = 0,
  legendre = 1,
  gll_warped = 2,
  equispaced = 3
};
enum class dpc_variant
{
  unset = 0
};
} // namespace element

namespace maps
{
enum class type
{
  identity = 0,
  contravariantPiola = 2
};
}

namespace quadrature
{
enum class type
{
  Default = 0,
  gauss_jacobi = 1
};

// Gauss-Jacobi nodes/weights for weight (1-x)^alpha on [-1, 1] (Newton on the recurrence)
inline void gauss_jacobi(int m, double alpha, std::vector<double>& x, std::vector<double>& w)
{
  x.assign(m, 0.0);
  w.assign(m, 0.0);
  auto eval = [&](double t, double& p, double& dp)
  {
    // P_n^{(alpha,0)}(t) by the three-term recurrence
    double p0 = 1.0, p1 = 0.5 * (alpha + (alpha + 2.0) * t);
    if (m == 0)
    {
      p = 1.0;
      dp = 0.0;
      return;
    }
    for (int n = 1; n < m; ++n)
    {
      const double a = alpha, b = 0.0;
      const double c1 = 2.0 * (n + 1) * (n + a + b + 1) * (2 * n + a + b);
      const double c2 = (2 * n + a + b + 1) * (a * a - b * b);
      const double c3 = (2 * n + a + b) * (2 * n + a + b + 1) * (2 * n + a + b + 2);
      const double c4 = 2.0 * (n + a) * (n + b) * (2 * n + a + b + 2);
      const double p2 = ((c2 + c3 * t) * p1 - c4 * p0) / c1;
      p0 = p1;
      p1 = p2;
    }
    p = p1;
    // derivative: (2n+a+b)(1-t^2) P_n' = n(a-b-(2n+a+b)t) P_n + 2(n+a)(n+b) P_{n-1}
    const double n = m, a = alpha, b = 0.0;
    dp = (n * (a - b - (2 * n + a + b) * t) * p1 + 2.0 * (n + a) * (n + b) * p0) / ((2 * n + a + b) * (1.0 - t * t));
  };
  for (int i = 0; i < m; ++i)
  {
    double t = -std::cos((2.0 * i + 1.0) * M_PI / (2.0 * m)); // Chebyshev guess
    for (int it = 0; it < 100; ++it)
    {
      // deflated Newton
      double p, dp;
      eval(t, p, dp);
      double s = 0.0;
      for (int j = 0; j < i; ++j)
        s += 1.0 / (t - x[j]);
      const double dt = p / (dp - s * p);
      t -= dt;
      if (std::fabs(dt) < 1e-16)
        break;
    }
    x[i] = t;
  }
  // weights from the moments: solve the (small) Vandermonde-type system in long double via
  // w_i = int prod_{j != i} (t - x_j)/(x_i - x_j) (1-t)^alpha dt, evaluated with a fine
  // Gauss-Legendre rule would be circular; use the closed form instead:
  //   w_i = Gamma-factor / ((1 - x_i^2) [P_n'(x_i)]^2), factor = 2^(a+1) for b = 0
  for (int i = 0; i < m; ++i)
  {
    double p, dp;
    eval(x[i], p, dp);
    w[i] = std::pow(2.0, alpha + 1.0) / ((1.0 - x[i] * x[i]) * dp * dp);
  }
}

// {points (flattened, row major), weights}; rule exact to degree `deg` with
// m = (deg + 2) / 2 points per direction
inline std::array<std::vector<double>, 2> make_quadrature(type, cell::type ct, int deg)
{
  const int m = (deg + 2) / 2;
  std::vector<double> xl, wl;
  gauss_jacobi(m, 0.0, xl, wl);
  if (ct == cell::type::interval)
  {
    std::vector<double> pts(m), wts(m);
    for (int i = 0; i < m; ++i)
    {
      pts[i] = 0.5 + 0.5 * xl[i];
      wts[i] = 0.5 * wl[i];
    }
    return {pts, wts};
  }
  if (ct == cell::type::triangle)
  {
    std::vector<double> xa, wa;
    gauss_jacobi(m, 1.0, xa, wa);
    std::vector<double> pts, wts;
    for (int i = 0; i < m; ++i)
    {
      const double u = 0.5 * (1.0 + xa[i]);
      for (int j = 0; j < m; ++j)
      {
        const double v = 0.5 * (1.0 + xl[j]);
        pts.push_back(u);
        pts.push_back((1.0 - u) * v);
        wts.push_back(wa[i] * wl[j] * 0.125);
      }
    }
    return {pts, wts};
  }
  throw std::runtime_error("basix shim: make_quadrature cell type");
}
inline std::array<std::vector<double>, 2> make_quadrature(cell::type ct, int deg)
{
  return make_quadrature(type::Default, ct, deg);
}
} // namespace quadrature

/// Polynomial finite element given by monomial coefficients.
class FiniteElement
{
public:
  FiniteElement() = default;

  /// monomials degree-major, x^a y^b with a descending inside a degree (1D: s^a)
  static std::vector<std::array<int, 2>> monomials(int tdim, int degree)
  {
    std::vector<std::array<int, 2>> m;
    for (int d = 0; d <= degree; ++d)
    {
      if (tdim == 1)
        m.push_back({d, 0});
      else
        for (int a = d; a >= 0; --a)
          m.push_back({a, d - a});
    }
    return m;
  }

  /// coef [ndofs][value_size][nmono]
  static FiniteElement from_monomials(cell::type ct, int degree, int poly_degree, int value_size, int ndofs,
                                      std::vector<double> coef, bool discontinuous, element::lagrange_variant lv,
                                      maps::type map, std::vector<double> X, std::array<std::size_t, 2> Xshape,
                                      std::vector<double> M, std::array<std::size_t, 2> Mshape,
                                      std::vector<std::vector<std::vector<int>>> entity_dofs,
                                      std::vector<std::vector<std::vector<int>>> entity_closure_dofs)
  {
    FiniteElement e;
    e._ct = ct;
    e._degree = degree;
    e._pdeg = poly_degree;
    e._vs = value_size;
    e._ndofs = ndofs;
    e._coef = std::move(coef);
    e._disc = discontinuous;
    e._lv = lv;
    e._map = map;
    e._X = std::move(X);
    e._Xshape = Xshape;
    e._M = std::move(M);
    e._Mshape = Mshape;
    e._edofs = std::move(entity_dofs);
    e._ecdofs = std::move(entity_closure_dofs);
    e._mon = monomials(cell::topological_dimension(ct), poly_degree);
    if (e._coef.size() != (std::size_t)ndofs * value_size * e._mon.size())
      throw std::runtime_error("basix shim: coefficient table size");
    return e;
  }

  int degree() const { return _degree; }
  int dim() const { return _ndofs; }
  cell::type cell_type() const { return _ct; }
  element::lagrange_variant lagrange_variant() const { return _lv; }
  bool discontinuous() const { return _disc; }
  maps::type map_type() const { return _map; }
  int value_size() const { return _vs; }
  const std::vector<std::vector<std::vector<int>>>& entity_dofs() const { return _edofs; }
  const std::vector<std::vector<std::vector<int>>>& entity_closure_dofs() const { return _ecdofs; }
  bool dof_transformations_are_identity() const { return true; }

  std::pair<std::vector<double>, std::array<std::size_t, 2>> points() const { return {_X, _Xshape}; }
  std::pair<std::vector<double>, std::array<std::size_t, 2>> interpolation_matrix() const { return {_M, _Mshape}; }

  std::array<std::size_t, 4> tabulate_shape(std::size_t nd, std::size_t npts) const
  {
    const std::size_t tdim = cell::topological_dimension(_ct);
    std::size_t nder = 1;
    for (std::size_t i = 1; i <= tdim; ++i)
      nder = nder * (nd + i) / i;
    return {nder, npts, (std::size_t)_ndofs, (std::size_t)_vs};
  }

  /// out[der][pt][dof][comp]; derivative order as Basix: 2D (0,0),(1,0),(0,1)
  void tabulate(int nd, std::span<const double> x, std::array<std::size_t, 2> shape, std::span<double> out) const
  {
    if (nd > 1)
      throw std::runtime_error("basix shim: only first derivatives");
    const std::size_t npts = shape[0], tdim = shape[1];
    const auto ts = tabulate_shape(nd, npts);
    if (out.size() < ts[0] * ts[1] * ts[2] * ts[3])
      throw std::runtime_error("basix shim: tabulate storage");
    const std::size_t nm = _mon.size();
    std::vector<double> mv(nm), mx(nm), my(nm);
    for (std::size_t p = 0; p < npts; ++p)
    {
      const double xx = x[p * tdim], yy = (tdim > 1) ? x[p * tdim + 1] : 0.0;
      for (std::size_t m = 0; m < nm; ++m)
      {
        const int a = _mon[m][0], b = _mon[m][1];
        mv[m] = ipow(xx, a) * ipow(yy, b);
        mx[m] = (a > 0) ? a * ipow(xx, a - 1) * ipow(yy, b) : 0.0;
        my[m] = (b > 0) ? b * ipow(xx, a) * ipow(yy, b - 1) : 0.0;
      }
      for (int i = 0; i < _ndofs; ++i)
        for (int c = 0; c < _vs; ++c)
        {
          const double* cf = _coef.data() + ((std::size_t)i * _vs + c) * nm;
          double v = 0, vx = 0, vy = 0;
          for (std::size_t m = 0; m < nm; ++m)
          {
            v += cf[m] * mv[m];
            vx += cf[m] * mx[m];
            vy += cf[m] * my[m];
          }
          const std::size_t o = (p * _ndofs + i) * _vs + c;
          const std::size_t plane = ts[1] * ts[2] * ts[3];
          out[o] = v;
          if (nd == 1)
          {
            out[plane + o] = vx;
            if (tdim > 1)
              out[2 * plane + o] = vy;
          }
        }
    }
  }

  /// Basix `map_fn`: (u, U, J, detJ, K) -> u = J U / detJ row-wise (contravariant Piola)
  template <typename O, typename P, typename Q, typename R>
  std::function<void(O&, const P&, const Q&, double, const R&)> map_fn() const
  {
    if (_map == maps::type::contravariantPiola)
      return [](O& u, const P& U, const Q& J, double detJ, const R&)
      {
        for (std::size_t p = 0; p < U.extent(0); ++p)
          for (std::size_t i = 0; i < J.extent(0); ++i)
          {
            typename O::value_type acc = 0;
            for (std::size_t k = 0; k < J.extent(1); ++k)
              acc += J(i, k) * U(p, k);
            u(p, i) = acc / detJ;
          }
      };
    return [](O& u, const P& U, const Q&, double, const R&)
    {
      for (std::size_t p = 0; p < U.extent(0); ++p)
        for (std::size_t i = 0; i < U.extent(1); ++i)
          u(p, i) = U(p, i);
    };
  }

private:
  static double ipow(double x, int n)
  {
    double r = 1.0;
    for (int i = 0; i < n; ++i)
      r *= x;
    return r;
  }
  cell::type _ct = cell::type::triangle;
  int _degree = 0, _pdeg = 0, _vs = 1, _ndofs = 0;
  bool _disc = false;
  element::lagrange_variant _lv = element::lagrange_variant::unset;
  maps::type _map = maps::type::identity;
  std::vector<double> _coef, _X, _M;
  std::array<std::size_t, 2> _Xshape{0, 0}, _Mshape{0, 0};
  std::vector<std::vector<std::vector<int>>> _edofs, _ecdofs;
  std::vector<std::array<int, 2>> _mon;
};

namespace element
{
/// Equispaced nodal Lagrange element in Basix' DOF order.
inline FiniteElement create_lagrange(cell::type ct, int degree, lagrange_variant lv, bool discontinuous)
{
  const int tdim = cell::topological_dimension(ct);
  std::vector<std::array<double, 2>> nodes;
  std::vector<std::vector<std::vector<int>>> edofs(tdim + 1), ecdofs(tdim + 1);
  const int p = degree;
  if (ct == cell::type::triangle)
  {
    edofs[0].resize(3);
    edofs[1].resize(3);
    edofs[2].resize(1);
    if (p == 0)
    {
      nodes.push_back({1.0 / 3.0, 1.0 / 3.0});
      edofs[2][0] = {0};
    }
    else
    {
      const double v[3][2] = {{0, 0}, {1, 0}, {0, 1}};
      const int fv[3][2] = {{1, 2}, {0, 2}, {0, 1}};
      for (int i = 0; i < 3; ++i)
      {
        edofs[0][i] = {i};
        nodes.push_back({v[i][0], v[i][1]});
      }
      for (int f = 0; f < 3; ++f)
        for (int i = 1; i < p; ++i)
        {
          const double t = (double)i / p;
          edofs[1][f].push_back(nodes.size());
          nodes.push_back({v[fv[f][0]][0] + t * (v[fv[f][1]][0] - v[fv[f][0]][0]),
                           v[fv[f][0]][1] + t * (v[fv[f][1]][1] - v[fv[f][0]][1])});
        }
      for (int j = 1; j < p; ++j)
        for (int i = 1; i < p - j; ++i)
        {
          edofs[2][0].push_back(nodes.size());
          nodes.push_back({(double)i / p, (double)j / p});
        }
    }
    // closures
    ecdofs[0] = edofs[0];
    ecdofs[1].resize(3);
    ecdofs[2].resize(1);
    const int fv[3][2] = {{1, 2}, {0, 2}, {0, 1}};
    for (int f = 0; f < 3; ++f)
    {
      if (p > 0)
      {
        ecdofs[1][f].push_back(fv[f][0]);
        ecdofs[1][f].push_back(fv[f][1]);
      }
      for (int d : edofs[1][f])
        ecdofs[1][f].push_back(d);
    }
    for (std::size_t i = 0; i < nodes.size(); ++i)
      ecdofs[2][0].push_back(i);
  }
  else if (ct == cell::type::interval)
  {
    edofs[0].resize(2);
    edofs[1].resize(1);
    if (p == 0)
    {
      nodes.push_back({0.5, 0});
      edofs[1][0] = {0};
    }
    else
    {
      nodes.push_back({0, 0});
      nodes.push_back({1, 0});
      edofs[0][0] = {0};
      edofs[0][1] = {1};
      for (int i = 1; i < p; ++i)
      {
        edofs[1][0].push_back(nodes.size());
        nodes.push_back({(double)i / p, 0});
      }
    }
    ecdofs[0] = edofs[0];
    ecdofs[1].resize(1);
    for (std::size_t i = 0; i < nodes.size(); ++i)
      ecdofs[1][0].push_back(i);
  }
  else
    throw std::runtime_error("basix shim: create_lagrange cell type");

  const auto mon = FiniteElement::monomials(tdim, p);
  const std::size_t n = nodes.size();
  if (mon.size() != n)
    throw std::runtime_error("basix shim: lagrange size");
  // V[i][m] = mono_m(node_i); coefficients C = V^-1 (columns = basis functions)
  std::vector<long double> V(n * n), C(n * n, 0.0L);
  for (std::size_t i = 0; i < n; ++i)
    for (std::size_t m = 0; m < n; ++m)
      V[i * n + m] = std::pow((long double)nodes[i][0], mon[m][0]) * std::pow((long double)nodes[i][1], mon[m][1]);
  for (std::size_t i = 0; i < n; ++i)
    C[i * n + i] = 1.0L;
  for (std::size_t c = 0; c < n; ++c)
  {
    std::size_t piv = c;
    for (std::size_t r = c + 1; r < n; ++r)
      if (std::fabs(V[r * n + c]) > std::fabs(V[piv * n + c]))
        piv = r;
    for (std::size_t j = 0; j < n; ++j)
    {
      std::swap(V[c * n + j], V[piv * n + j]);
      std::swap(C[c * n + j], C[piv * n + j]);
    }
    const long double inv = 1.0L / V[c * n + c];
    for (std::size_t j = 0; j < n; ++j)
    {
      V[c * n + j] *= inv;
      C[c * n + j] *= inv;
    }
    for (std::size_t r = 0; r < n; ++r)
      if (r != c && V[r * n + c] != 0.0L)
      {
        const long double f = V[r * n + c];
        for (std::size_t j = 0; j < n; ++j)
        {
          V[r * n + j] -= f * V[c * n + j];
          C[r * n + j] -= f * C[c * n + j];
        }
      }
  }
  // C[m][i] = coefficient of monomial m in basis function i
  std::vector<double> coef(n * n);
  for (std::size_t i = 0; i < n; ++i)
    for (std::size_t m = 0; m < n; ++m)
      coef[i * n + m] = (double)C[m * n + i];
  if (discontinuous)
  {
    // all DOFs belong to the cell interior
    std::vector<std::vector<std::vector<int>>> e2(tdim + 1), c2(tdim + 1);
    for (int d = 0; d <= tdim; ++d)
    {
      e2[d].resize(edofs[d].size());
      c2[d].resize(edofs[d].size());
    }
    for (std::size_t i = 0; i < n; ++i)
    {
      e2[tdim][0].push_back(i);
      c2[tdim][0].push_back(i);
    }
    edofs = e2;
    ecdofs = c2;
  }
  std::vector<double> X;
  for (auto& nd : nodes)
    for (int d = 0; d < tdim; ++d)
      X.push_back(nd[d]);
  return FiniteElement::from_monomials(ct, degree, degree, 1, (int)n, coef, discontinuous, lv, maps::type::identity, X,
                                       {n, (std::size_t)tdim}, {}, {0, 0}, edofs, ecdofs);
}
} // namespace element

inline FiniteElement create_element(element::family fam, cell::type ct, int degree, element::lagrange_variant lv,
                                    element::dpc_variant, bool discontinuous)
{
  if (fam != element::family::P)
    throw std::runtime_error("basix shim: create_element family");
  return element::create_lagrange(ct, degree, lv, discontinuous);
}
} // namespace basix
