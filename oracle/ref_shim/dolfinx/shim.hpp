// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
// Stand-in for the handful of DOLFINx 0.6 types the reference hot path touches
// (mesh::Mesh/Topology/Geometry, graph::AdjacencyList, common::IndexMap, fem::FunctionSpace/
// FiniteElement/DofMap/Function/Form/CoordinateElement, la::Vector), so that the UNCHANGED
// reference sources under /root/reference/cpp compile here.  Plain data holders filled by
// oracle/ref_driver.cpp from the same arrays the product's C ABI receives (`eqlb_mesh`);
// written from the public DOLFINx interface, not copied.
#pragma once

#include <basix/finite-element.h>
#include <basix/mdspan.hpp>

#include <array>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <span>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace dolfinx
{
namespace common
{
class IndexMap
{
public:
  IndexMap(std::int32_t local, std::int32_t ghosts = 0) : _local(local), _ghosts(ghosts) {}
  std::int32_t size_local() const { return _local; }
  std::int32_t num_ghosts() const { return _ghosts; }

private:
  std::int32_t _local, _ghosts;
};
} // namespace common

namespace math
{
/// C += A B  (or A^T B^T when transpose), bounded by the extents of C
template <typename U, typename V, typename P>
void dot(const U& A, const V& B, P&& C, bool transpose = false)
{
  if (transpose)
  {
    for (std::size_t i = 0; i < C.extent(0); ++i)
      for (std::size_t j = 0; j < C.extent(1); ++j)
        for (std::size_t k = 0; k < A.extent(0); ++k)
          C(i, j) += A(k, i) * B(j, k);
  }
  else
  {
    for (std::size_t i = 0; i < C.extent(0); ++i)
      for (std::size_t j = 0; j < C.extent(1); ++j)
        for (std::size_t k = 0; k < A.extent(1); ++k)
          C(i, j) += A(i, k) * B(k, j);
  }
}
} // namespace math

namespace graph
{
template <typename T>
class AdjacencyList
{
public:
  AdjacencyList() = default;
  AdjacencyList(std::vector<T> data, std::vector<std::int32_t> offsets) : _data(std::move(data)), _off(std::move(offsets)) {}
  /// regular list with `width` links per node
  static AdjacencyList regular(const T* data, std::int32_t n, int width)
  {
    std::vector<std::int32_t> off(n + 1);
    for (std::int32_t i = 0; i <= n; ++i)
      off[i] = i * width;
    return AdjacencyList(std::vector<T>(data, data + (std::size_t)n * width), std::move(off));
  }
  std::span<const T> links(std::size_t i) const { return std::span<const T>(_data.data() + _off[i], _off[i + 1] - _off[i]); }
  int num_links(std::size_t i) const { return _off[i + 1] - _off[i]; }
  std::int32_t num_nodes() const { return (std::int32_t)_off.size() - 1; }
  const std::vector<T>& array() const { return _data; }
  const std::vector<std::int32_t>& offsets() const { return _off; }

private:
  std::vector<T> _data;
  std::vector<std::int32_t> _off;
};
} // namespace graph

namespace mesh
{
enum class CellType
{
  point = 0,
  interval = 1,
  triangle = 2,
  tetrahedron = 3
};
inline basix::cell::type cell_type_to_basix_type(CellType c) { return static_cast<basix::cell::type>(c); }
inline CellType cell_type_from_basix_type(basix::cell::type c) { return static_cast<CellType>(c); }
} // namespace mesh

namespace fem
{
/// Affine P1 triangle geometry map
class CoordinateElement
{
public:
  int dim() const { return 3; }
  bool is_affine() const { return true; }
  std::array<std::size_t, 4> tabulate_shape(std::size_t nd, std::size_t npts) const { return {nd == 0 ? 1u : 3u, npts, 3, 1}; }
  void tabulate(int nd, std::span<const double> X, std::array<std::size_t, 2> shape, std::span<double> basis) const
  {
    const std::size_t npts = shape[0];
    for (std::size_t p = 0; p < npts; ++p)
    {
      const double x = X[p * shape[1]], y = X[p * shape[1] + 1];
      basis[(0 * npts + p) * 3 + 0] = 1.0 - x - y;
      basis[(0 * npts + p) * 3 + 1] = x;
      basis[(0 * npts + p) * 3 + 2] = y;
      if (nd > 0)
      {
        const double dx[3] = {-1, 1, 0}, dy[3] = {-1, 0, 1};
        for (int i = 0; i < 3; ++i)
        {
          basis[(1 * npts + p) * 3 + i] = dx[i];
          basis[(2 * npts + p) * 3 + i] = dy[i];
        }
      }
    }
  }
  /// J = coords^T dphi^T
  template <typename U, typename V, typename W>
  static void compute_jacobian(const U& dphi, const V& cell_geometry, W&& J)
  {
    math::dot(cell_geometry, dphi, J, true);
  }
  template <typename U, typename V>
  static void compute_jacobian_inverse(const U& J, V&& K)
  {
    const double det = J(0, 0) * J(1, 1) - J(0, 1) * J(1, 0);
    K(0, 0) = J(1, 1) / det;
    K(0, 1) = -J(0, 1) / det;
    K(1, 0) = -J(1, 0) / det;
    K(1, 1) = J(0, 0) / det;
  }
  template <typename U>
  static double compute_jacobian_determinant(const U& J, std::span<double>)
  {
    return J(0, 0) * J(1, 1) - J(0, 1) * J(1, 0);
  }
};
} // namespace fem

namespace mesh
{
class Geometry
{
public:
  Geometry(int dim, graph::AdjacencyList<std::int32_t> dofmap, std::vector<double> x) : _dim(dim), _dofmap(std::move(dofmap)), _x(std::move(x)) {}
  int dim() const { return _dim; }
  const graph::AdjacencyList<std::int32_t>& dofmap() const { return _dofmap; }
  std::span<const double> x() const { return _x; }
  const fem::CoordinateElement& cmap() const { return _cmap; }

private:
  int _dim;
  graph::AdjacencyList<std::int32_t> _dofmap;
  std::vector<double> _x;
  fem::CoordinateElement _cmap;
};

class Topology
{
public:
  using AL = graph::AdjacencyList<std::int32_t>;
  Topology() : _conn(3, std::vector<std::shared_ptr<const AL>>(3)), _imap(3) {}
  int dim() const { return 2; }
  CellType cell_type() const { return CellType::triangle; }
  std::shared_ptr<const AL> connectivity(int d0, int d1) const { return _conn[d0][d1]; }
  std::shared_ptr<const common::IndexMap> index_map(int d) const { return _imap[d]; }
  const std::vector<std::uint8_t>& get_facet_permutations() const { return _fct_perms; }
  const std::vector<std::uint32_t>& get_cell_permutation_info() const { return _cell_perm_info; }
  void create_entity_permutations() const {}
  void create_connectivity(int, int) const {}

  void set_connectivity(std::shared_ptr<const AL> c, int d0, int d1) { _conn[d0][d1] = std::move(c); }
  void set_index_map(int d, std::shared_ptr<const common::IndexMap> m) { _imap[d] = std::move(m); }
  void set_permutations(std::vector<std::uint8_t> fp, std::vector<std::uint32_t> ci)
  {
    _fct_perms = std::move(fp);
    _cell_perm_info = std::move(ci);
  }

private:
  std::vector<std::vector<std::shared_ptr<const AL>>> _conn;
  std::vector<std::shared_ptr<const common::IndexMap>> _imap;
  std::vector<std::uint8_t> _fct_perms;
  std::vector<std::uint32_t> _cell_perm_info;
};

class Mesh
{
public:
  Mesh(Topology t, Geometry g) : _t(std::move(t)), _g(std::move(g)) {}
  const Topology& topology() const { return _t; }
  const Topology& topology_mutable() const { return _t; }
  const Geometry& geometry() const { return _g; }

private:
  Topology _t;
  Geometry _g;
};
} // namespace mesh

namespace la
{
template <typename T>
class Vector
{
public:
  /// non-owning view of a caller-owned array
  Vector(T* p, std::size_t n) : _p(p), _n(n) {}
  std::span<const T> array() const { return std::span<const T>(_p, _n); }
  std::span<T> mutable_array() { return std::span<T>(_p, _n); }

private:
  T* _p;
  std::size_t _n;
};
} // namespace la

namespace fem
{
namespace impl
{
template <typename T>
using scalar_value_type_t = T;
}
/// form kernel signature (FFCx `tabulate_tensor`)
template <class U, class T>
concept FEkernel = std::is_invocable_v<U, T*, const T*, const T*, const impl::scalar_value_type_t<T>*, const int*, const std::uint8_t*>;

enum class IntegralType : std::int8_t
{
  cell = 0,
  exterior_facet = 1,
  interior_facet = 2,
  vertex = 3
};

class ElementDofLayout
{
public:
  ElementDofLayout() = default;
  explicit ElementDofLayout(std::vector<std::vector<std::vector<int>>> entity_dofs) : _edofs(std::move(entity_dofs)) {}
  const std::vector<int>& entity_dofs(int dim, int i) const { return _edofs[dim][i]; }

private:
  std::vector<std::vector<std::vector<int>>> _edofs;
};

/// T = element transformation applied in place to `data` viewed as (ndofs x block)
/// [pre-apply] or (block x ndofs) [to-transpose]; identity unless the driver installs one.
using dof_transform_fn_d
    = std::function<void(const std::span<double>&, const std::span<const std::uint32_t>&, std::int32_t, int)>;

class FiniteElement
{
public:
  FiniteElement(basix::FiniteElement e, int block_size = 1) : _e(std::move(e)), _bs(block_size) {}
  const basix::FiniteElement& basix_element() const { return _e; }
  int space_dimension() const { return _space_dim >= 0 ? _space_dim : _e.dim() * _bs; }
  int block_size() const { return _bs; }
  bool needs_dof_transformations() const { return _needs_trafo; }

  template <typename T>
  std::function<void(const std::span<T>&, const std::span<const std::uint32_t>&, std::int32_t, int)>
  get_dof_transformation_function(bool inverse = false, bool transpose = false, bool /*scalar_element*/ = false) const
  {
    if (!_needs_trafo)
      return [](const std::span<T>&, const std::span<const std::uint32_t>&, std::int32_t, int) {};
    if constexpr (std::is_same_v<T, double>)
    {
      if (!inverse && !transpose)
        return _trafo;
      if (inverse && transpose)
        return _trafo_inv_t;
    }
    throw std::runtime_error("dolfinx shim: dof transformation variant not available");
  }
  template <typename T>
  std::function<void(const std::span<T>&, const std::span<const std::uint32_t>&, std::int32_t, int)>
  get_dof_transformation_to_transpose_function(bool /*inverse*/ = false, bool /*transpose*/ = false,
                                               bool /*scalar_element*/ = false) const
  {
    if (!_needs_trafo)
      return [](const std::span<T>&, const std::span<const std::uint32_t>&, std::int32_t, int) {};
    if constexpr (std::is_same_v<T, double>)
      return _trafo_to_t;
    throw std::runtime_error("dolfinx shim: dof transformation variant not available");
  }

  void set_space_dimension(int n) { _space_dim = n; }
  void set_transformations(dof_transform_fn_d t, dof_transform_fn_d t_to_t, dof_transform_fn_d t_inv_t)
  {
    _needs_trafo = true;
    _trafo = std::move(t);
    _trafo_to_t = std::move(t_to_t);
    _trafo_inv_t = std::move(t_inv_t);
  }

private:
  basix::FiniteElement _e;
  int _bs, _space_dim = -1;
  bool _needs_trafo = false;
  dof_transform_fn_d _trafo, _trafo_to_t, _trafo_inv_t;
};

class DofMap
{
public:
  DofMap(graph::AdjacencyList<std::int32_t> list, std::int32_t ndofs, int bs, ElementDofLayout layout)
      : index_map(std::make_shared<common::IndexMap>(ndofs)), _list(std::move(list)), _bs(bs), _layout(std::move(layout))
  {
  }
  const graph::AdjacencyList<std::int32_t>& list() const { return _list; }
  std::span<const std::int32_t> cell_dofs(int c) const { return _list.links(c); }
  int index_map_bs() const { return _bs; }
  int bs() const { return _bs; }
  const ElementDofLayout& element_dof_layout() const { return _layout; }
  std::shared_ptr<const common::IndexMap> index_map;

private:
  graph::AdjacencyList<std::int32_t> _list;
  int _bs;
  ElementDofLayout _layout;
};

class FunctionSpace
{
public:
  FunctionSpace(std::shared_ptr<const mesh::Mesh> m, std::shared_ptr<const FiniteElement> e, std::shared_ptr<const DofMap> d)
      : _mesh(std::move(m)), _element(std::move(e)), _dofmap(std::move(d))
  {
  }
  std::shared_ptr<const mesh::Mesh> mesh() const { return _mesh; }
  std::shared_ptr<const FiniteElement> element() const { return _element; }
  std::shared_ptr<const DofMap> dofmap() const { return _dofmap; }
  std::shared_ptr<const FunctionSpace> sub(const std::vector<int>& component) const { return _subs.at(component.at(0)); }
  bool contains(const FunctionSpace& V) const { return this == &V; }
  void add_sub(std::shared_ptr<const FunctionSpace> s) { _subs.push_back(std::move(s)); }

private:
  std::shared_ptr<const mesh::Mesh> _mesh;
  std::shared_ptr<const FiniteElement> _element;
  std::shared_ptr<const DofMap> _dofmap;
  std::vector<std::shared_ptr<const FunctionSpace>> _subs;
};

template <typename T>
class Function
{
public:
  Function(std::shared_ptr<const FunctionSpace> V, std::shared_ptr<la::Vector<T>> x) : _V(std::move(V)), _x(std::move(x)) {}
  std::shared_ptr<const FunctionSpace> function_space() const { return _V; }
  std::shared_ptr<la::Vector<T>> x() { return _x; }
  std::shared_ptr<const la::Vector<T>> x() const { return _x; }
  std::string name = "u";

private:
  std::shared_ptr<const FunctionSpace> _V;
  std::shared_ptr<la::Vector<T>> _x;
};

template <typename T>
class Constant
{
public:
  std::vector<T> value;
};

template <typename T>
class DirichletBC
{
public:
  std::shared_ptr<const FunctionSpace> function_space() const { return nullptr; }
  void mark_dofs(std::span<std::int8_t>) const {}
  void dof_values(std::span<T>) const {}
};

template <typename T>
class Form
{
public:
  using kernel_t = std::function<void(T*, const T*, const T*, const double*, const int*, const std::uint8_t*)>;
  Form(std::vector<std::shared_ptr<const FunctionSpace>> spaces, kernel_t cell_kernel,
       std::vector<std::shared_ptr<const Function<T>>> coefficients, std::vector<std::shared_ptr<const Constant<T>>> constants,
       std::shared_ptr<const mesh::Mesh> mesh)
      : _spaces(std::move(spaces)), _kernel(std::move(cell_kernel)), _coeffs(std::move(coefficients)),
        _consts(std::move(constants)), _mesh(std::move(mesh))
  {
    const std::int32_t nc = _mesh->topology().index_map(2)->size_local();
    _cells.resize(nc);
    for (std::int32_t i = 0; i < nc; ++i)
      _cells[i] = i;
  }
  std::shared_ptr<const mesh::Mesh> mesh() const { return _mesh; }
  const std::vector<std::shared_ptr<const FunctionSpace>>& function_spaces() const { return _spaces; }
  const kernel_t& kernel(IntegralType, int) const { return _kernel; }
  bool needs_facet_permutations() const { return false; }
  const std::vector<std::shared_ptr<const Function<T>>>& coefficients() const { return _coeffs; }
  const std::vector<std::shared_ptr<const Constant<T>>>& constants() const { return _consts; }
  std::vector<int> coefficient_offsets() const
  {
    std::vector<int> n{0};
    for (auto& c : _coeffs)
      n.push_back(n.back() + c->function_space()->element()->space_dimension());
    return n;
  }
  const std::vector<std::int32_t>& cell_domains(int) const { return _cells; }
  std::vector<int> integral_ids(IntegralType t) const { return t == IntegralType::cell ? std::vector<int>{-1} : std::vector<int>{}; }
  const std::vector<std::int32_t>& exterior_facet_domains(int) const { return _none; }
  const std::vector<std::int32_t>& interior_facet_domains(int) const { return _none; }

private:
  std::vector<std::shared_ptr<const FunctionSpace>> _spaces;
  kernel_t _kernel;
  std::vector<std::shared_ptr<const Function<T>>> _coeffs;
  std::vector<std::shared_ptr<const Constant<T>>> _consts;
  std::shared_ptr<const mesh::Mesh> _mesh;
  std::vector<std::int32_t> _cells, _none;
};

template <typename T>
std::map<std::pair<IntegralType, int>, std::pair<std::vector<T>, int>> allocate_coefficient_storage(const Form<T>& form)
{
  std::map<std::pair<IntegralType, int>, std::pair<std::vector<T>, int>> out;
  const int cstride = form.coefficient_offsets().back();
  out[{IntegralType::cell, -1}] = {std::vector<T>(form.cell_domains(-1).size() * (std::size_t)cstride), cstride};
  return out;
}

template <typename T>
std::vector<T> pack_constants(const Form<T>& form)
{
  std::vector<T> out;
  for (auto& c : form.constants())
    out.insert(out.end(), c->value.begin(), c->value.end());
  return out;
}

template <typename T>
std::map<std::pair<IntegralType, int>, std::pair<std::span<const T>, int>>
make_coefficients_span(const std::map<std::pair<IntegralType, int>, std::pair<std::vector<T>, int>>& coeffs)
{
  std::map<std::pair<IntegralType, int>, std::pair<std::span<const T>, int>> out;
  for (auto& [key, val] : coeffs)
    out[key] = {std::span<const T>(val.first), val.second};
  return out;
}

/// per cell: [coefficient 0 dofs (dof-major, block inner)] [coefficient 1 ...] - the layout
/// DOLFINx hands to FFCx kernels (no DOF transformations needed for the P1/DG coefficients
/// of the fixed EV forms)
template <typename T>
void pack_coefficients(const Form<T>& form, std::map<std::pair<IntegralType, int>, std::pair<std::vector<T>, int>>& coeffs)
{
  auto& [data, cstride] = coeffs.at({IntegralType::cell, -1});
  const std::vector<int> offs = form.coefficient_offsets();
  const auto& cells = form.cell_domains(-1);
  for (std::size_t ci = 0; ci < form.coefficients().size(); ++ci)
  {
    const auto& f = form.coefficients()[ci];
    const int bs = f->function_space()->element()->block_size();
    std::span<const T> x = f->x()->array();
    const auto& dm = f->function_space()->dofmap()->list();
    for (std::size_t c = 0; c < cells.size(); ++c)
    {
      auto dofs = dm.links(cells[c]);
      T* dst = data.data() + c * cstride + offs[ci];
      for (std::size_t j = 0; j < dofs.size(); ++j)
        for (int k = 0; k < bs; ++k)
          dst[j * bs + k] = x[bs * dofs[j] + k];
    }
  }
}
} // namespace fem
} // namespace dolfinx
