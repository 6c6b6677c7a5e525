// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
// CPU restatement ("oracle") of the dolfinx_eqlb hot path. Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may load this.  PARITY PINNING: PINNED against the reference's own code.  oracle/_ref/libeqlb_ref.so
// is the reference's C++ (se/, ev/, base/ sources and header templates) compiled UNCHANGED against the
// stand-in headers of oracle/ref_shim (oracle/Makefile target `ref`); tests/test_ref_pinning.py and
// tests/test_ref_bcs.py compare this restatement with it (integer maps bit-exact, DOF vectors <= 1e-11),
// the goldens under tests/golden/ are its outputs, and the RT element tables are checked against the
// reference's e_raviart_thomas.py executed on a stand-in basix module (tests/golden/ref_rt_element.npz).
// Restated rather than compiled (third party, absent): Basix quadrature points / Lagrange variants, FFCx
// form kernels, DOLFINx DOF transformations, Eigen - see DESIGN.md section 6 for what that leaves open.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/eqlb_b200.h"

namespace oracle
{

struct Links
{
  const int32_t* p;
  int n;
  int32_t operator[](int i) const { return p[i]; }
  int size() const { return n; }
  const int32_t* begin() const { return p; }
  const int32_t* end() const { return p + n; }
};

struct MeshView
{
  const eqlb_mesh* m;
  Links node_to_cell(int i) const
  {
    return {m->node_cell + m->node_cell_off[i], m->node_cell_off[i + 1] - m->node_cell_off[i]};
  }
  Links node_to_fct(int i) const
  {
    return {m->node_fct + m->node_fct_off[i], m->node_fct_off[i + 1] - m->node_fct_off[i]};
  }
  Links fct_to_cell(int i) const
  {
    return {m->fct_cell + m->fct_cell_off[i], m->fct_cell_off[i + 1] - m->fct_cell_off[i]};
  }
  Links fct_to_node(int i) const { return {m->fct_node + 2 * i, 2}; }
  Links cell_to_fct(int i) const { return {m->cell_fct + 3 * i, 3}; }
  Links cell_to_node(int i) const { return {m->cell_node + 3 * i, 3}; }
};

// wire values of base/Patch.hpp:20-33
enum PatchType : int8_t { internal = 0, bound_essnt_dual = 1, bound_essnt_primal = 2, bound_mixed = 3 };
enum PatchFacetType : int8_t { f_internal = 0, essnt_primal = 1, essnt_dual = 2 };

// Dense column-major-free helpers (row-major, leading dimension ld)
// Cholesky A = L L^T in place (lower), same role as Eigen::LLT (se/PatchData.hpp:831)
inline void llt_factor(double* A, int n, int ld)
{
  for (int j = 0; j < n; ++j)
  {
    double d = A[j * ld + j];
    for (int k = 0; k < j; ++k)
      d -= A[j * ld + k] * A[j * ld + k];
    d = std::sqrt(d);
    A[j * ld + j] = d;
    for (int i = j + 1; i < n; ++i)
    {
      double s = A[i * ld + j];
      for (int k = 0; k < j; ++k)
        s -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = s / d;
    }
  }
}

inline void llt_solve(const double* A, int n, int ld, double* b)
{
  for (int i = 0; i < n; ++i)
  {
    double s = b[i];
    for (int k = 0; k < i; ++k)
      s -= A[i * ld + k] * b[k];
    b[i] = s / A[i * ld + i];
  }
  for (int i = n - 1; i >= 0; --i)
  {
    double s = b[i];
    for (int k = i + 1; k < n; ++k)
      s -= A[k * ld + i] * b[k];
    b[i] = s / A[i * ld + i];
  }
}

// LU with partial pivoting in place, same role as Eigen::PartialPivLU
// (se/PatchData.hpp:832, ev/solve_patch.hpp:94)
inline void lu_factor(double* A, int n, int ld, int* piv)
{
  for (int k = 0; k < n; ++k)
  {
    int p = k;
    double mx = std::fabs(A[k * ld + k]);
    for (int i = k + 1; i < n; ++i)
    {
      double v = std::fabs(A[i * ld + k]);
      if (v > mx)
      {
        mx = v;
        p = i;
      }
    }
    piv[k] = p;
    if (p != k)
      for (int j = 0; j < n; ++j)
        std::swap(A[k * ld + j], A[p * ld + j]);
    double d = A[k * ld + k];
    for (int i = k + 1; i < n; ++i)
    {
      double l = A[i * ld + k] / d;
      A[i * ld + k] = l;
      for (int j = k + 1; j < n; ++j)
        A[i * ld + j] -= l * A[k * ld + j];
    }
  }
}

inline void lu_solve(const double* A, int n, int ld, const int* piv, double* b)
{
  for (int k = 0; k < n; ++k)
    if (piv[k] != k)
      std::swap(b[k], b[piv[k]]);
  for (int i = 0; i < n; ++i)
  {
    double s = b[i];
    for (int k = 0; k < i; ++k)
      s -= A[i * ld + k] * b[k];
    b[i] = s;
  }
  for (int i = n - 1; i >= 0; --i)
  {
    double s = b[i];
    for (int k = i + 1; k < n; ++k)
      s -= A[i * ld + k] * b[k];
    b[i] = s / A[i * ld + i];
  }
}

// base::KernelData::compute_jacobian (base/KernelData.cpp:66-90): affine P1 map,
// J = [x1-x0, x2-x0] (columns), K = J^-1, returns signed detJ
inline double compute_jacobian(double J[4], double K[4], const double* x0, const double* x1, const double* x2)
{
  J[0] = x1[0] - x0[0];
  J[1] = x2[0] - x0[0];
  J[2] = x1[1] - x0[1];
  J[3] = x2[1] - x0[1];
  const double det = J[0] * J[3] - J[1] * J[2];
  K[0] = J[3] / det;
  K[1] = -J[1] / det;
  K[2] = -J[2] / det;
  K[3] = J[0] / det;
  return det;
}

} // namespace oracle
