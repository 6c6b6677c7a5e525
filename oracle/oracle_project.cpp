// TEST INFRASTRUCTURE - NOT PRODUCT CODE (see oracle_common.hpp).
// Restatement of the cell-wise local solver (base/local_solver.hpp:38-187) for the mass
// form of lsolver/projection.py:17-77: per cell A_e = (phi_i, phi_j), L_e = (phi_i, f),
// Cholesky (Eigen::LLT), solve, assignment into the DG vector.
#include "oracle_common.hpp"

extern "C" int oracle_local_project(const eqlb_mesh* mesh, const eqlb_tables* t, int nfun, const double* const* qvals,
                                    double* const* out)
{
  using namespace oracle;
  const int ndg = t->ndg, nq = t->nq;
  std::vector<double> A((size_t)ndg * ndg), L(ndg);
  for (int c = 0; c < mesh->ncell; ++c)
  {
    const int32_t* xd = mesh->cell_node + 3 * c;
    double J[4], K[4];
    const double detJ = compute_jacobian(J, K, mesh->x + 3 * xd[0], mesh->x + 3 * xd[1], mesh->x + 3 * xd[2]);
    std::fill(A.begin(), A.end(), 0.0);
    for (int q = 0; q < nq; ++q)
      for (int i = 0; i < ndg; ++i)
        for (int j = 0; j < ndg; ++j)
          A[i * ndg + j] += t->qwts[q] * std::fabs(detJ) * t->dg_q[(size_t)q * ndg + i] * t->dg_q[(size_t)q * ndg + j];
    llt_factor(A.data(), ndg, ndg);
    for (int f = 0; f < nfun; ++f)
    {
      std::fill(L.begin(), L.end(), 0.0);
      for (int q = 0; q < nq; ++q)
        for (int i = 0; i < ndg; ++i)
          L[i] += t->qwts[q] * std::fabs(detJ) * t->dg_q[(size_t)q * ndg + i] * qvals[f][(size_t)c * nq + q];
      llt_solve(A.data(), ndg, ndg, L.data());
      for (int i = 0; i < ndg; ++i)
        out[f][mesh->dg_dofmap[(size_t)c * ndg + i]] = L[i];
    }
  }
  return 0;
}
