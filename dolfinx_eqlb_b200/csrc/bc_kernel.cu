// Device-side construction of the boundary data the hot path reads:
// `base::BoundaryData` constructor, `base/BoundaryData.cpp:279-633`, for tractions given as
// polynomials on the boundary facets - the reference's interpolation branch (`:580-597`,
// `KernelDataBC::interpolate_flux :149-161,229-250`).  On an affine cell the interpolation of a flux
// with normal trace g (prescribed along the outward unit normal n) on facet E with local id f into
// the hierarchic RT element is closed form:
//     dof_j = int_E (g n) . n_ref^phys s^j ds = +-|E| sgn(det J) int_0^1 g(s) s^j ds,
// "+" iff the reference normal of f points outward (push-forward and pull-back cancel).
// With g(s) = sum_i c_i s^i:  int_0^1 g s^j = sum_i c_i / (i + j + 1).
// One thread per boundary facet; replaces the host loop of `dolfinx_eqlb_b200/eqlb.py::BoundaryData`.
#include "eqlb_internal.cuh"

namespace
{
__global__ void bc_prime_kernel(int n, const int32_t* __restrict__ fcts, int8_t* __restrict__ ftype)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    ftype[fcts[i]] = EQLB_FCT_ESSNT_PRIMAL;
}

__device__ __forceinline__ int local_facet(const MeshView& mv, int32_t c, int32_t f)
{
  int lf = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j)
    if (mv.cell_fct[3 * (size_t)c + j] == f)
      lf = j;
  return lf;
}

__global__ void bc_poly_kernel(MeshView mv, int k, int nrt, int n, const int32_t* __restrict__ fcts, int ncoef,
                               const double* __restrict__ coeffs, int8_t* __restrict__ ftype, double* __restrict__ bflux,
                               int8_t* __restrict__ local_fct_id, int32_t* __restrict__ node_cnt)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const int32_t f = fcts[i];
  const int32_t c = mv.fct_cell[mv.fct_cell_off[f]];
  const int lf = local_facet(mv, c, f);
  const int32_t n0 = mv.cell_node[3 * (size_t)c], n1 = mv.cell_node[3 * (size_t)c + 1], n2 = mv.cell_node[3 * (size_t)c + 2];
  const int32_t cn[3] = {n0, n1, n2};
  // facet f is opposite local vertex f: vertices {1,2}, {0,2}, {0,1}
  const int32_t va = cn[lf == 0 ? 1 : 0], vb = cn[lf == 2 ? 1 : 2];
  const double dx = mv.x[3 * (size_t)va] - mv.x[3 * (size_t)vb], dy = mv.x[3 * (size_t)va + 1] - mv.x[3 * (size_t)vb + 1];
  const double len = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
  const double J00 = mv.x[3 * (size_t)n1] - mv.x[3 * (size_t)n0], J01 = mv.x[3 * (size_t)n2] - mv.x[3 * (size_t)n0];
  const double J10 = mv.x[3 * (size_t)n1 + 1] - mv.x[3 * (size_t)n0 + 1], J11 = mv.x[3 * (size_t)n2 + 1] - mv.x[3 * (size_t)n0 + 1];
  const double det = __dsub_rn(__dmul_rn(J00, J11), __dmul_rn(J01, J10));
  const double sgn = det > 0.0 ? 1.0 : (det < 0.0 ? -1.0 : 0.0);
  const double pre = ((lf == 1) ? 1.0 : -1.0) * sgn;  // reference normal outward: {false, true, false}
  for (int j = 0; j < k; ++j)
  {
    double mom = 0.0;
    for (int q = 0; q < ncoef; ++q)
      mom = __dadd_rn(mom, coeffs[(size_t)i * ncoef + q] / (double)(q + j + 1));
    bflux[(size_t)c * nrt + lf * k + j] = __dmul_rn(__dmul_rn(pre, len), mom);
  }
  ftype[f] = EQLB_FCT_ESSNT_DUAL;
  local_fct_id[f] = (int8_t)lf;
  if (node_cnt)
  {
    atomicAdd(node_cnt + mv.fct_node[2 * (size_t)f], 1);
    atomicAdd(node_cnt + mv.fct_node[2 * (size_t)f + 1], 1);
  }
}

// node is on the essential stress boundary iff it was hit by 4 traction facets over the 2 rows
// (`base/BoundaryData.cpp:624-632`)
__global__ void bc_node_marker_kernel(int n, const int32_t* __restrict__ cnt, int8_t* __restrict__ marker)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    marker[i] = (cnt[i] == 4) ? 1 : 0;
}

// `local_fct_id` handed to eqlb_set_bcs has to be the cell-local id of every flux-BC facet
__global__ void bc_check_local_id_kernel(MeshView mv, int nrhs, const int8_t* __restrict__ ftype, const int8_t* __restrict__ lid,
                                         int* __restrict__ bad)
{
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= mv.nfct)
    return;
  bool dual = false;
  for (int r = 0; r < nrhs; ++r)
    dual |= ftype[(size_t)r * mv.nfct + f] == EQLB_FCT_ESSNT_DUAL;
  if (!dual)
    return;
  const int32_t c = mv.fct_cell[mv.fct_cell_off[f]];
  if (local_facet(mv, c, f) != lid[f])
    atomicAdd(bad, 1);
}
} // namespace

void launch_bc_poly(eqlb_handle* h, int r, int nprime, const int32_t* d_prime, int nb, const int32_t* d_fcts, int ncoef,
                    const double* d_coeffs, int8_t* d_local_fct_id, int32_t* d_node_cnt)
{
  const int bs = 128;
  int8_t* ft = h->d_facet_type.p + (size_t)r * h->nfct;
  if (nprime > 0)
  {
    bc_prime_kernel<<<(nprime + bs - 1) / bs, bs, 0, h->stream>>>(nprime, d_prime, ft);
    CUDA_CHECK(cudaGetLastError());
    h->launches++;
  }
  if (nb > 0)
  {
    bc_poly_kernel<<<(nb + bs - 1) / bs, bs, 0, h->stream>>>(h->mesh_view(), h->k, h->nrt, nb, d_fcts, ncoef, d_coeffs, ft,
                                                              h->d_bflux.p + (size_t)r * h->ncell * h->nrt, d_local_fct_id,
                                                              d_node_cnt);
    CUDA_CHECK(cudaGetLastError());
    h->launches++;
  }
}

void launch_bc_node_markers(eqlb_handle* h, const int32_t* d_cnt)
{
  const int bs = 256;
  bc_node_marker_kernel<<<(h->nnode + bs - 1) / bs, bs, 0, h->stream>>>(h->nnode, d_cnt, h->d_node_on_bnd.p);
  CUDA_CHECK(cudaGetLastError());
  h->launches++;
}

int count_bad_local_fct_ids(eqlb_handle* h, const int8_t* d_lid)
{
  DevBuf<int> bad;
  bad.alloc(1);
  bad.zero(h->stream);
  const int bs = 256;
  bc_check_local_id_kernel<<<(h->nfct + bs - 1) / bs, bs, 0, h->stream>>>(h->mesh_view(), h->nrhs, h->d_facet_type.p, d_lid, bad.p);
  CUDA_CHECK(cudaGetLastError());
  int out = 0;
  CUDA_CHECK(cudaMemcpyAsync(&out, bad.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return out;
}
