// Squared Korn constants of the patches (Kim 2011), reference semantics:
// OrientedPatch::estimate_squared_korn_constant `se/Patch.cpp:130-334`, accumulation
// x_korn[cell] += (dim+1) * c  `se/reconstruction.hpp:248-260`.
// Geometry only: thread per patch on the compact patch records, coloured launches.
#include "eqlb_internal.cuh"

namespace
{
__device__ __forceinline__ double angle_between(double ax, double ay, double bx, double by)
{
  return acos((ax * bx + ay * by) / (sqrt(ax * ax + ay * ay) * sqrt(bx * bx + by * by)));
}

template <int NCMAX>
__global__ void korn_kernel(PatchView pv, int first, int count, const double* __restrict__ x,
                            const int32_t* __restrict__ cell_node, double* __restrict__ korn, int use_atomics)
{
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= count)
    return;
  const size_t ip = (size_t)first + tid;
  const int nc = pv.ncells[ip];
  const int node = pv.node[ip];
  const bool internal = (pv.rhsinfo[ip] & 3) == EQLB_PATCH_INTERNAL;
  const double xz = x[3 * (size_t)node], yz = x[3 * (size_t)node + 1];
  int32_t cell[NCMAX];
  uint16_t info[NCMAX];
  for (int a = 0; a < nc; ++a)
  {
    cell[a] = pv.cell[(size_t)a * pv.stride + ip];
    info[a] = pv.info[(size_t)a * pv.stride + ip];
  }
  double theta_min;
  if (internal)
  {
    theta_min = 0.5 * M_PI;
    for (int a = 0; a < nc; ++a)
    {
      const int v = info[a] & 3;
      // boundary nodes of the cell in local vertex order
      const int l0 = (v == 0) ? 1 : 0, l1 = (v == 2) ? 1 : 2;
      const int32_t n0 = cell_node[3 * (size_t)cell[a] + l0], n1 = cell_node[3 * (size_t)cell[a] + l1];
      const double b0x = x[3 * (size_t)n0], b0y = x[3 * (size_t)n0 + 1];
      const double b1x = x[3 * (size_t)n1], b1y = x[3 * (size_t)n1 + 1];
      const double v2x = b1x - b0x, v2y = b1y - b0y;
      theta_min = fmin(theta_min, angle_between(xz - b0x, yz - b0y, v2x, v2y));
      theta_min = fmin(theta_min, angle_between(xz - b1x, yz - b1y, -v2x, -v2y));
    }
  }
  else
  {
    const int nf = nc + 1;
    // outer nodes of the facets E_0 .. E_nc
    double ox[NCMAX + 1], oy[NCMAX + 1];
    {
      const int v = info[0] & 3, fm = (info[0] >> 2) & 3;
      const int32_t n = cell_node[3 * (size_t)cell[0] + (3 - fm - v)];
      ox[0] = x[3 * (size_t)n];
      oy[0] = x[3 * (size_t)n + 1];
    }
    for (int a = 0; a < nc; ++a)
    {
      const int v = info[a] & 3, fp = (info[a] >> 4) & 3;
      const int32_t n = cell_node[3 * (size_t)cell[a] + (3 - fp - v)];
      ox[a + 1] = x[3 * (size_t)n];
      oy[a + 1] = x[3 * (size_t)n + 1];
    }
    auto centroid = [&](int a1, double& cx, double& cy)  // 1-based patch cell index
    {
      cx = 0.0;
      cy = 0.0;
      for (int j = 0; j < 3; ++j)
      {
        const int32_t n = cell_node[3 * (size_t)cell[a1 - 1] + j];
        cx += x[3 * (size_t)n] / 3;
        cy += x[3 * (size_t)n + 1] / 3;
      }
    };
    auto midpoint = [&](int f, double& cx, double& cy)
    {
      cx = 0.5 * xz + 0.5 * ox[f];
      cy = 0.5 * yz + 0.5 * oy[f];
    };
    double cnx[3], cny[3];
    if (nc % 2 == 0)
    {
      const int h = nc / 2;
      centroid(h, cnx[0], cny[0]);
      centroid(h + 1, cnx[1], cny[1]);
      midpoint(h, cnx[2], cny[2]);
    }
    else
    {
      const int h = nf / 2;
      midpoint(h, cnx[0], cny[0]);
      midpoint(h - 1, cnx[1], cny[1]);
      centroid(h, cnx[2], cny[2]);
    }
    double phi_min[3] = {M_PI, M_PI, M_PI};
    double xi = xz, yi = yz;
    double v2x = ox[nc] - xi, v2y = oy[nc] - yi;
    for (int i = 0; i < nf; ++i)
    {
      const double v3x = ox[i] - xi, v3y = oy[i] - yi;
      for (int j = 0; j < 3; ++j)
      {
        const double v1x = cnx[j] - xi, v1y = cny[j] - yi;
        phi_min[j] = fmin(phi_min[j], angle_between(v1x, v1y, v2x, v2y));
        phi_min[j] = fmin(phi_min[j], angle_between(v1x, v1y, v3x, v3y));
      }
      xi = ox[i];
      yi = oy[i];
      v2x = -v3x;
      v2y = -v3y;
    }
    theta_min = fmax(fmax(phi_min[0], phi_min[1]), phi_min[2]);
  }
  const double sh = sin(0.5 * theta_min);
  const double cks = 3.0 * 2.0 / (sh * sh);
  for (int a = 0; a < nc; ++a)
  {
    if (use_atomics)
      atomicAdd(korn + cell[a], cks);
    else
      korn[cell[a]] += cks;
  }
}
} // namespace

void launch_korn(eqlb_handle* h, double* dKorn)
{
  const PatchView pv = h->patch_view();
  const int bs = 128;
  auto kern = (h->ncmax <= 8) ? korn_kernel<8> : korn_kernel<EQLB_NCMAX>;
  const int ngrouped = h->h_group_off.empty() ? 0 : h->h_group_off.back();
  if (ngrouped > 0)
  {
    kern<<<(ngrouped + bs - 1) / bs, bs, 0, h->stream>>>(pv, 0, ngrouped, h->d_x.p, h->d_cell_node.p, dKorn, 1);
    CUDA_CHECK(cudaGetLastError());
    h->launches++;
  }
  const bool atomics = (h->flags & EQLB_FLAG_ATOMIC) != 0;
  if (atomics)
  {
    const int count = h->nactive - ngrouped;
    if (count > 0)
    {
      kern<<<(count + bs - 1) / bs, bs, 0, h->stream>>>(pv, ngrouped, count, h->d_x.p, h->d_cell_node.p, dKorn, 1);
      CUDA_CHECK(cudaGetLastError());
      h->launches++;
    }
    return;
  }
  // the launch-segment window of eqlb_set_part applies here like in launch_se (interface / interior patches)
  for (int c = std::max(0, h->win_lo); c < std::min(h->nseg, h->win_hi); ++c)
  {
    const int first = h->h_colour_off[c], count = h->h_colour_off[c + 1] - first;
    if (count == 0)
      continue;
    kern<<<(count + bs - 1) / bs, bs, 0, h->stream>>>(pv, first, count, h->d_x.p, h->d_cell_node.p, dKorn, 0);
    CUDA_CHECK(cudaGetLastError());
    h->launches++;
  }
}
