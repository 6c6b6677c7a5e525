// Internal declarations shared by the translation units of libeqlb_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/eqlb_b200.h"

#define EQLB_NCMAX 32      // hard upper bound of cells per patch supported by the kernels
#define EQLB_KMAX_WIDE 3   // ... patches with 17..32 cells: flux degrees <= 3 only (generic kernel frame size)
#define EQLB_MAXRHS 8  // max number of simultaneously equilibrated fluxes

// Reciprocal for the pivots / determinants of the patch kernels: hardware approximation
// (MUFU.RCP64H, relative error 2^-23) + two Newton steps = full double precision for normal,
// finite arguments in 5 instructions instead of the ~13 of the IEEE division sequence (no
// denormal / infinity handling: pivots of SPD patch systems and cell determinants are neither).
#ifdef __CUDACC__
__device__ __forceinline__ double eqlb_rcp(double a)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  return r;
}
#endif

// device pointers of the per-RHS vectors, passed by value as kernel parameter
struct RhsPtrs
{
  const double* G[EQLB_MAXRHS];
  const double* F[EQLB_MAXRHS];
  double* S[EQLB_MAXRHS];
};

// ---------------------------------------------------------------------------
// error handling: mirror the reference's `throw std::runtime_error` -> status code
// ---------------------------------------------------------------------------
struct EqlbError : std::runtime_error
{
  int code;
  EqlbError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CUDA_CHECK(expr)                                                                            \
  do                                                                                                \
  {                                                                                                 \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      throw EqlbError(EQLB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
  } while (0)

void eqlb_set_error(const std::string& msg);

// ---------------------------------------------------------------------------
// simple owning device buffer
// ---------------------------------------------------------------------------
// pageable host <-> device through the pinned staging pool (staged_copy.cu); blocking
bool host_is_pinned(const void* p);
void eqlb_h2d(void* dst_dev, const void* src_host, size_t bytes);
void eqlb_d2h(void* dst_host, const void* src_dev, size_t bytes);

// device memory from the stream-ordered pool (freed blocks stay cached: a new handle of the same process does not
// pay the 1-2 ms per cudaMalloc of a fresh 50-270 MB block again); staged_copy.cu
void* eqlb_dev_alloc(size_t bytes);
void eqlb_dev_free(void* p);

template <typename T>
struct DevBuf
{
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release()
  {
    if (p)
      eqlb_dev_free(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count)
  {
    if (count == n && p)
      return;
    release();
    if (count == 0)
      return;
    p = static_cast<T*>(eqlb_dev_alloc(count * sizeof(T)));
    n = count;
  }
  void upload(const T* host, size_t count)
  {
    alloc(count);
    if (count)
      eqlb_h2d(p, host, count * sizeof(T));
  }
  void zero(cudaStream_t s = 0)
  {
    if (n)
      CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
};

// ---------------------------------------------------------------------------
// device views
// ---------------------------------------------------------------------------

// Reference-matrix tables in one flat device array (doubles); offsets in doubles.
// Kernels copy the block into shared memory (a few KB) at CTA start.
struct TableView
{
  const double* data;  // device
  int ndoubles;
  int k, p, nrt, ndg, ndg_fct, ndiv, nadd;
  int o_rt_mass;     // [3][nrt][nrt]
  int o_fct_mom;     // [3][3][k][ndg]
  int o_cell_mom_f;  // [3][1+ndiv][ndg]
  int o_cell_mom_g;  // [3][1+ndiv][ndg][2]
  int o_bc_mat;      // [3][3][k][k]
  int o_trafo;       // [k][k]
  int o_rt_p1;       // [nrt][2][3]
  int o_dg_mono;     // [1+ndiv][ndg]
  int o_hat_dg_rt;   // [3][ndg][nrt][2]
  int o_mono_int;    // [1+ndiv]
};

// Mesh connectivity on the device (patch builder input)
struct MeshView
{
  int nnode, ncell, nfct;
  const double* x;  // [nnode*3]
  const int32_t *cell_node, *cell_fct, *fct_node, *fct_cell_off, *fct_cell, *node_cell_off, *node_cell, *node_fct_off,
      *node_fct;
  const uint8_t* fct_perms;
};

// Compact patch records, stored in colour-sorted order (position i <-> node order[i]).
// SoA with the patch index fastest so that a warp of consecutive patches loads
// coalesced.  info byte per (cell slot a, patch): bits 0-1 local id of the patch
// node in T_a, bits 2-3 local facet id of E_{a-1} in T_a, bits 4-5 local facet id of
// E_a in T_a, bit 6 reversed(a,0), bit 7 reversed(a,1), bit 8/9 facet-reflection bit
// (fct_perms) of E_{a-1} / E_a in T_a.
// rhs byte per (rhs, patch): bits 0-1 patch type, bit 2 reversion_required,
// bit 3 E_0 carries a flux BC, bit 4 E_n carries a flux BC.
struct PatchView
{
  int npatch;       // number of patches (== nnode)
  int ncmax;        // max cells per patch
  int nrhs;
  size_t stride;    // padded npatch (multiple of 32)
  const int32_t* node;   // [stride]     patch-central node
  const uint8_t* ncells; // [stride]
  const int32_t* cell;   // [ncmax][stride]
  const uint16_t* info;  // [ncmax][stride]
  const uint8_t* rhsinfo; // [nrhs][stride]
  // Lane records of the warp-cooperative kernels: one int4 per (patch, lane) with a
  // fixed number S of lanes per patch inside a launch segment, so that a warp reads 512
  // contiguous bytes and no load depends on another record load.
  // x = cell, y = info | ncells << 16, z = global facet E_{a-1}, w = global facet E_a
  // (lanes >= ncells: x = z = w = 0, info = 0).
  const int4* rec;
};

struct ColouringJob;  // patch_builder.cu

struct eqlb_handle
{
  int device = 0;
  cudaStream_t stream = 0;
  uint32_t flags = 0;
  int nrhs = 0;
  bool bcs_set = false;
  int64_t launches = 0;

  // sizes
  int nnode = 0, ncell = 0, nfct = 0;
  int k = 0, p = 0, nrt = 0, ndg = 0, ndg_fct = 0, ndiv = 0, nadd = 0, nq = 0, nqf = 0;
  int ncmax = 0;
  int nactive = 0;  // number of equilibrated patches (owned nodes)
  bool dg_identity = true;
  std::vector<uint8_t> h_owned;

  // mesh on device
  DevBuf<double> d_x;
  DevBuf<int32_t> d_cell_node, d_cell_fct, d_fct_node, d_fct_cell_off, d_fct_cell, d_node_cell_off, d_node_cell,
      d_node_fct_off, d_node_fct, d_dg_dofmap;
  DevBuf<uint8_t> d_fct_perms;
  DevBuf<double> d_cellJ;  // [ncell][4] Jacobians (J00,J01,J10,J11)

  // host topology used by the colouring (see eqlb_create)
  struct HostTopo
  {
    const int32_t *node_cell_off = nullptr, *node_cell = nullptr, *cell_node = nullptr, *node_fct_off = nullptr,
                  *node_fct = nullptr, *fct_node = nullptr;
  } topo;
  // host copies needed for grouping / recolouring (stress handles only)
  std::vector<int32_t> h_node_cell_off, h_node_cell, h_cell_node, h_node_fct_off, h_node_fct, h_fct_node;
  std::vector<uint8_t> h_grouped;      // node is member of a grouped boundary patch set
  std::vector<int32_t> h_group_off;    // offsets of the groups in h_order (first member = inner patch)

  // tables
  DevBuf<double> d_tables;
  TableView tv{};
  std::vector<double> h_tables_q;  // quadrature style tables needed by the projector (dg_q, qwts)
  DevBuf<double> d_proj;           // [ndg][nq] projection operator (mass^-1 * dg_q^T * w)
  DevBuf<double> d_primal;         // primal-space / estimator tables (offsets o_pr, see eqlb_create)
  int o_pr[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int npk = 0;
  DevBuf<int32_t> d_pk_dofmap;     // [ncell][npk] cell dofmap of the primal P_k space (eqlb_set_primal_space)
  int64_t pk_ndofs = 0;
  DevBuf<double> d_basix;          // [k][k] + [(k*k-k)][nrt] hierarchic RT -> Basix RT (Legendre) maps

  // boundary data
  DevBuf<int8_t> d_facet_type;  // [nrhs][nfct]
  DevBuf<double> d_bflux;       // [nrhs][ncell*nrt] (zeros where absent)
  std::vector<uint8_t> h_has_bflux;
  bool bflux_dummy = false;  // no flux-BC facet at all: d_bflux is a 2-element placeholder (never dereferenced)
  DevBuf<int8_t> d_node_on_bnd;
  DevBuf<int8_t> d_local_fct_id;  // [nfct] cell-local id of the flux-BC facets

  // patches
  std::vector<int32_t> h_order;       // colour-sorted node order
  std::vector<int32_t> h_colour_off;  // [ncolours+1]
  std::vector<int32_t> h_colour_fast; // [ncolours] number of leading patches of the colour eligible for the k=2 kernel
  DevBuf<double> d_k1tab;             // gathered tables of the degree-1 kernel
  DevBuf<double> d_k2tab;             // gathered tables of the k=2 streaming kernel
  DevBuf<double> d_kwtab;             // gathered tables of the general warp-cooperative kernel
  std::vector<int32_t> h_colour;      // [nnode]
  int ncolours = 0;
  bool coloured_with_groups = false;  // h_order currently starts with grouped boundary patches
  int nseg = 0;  // launch segments = spatial chunks x colours (h_colour_* arrays are per segment)
  size_t pstride = 0;
  DevBuf<int32_t> d_pnode, d_pcell;
  DevBuf<uint8_t> d_pncells, d_prhs;
  DevBuf<uint16_t> d_pinfo;
  DevBuf<int4> d_prec;                  // lane records of the eligible head of every segment
  // the eligible head of a segment is ordered by lane class (patches with <= 4 / 8 / 16 facets) and launched
  // class by class, so that one high-degree vertex does not widen the lanes of a whole colour
  struct FastSub
  {
    int32_t first, count, lanes;  // range in h_order, lanes per patch (4, 8, 16)
    int64_t recoff;               // offset of its lane records in d_prec
  };
  std::shared_ptr<ColouringJob> colouring_job;  // queued by eqlb_create, consumed by colour_patches
  DevBuf<int32_t> d_order;      // [nactive] patches in launch order (grouped patches first)
  DevBuf<uint16_t> d_rawkey;    // [nnode] ((chunk*64 + colour) << 2 | lane class) of the device launch order, 0xFFFF = not launched
  bool ordered = false;         // colouring + launch order are valid
  bool slabs_pending = false;   // result slabs of the host pipeline still to be computed from d_rawkey (ensure_slabs)
  std::vector<std::vector<FastSub>> h_seg_subs;  // [nseg] (empty: no lane-per-cell launch)
  int nsub = 0;
  DevBuf<int64_t> d_seginfo;            // [nsub][4] first, count, lanes, recoff

  // host pipeline (EQLB_FLAG_HOST_PIPELINE): spatial stages = chunks of the cell range
  int nchunk = 1;
  bool interface_first = false;  // chunk 0 = interface patches of a distributed run, chunk 1 = interior
  int part = EQLB_PART_ALL;
  int win_lo = 0, win_hi = 1 << 30;  // window of launch segments executed by launch_se / launch_ev
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_done;
  cudaEvent_t ev_start = nullptr;
  struct OutSlab
  {
    size_t off, len;  // range of the per-RHS result vector (doubles)
    int final_stage;  // last stage that adds into the range
  };
  std::vector<OutSlab> h_se_slabs, h_ev_slabs;
  ~eqlb_handle();

  // staging buffers for host-pointer calls
  DevBuf<double> d_stage_G, d_stage_f, d_stage_sigma, d_stage_korn;

  MeshView mesh_view() const;
  PatchView patch_view() const;
};

// kernels launchers (defined in the .cu files)
void launch_compute_cellJ(eqlb_handle* h);
// exact sequential first-fit colouring, computed on the device (patch_builder.cu)
struct ColouringJob;
const int* colouring_job_colours(const ColouringJob& job);  // device colours of a job
void device_order_histogram(eqlb_handle* h, const int* d_colour, int nchunk, std::vector<uint32_t>& hist);
void device_order_sort(eqlb_handle* h, const std::vector<uint16_t>& rank, int offset, int count);
void device_slab_stages(eqlb_handle* h, int nchunk, std::vector<int>& cfin, std::vector<int>& ffin);
std::shared_ptr<ColouringJob> device_greedy_colouring_start(eqlb_handle* h, const uint8_t* h_skip);
int device_greedy_colouring_finish(eqlb_handle* h, ColouringJob& job, std::vector<int32_t>& colour);
void launch_patch_builder(eqlb_handle* h, int32_t* d_ncells_out, int32_t* d_cells, int32_t* d_fcts, int8_t* d_inodes,
                          int8_t* d_fcts_local, int8_t* d_type, uint8_t* d_reversed, uint8_t* d_reversion);
void launch_se_dofmaps(eqlb_handle* h, int32_t* d_dofmap, int32_t* d_projflux, int8_t* d_bmarkers, int ndpc, int hzmax);
void launch_ev_dofmaps(eqlb_handle* h, int32_t* d_ncells, int32_t* d_cells, int32_t* d_fcts, int8_t* d_inod, int32_t* d_elmt,
                       int32_t* d_patch, int32_t* d_global, int32_t* d_lpatch, int32_t* d_lglobal);
void launch_se(eqlb_handle* h, const double* const* dG, const double* const* dF, double* const* dSigma, double* dKorn);
void launch_ev(eqlb_handle* h, const double* const* dG, const double* const* dF, double* const* dSigma);
void launch_project(eqlb_handle* h, int nfun, const double* const* dq, double* const* dout);
void launch_korn(eqlb_handle* h, double* dKorn);
void launch_bc_poly(eqlb_handle* h, int r, int nprime, const int32_t* d_prime, int nb, const int32_t* d_fcts, int ncoef,
                    const double* d_coeffs, int8_t* d_local_fct_id, int32_t* d_node_cnt);
void launch_bc_node_markers(eqlb_handle* h, const int32_t* d_cnt);
int count_bad_local_fct_ids(eqlb_handle* h, const int8_t* d_lid);
void launch_flux_norm(eqlb_handle* h, int nfun, const double* const* dsig, double* const* dout);
void launch_primal_project(eqlb_handle* h, int nfun, const double* const* uh, const double* const* fh, double* const* G,
                           double* const* F);
void launch_estimate_poisson(eqlb_handle* h, int nfun, const double* const* sigma, const double* const* uh, const double* const* fh,
                             double* const* e_sig, double* const* e_osc, int is_ev);
void launch_estimate_elasticity(eqlb_handle* h, const double* const* ds, const double* const* sh, const double* const* fh,
                                const double* korn, double pi_1, double* const* eta);
void launch_ev_to_basix(eqlb_handle* h, int nfun, const double* const* din, double* const* dout);
void build_k1_tables(eqlb_handle* h, const eqlb_tables* t);
void launch_k1(eqlb_handle* h, bool ev, const RhsPtrs& ptrs, int first, int count, int use_atomics, int lanes, int64_t recoff,
               bool pdl);
void build_k2_tables(eqlb_handle* h, const eqlb_tables* t);
bool kw_supported(int k, int ndg);
void build_kw_tables(eqlb_handle* h, const eqlb_tables* t);
void launch_kw(eqlb_handle* h, bool ev, const RhsPtrs& ptrs, int first, int count, int use_atomics, int lanes, int64_t recoff);
void launch_k2(eqlb_handle* h, bool ev, const RhsPtrs& ptrs, int first, int count, int use_atomics, int maxnf, int lanes,
               int64_t recoff, bool stress, bool pdl);
