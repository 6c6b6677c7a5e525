// Host side of the C ABI (include/eqlb_b200.h): validation, device residency of
// mesh / tables / boundary data, patch colouring, staging of host buffers.
// Mirrors the drivers se/reconstruction.hpp:337-407 and ev/reconstruction.hpp:157-177
// (input checks and error texts), everything numerical runs in CUDA kernels.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>

#include <chrono>
#include <exception>
#include <thread>

#include "eqlb_internal.cuh"

static thread_local std::string g_last_error;
void eqlb_set_error(const std::string& msg) { g_last_error = msg; }

void ensure_host_topology(eqlb_handle* h);  // host copies of the vertex connectivity, fetched from the device on first use

namespace
{

template <typename F>
int guarded(F&& f)
{
  try
  {
    f();
    return EQLB_OK;
  }
  catch (const EqlbError& e)
  {
    g_last_error = e.what();
    return e.code;
  }
  catch (const std::exception& e)
  {
    g_last_error = e.what();
    return EQLB_ERR_INPUT;
  }
}

void require_device()
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    throw EqlbError(EQLB_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ")
                                       + cudaGetErrorString(e));
}

// Greedy vertex colouring such that two patches of one colour never share a cell
// (two vertices of a common cell get different colours).  Deterministic: vertices
// in ascending order, smallest free colour.
// wall-clock breakdown of the setup calls on stderr when EQLB_TIMING is set
struct StageTimer
{
  const bool on = getenv("EQLB_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char* what)
  {
    if (!on)
      return;
    cudaDeviceSynchronize();
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[eqlb timing] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

// (segment, lane class) bookkeeping shared by the host and the device ordering: a lane class with few members
// joins the next wider one (saves a launch)
static void merge_small_classes(int32_t* cc, int8_t* cm)
{
  cm[0] = 0, cm[1] = 1, cm[2] = 2, cm[3] = 3;
  const int thr = std::max(8192, (cc[0] + cc[1] + cc[2]) / 16);
  if (cc[1] > 0 && cc[1] < thr && cc[2] > 0)
    cm[1] = 2, cc[2] += cc[1], cc[1] = 0;
  if (cc[0] > 0 && cc[0] < thr && cc[1] + cc[2] > 0)
  {
    cm[0] = cc[1] > 0 ? 1 : 2;
    cc[cm[0]] += cc[0], cc[0] = 0;
  }
}

// Launch order computed on the device (patch_builder.cu): raw keys + histogram, the host derives segment offsets,
// lane-class merging and the launch list from the 4096 counts, a stable radix sort writes h->d_order.
static void order_on_device(eqlb_handle* h, ColouringJob& job, int ncol, int nchunk, int ngrouped)
{
  std::vector<uint32_t> hist;
  device_order_histogram(h, colouring_job_colours(job), nchunk, hist);
  h->nseg = nchunk * ncol;
  h->h_colour_off.assign(h->nseg + 1, 0);
  h->h_colour_fast.assign(h->nseg, 0);
  h->h_seg_subs.assign(h->nseg, {});
  std::vector<uint16_t> rank(4096, 0xFFFF);
  int32_t at = ngrouped;  // grouped patches occupy the head of the order
  for (int chunk = 0; chunk < nchunk; ++chunk)
    for (int col = 0; col < ncol; ++col)
    {
      const int sg = chunk * ncol + col;
      const uint32_t* hh = &hist[(size_t)((chunk * 64 + col) << 2)];
      int32_t cc[3] = {(int32_t)hh[0], (int32_t)hh[1], (int32_t)hh[2]};
      int8_t cm[4];
      merge_small_classes(cc, cm);
      h->h_colour_off[sg] = at;
      for (int cl = 0; cl < 4; ++cl)
      {
        const int32_t first = at;
        for (int raw = 0; raw < 4; ++raw)
          if (cm[raw] == cl)
          {
            at += (int32_t)hh[raw];
            rank[(size_t)(((chunk * 64 + col) << 2) | raw)] = (uint16_t)(sg * 4 + cl);
          }
        if (cl < 3 && at > first)
        {
          h->h_seg_subs[sg].push_back({first, at - first, 4 << cl, -1});
          h->h_colour_fast[sg] += at - first;
        }
      }
    }
  h->h_colour_off[h->nseg] = at;
  if (at != h->nactive)
    throw EqlbError(EQLB_ERR_STATE, "device launch order: patch count mismatch");
  h->h_order.clear();
  device_order_sort(h, rank, ngrouped, at - ngrouped);
  h->h_se_slabs.clear();
  h->h_ev_slabs.clear();
  h->slabs_pending = nchunk > 1;
  h->ordered = true;
}

// result slabs of the host pipeline (a range of DOFs can go back to the host after the last stage with a patch
// that adds into it), computed on first use when the launch order came from the device
void ensure_slabs(eqlb_handle* h)
{
  if (!h->slabs_pending)
    return;
  const int nchunk = h->nchunk;
  std::vector<int> cfin, ffin;
  device_slab_stages(h, nchunk, cfin, ffin);
  auto slab_lo = [&](long cnt, int sidx) { return (size_t)(cnt * sidx / nchunk); };
  const int kk = h->k;
  const size_t ncd = (size_t)(kk * kk - kk);
  h->h_se_slabs.clear();
  h->h_ev_slabs.clear();
  for (int sidx = 0; sidx < nchunk; ++sidx)
  {
    const size_t c0 = slab_lo(h->ncell, sidx), c1 = slab_lo(h->ncell, sidx + 1);
    h->h_se_slabs.push_back({c0 * h->nrt, (c1 - c0) * h->nrt, cfin[sidx]});
    if (ncd)
      h->h_ev_slabs.push_back({(size_t)h->nfct * kk + c0 * ncd, (c1 - c0) * ncd, cfin[sidx]});
    const size_t f0 = slab_lo(h->nfct, sidx), f1 = slab_lo(h->nfct, sidx + 1);
    h->h_ev_slabs.push_back({f0 * kk, (f1 - f0) * kk, ffin[sidx]});
  }
  h->slabs_pending = false;
}

void colour_patches(eqlb_handle* h)
{
  // topology arrays: the caller's (during eqlb_create) or the handle's copies (stress handles,
  // which may have to recolour when a BC set groups boundary patches)
  if (h->nactive == 0 && h->nseg > 0)
    return;  // nothing owned: the (empty) colouring of eqlb_create stands
  const int n = h->nnode;
  int ncol = 0;
  const auto t_enter = std::chrono::steady_clock::now();
  const int ngrouped = h->h_group_off.empty() ? 0 : h->h_group_off.back();
  static const bool host_col = getenv("EQLB_HOST_COLOURING") && atoi(getenv("EQLB_HOST_COLOURING")) != 0;
  std::shared_ptr<ColouringJob> job;
  if (host_col || !h->d_node_cell.p || !h->d_cell_node.p)
  {
    // sequential first-fit colouring on the host (reference implementation of the device kernel)
    if (!h->topo.node_cell_off)
      ensure_host_topology(h);
    const eqlb_handle::HostTopo& T = h->topo;
    h->h_colour.assign(n, -1);
    for (int z = 0; z < n; ++z)
    {
      if (!h->h_owned[z] || h->h_grouped[z])
        continue;
      uint64_t used = 0;
      for (int i = T.node_cell_off[z]; i < T.node_cell_off[z + 1]; ++i)
      {
        const int32_t* cn = &T.cell_node[3 * (size_t)T.node_cell[i]];
        for (int j = 0; j < 3; ++j)
        {
          const int c = h->h_colour[cn[j]];
          if (c >= 0)
            used |= (uint64_t(1) << c);
        }
      }
      int c = 0;
      while (used & (uint64_t(1) << c))
        ++c;
      h->h_colour[z] = c;
      ncol = std::max(ncol, c + 1);
    }
  }
  else
  {
    // the same colouring, computed on the device from the uploaded connectivity (patch_builder.cu); eqlb_create
    // has queued the kernel already (h->colouring_job) and uploaded the rest of the mesh meanwhile
    job = h->colouring_job;
    h->colouring_job.reset();
    if (!job)
    {
      std::vector<uint8_t> skip;
      if (h->nactive < n || ngrouped > 0)
      {
        skip.resize(n);
        for (int z = 0; z < n; ++z)
          skip[z] = (!h->h_owned[z] || h->h_grouped[z]) ? 1 : 0;
      }
      job = device_greedy_colouring_start(h, skip.empty() ? nullptr : skip.data());
      CUDA_CHECK(cudaStreamSynchronize(h->stream));  // `skip` is read by an asynchronous copy
    }
    ncol = device_greedy_colouring_finish(h, *job, h->h_colour);
  }
  StageTimer ctm;
  ctm.t0 = t_enter;
  ctm.lap("  colouring: first-fit colours");
  h->ncolours = ncol;
  // Launch segments: (spatial chunk, colour).  One colour of the whole mesh streams all
  // inputs through HBM once per colour (ncu round 1: 3 x 0.68 GB read per step); with the
  // mesh cut into chunks whose working set fits the 126 MB L2 and the colours of a chunk
  // launched back to back, the 2nd and 3rd colour of a chunk hit L2.  Launches stay
  // serialised on one stream, so the summation order per DOF is fixed (deterministic).
  long chunk_cells = 1L << 40;  // default: one chunk (measured on B200: chunking only adds launch tails, the kernel is not DRAM bound)
  if (const char* e = getenv("EQLB_CHUNK_CELLS"))
    chunk_cells = std::max(1L, atol(e));
  int nchunk = (int)std::max<long>(1, ((long)h->ncell + chunk_cells - 1) / chunk_cells);
  if (h->flags & EQLB_FLAG_HOST_PIPELINE)
  {
    // host pipeline: about 128k cells per stage, 4..16 stages
    nchunk = (int)std::min<long>(16, std::max<long>(4, h->ncell / 262144));
    if (const char* e = getenv("EQLB_PIPE_STAGES"))
      nchunk = std::max(1, atoi(e));
    if (h->ncell < 4096)
      nchunk = 1;
  }
  // distributed runs: interface patches (a cell of the patch holds a vertex of another rank)
  // first, so that their halo exchange can overlap the interior patches (eqlb_set_part)
  h->interface_first = false;
  if ((h->flags & EQLB_FLAG_INTERFACE_FIRST) && !(h->flags & EQLB_FLAG_HOST_PIPELINE) && h->nactive < h->nnode)
  {
    h->interface_first = true;
    nchunk = 2;
  }
  h->nchunk = nchunk;
  static const bool host_order = getenv("EQLB_HOST_ORDER") && atoi(getenv("EQLB_HOST_ORDER")) != 0;
  if (job && !h->interface_first && nchunk <= 16 && !host_order)
  {
    order_on_device(h, *job, ncol, nchunk, ngrouped);
    ctm.lap("  colouring: launch order (device)");
    return;
  }
  if (!h->topo.node_cell_off)
    ensure_host_topology(h);
  const eqlb_handle::HostTopo& T = h->topo;
  // stage of a patch = chunk of its LAST cell: all its inputs are on the device once the
  // cell slabs 0..stage have arrived
  auto chunk_of = [&](int z)
  {
    if (h->interface_first)
    {
      for (int i = T.node_cell_off[z]; i < T.node_cell_off[z + 1]; ++i)
      {
        const int32_t* cn = &T.cell_node[3 * (size_t)T.node_cell[i]];
        if (!h->h_owned[cn[0]] || !h->h_owned[cn[1]] || !h->h_owned[cn[2]])
          return 0;
      }
      return 1;
    }
    int32_t cmax = 0;
    for (int i = T.node_cell_off[z]; i < T.node_cell_off[z + 1]; ++i)
      cmax = std::max(cmax, T.node_cell[i]);
    return (int)((long)cmax * nchunk / std::max(h->ncell, 1));
  };
  h->nseg = nchunk * ncol;
  h->h_colour_off.assign(h->nseg + 1, 0);
  // segment and lane class of every active patch, evaluated once (several host threads).
  // Lane class: patches eligible for the lane-per-cell kernels (interior patches, or any patch of a
  // single-RHS problem: `reversion_required` cannot occur there; at most 16 facets) by facet count
  // 4 / 8 / 16 -> 0 / 1 / 2, everything else 3.
  std::vector<int32_t> vseg(n, -1);
  std::vector<int8_t> vcls(n, 3);
  {
    const int nthr = n > (1 << 16) ? std::max(1, std::min(6, (int)std::thread::hardware_concurrency() / 2)) : 1;
    auto part = [&](int t)
    {
      const int z0 = (int)((long)n * t / nthr), z1 = (int)((long)n * (t + 1) / nthr);
      for (int z = z0; z < z1; ++z)
      {
        if (!h->h_owned[z] || h->h_grouped[z])
          continue;
        vseg[z] = chunk_of(z) * ncol + h->h_colour[z];
        const int nf = T.node_fct_off[z + 1] - T.node_fct_off[z], nc = T.node_cell_off[z + 1] - T.node_cell_off[z];
        vcls[z] = (nf > 16 || !(h->nrhs == 1 || nf == nc)) ? 3 : (nf <= 4 ? 0 : (nf <= 8 ? 1 : 2));
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthr; ++t)
      th.emplace_back(part, t);
    part(0);
    for (auto& x : th)
      x.join();
  }
  ctm.lap("  colouring: segment / lane class");
  // counting sort of the active patches by (segment, lane class), vertex order within a class; the histogram and
  // the scatter run on host threads over contiguous vertex ranges (same result for any thread count)
  const int nthr_o = n > (1 << 16) ? std::max(1, std::min(6, (int)std::thread::hardware_concurrency() / 2)) : 1;
  const size_t nkey = 4 * (size_t)h->nseg;
  std::vector<int32_t> hist((size_t)nthr_o * nkey, 0);  // [thread][segment][raw class 0..3]
  auto zrange = [&](int t, int& z0, int& z1)
  {
    z0 = (int)((long)n * t / nthr_o);
    z1 = (int)((long)n * (t + 1) / nthr_o);
  };
  auto run_threads = [&](auto&& fn)
  {
    std::vector<std::thread> th;
    for (int t = 1; t < nthr_o; ++t)
      th.emplace_back(fn, t);
    fn(0);
    for (auto& x : th)
      x.join();
  };
  run_threads(
      [&](int t)
      {
        int z0, z1;
        zrange(t, z0, z1);
        int32_t* hh = &hist[(size_t)t * nkey];
        for (int z = z0; z < z1; ++z)
          if (vseg[z] >= 0)
            hh[4 * vseg[z] + vcls[z]]++;
      });
  // class sizes per segment; a lane class with few members joins the next wider one (saves a launch)
  std::vector<int32_t> ccount(3 * (size_t)h->nseg, 0);
  for (int sg = 0; sg < h->nseg; ++sg)
  {
    int32_t tot = 0;
    for (int t = 0; t < nthr_o; ++t)
      for (int cl = 0; cl < 4; ++cl)
      {
        const int32_t v = hist[(size_t)t * nkey + 4 * sg + cl];
        tot += v;
        if (cl < 3)
          ccount[3 * sg + cl] += v;
      }
    h->h_colour_off[sg + 1] = tot;
  }
  h->h_colour_off[0] = ngrouped;  // grouped patches occupy the head of h_order
  for (int c = 0; c < h->nseg; ++c)
    h->h_colour_off[c + 1] += h->h_colour_off[c];
  h->h_order.resize(h->nactive);
  std::vector<int8_t> cmap(4 * (size_t)h->nseg);
  for (int sg = 0; sg < h->nseg; ++sg)
  {
    int32_t* cc = &ccount[3 * sg];
    int8_t* cm = &cmap[4 * sg];
    cm[0] = 0, cm[1] = 1, cm[2] = 2, cm[3] = 3;
    const int thr = std::max(8192, (cc[0] + cc[1] + cc[2]) / 16);
    if (cc[1] > 0 && cc[1] < thr && cc[2] > 0)
      cm[1] = 2, cc[2] += cc[1], cc[1] = 0;
    if (cc[0] > 0 && cc[0] < thr && cc[1] + cc[2] > 0)
    {
      cm[0] = cc[1] > 0 ? 1 : 2;
      cc[cm[0]] += cc[0], cc[0] = 0;
    }
  }
  h->h_colour_fast.assign(h->nseg, 0);
  h->h_seg_subs.assign(h->nseg, {});
  // start of (segment, merged class, thread): classes in order, threads in order within a class
  std::vector<int32_t> start((size_t)nthr_o * nkey, 0);
  for (int sg = 0; sg < h->nseg; ++sg)
  {
    int32_t at = h->h_colour_off[sg];
    for (int cl = 0; cl < 4; ++cl)
    {
      const int32_t first = at;
      for (int t = 0; t < nthr_o; ++t)
      {
        start[(size_t)t * nkey + 4 * sg + cl] = at;
        for (int raw = 0; raw < 4; ++raw)
          if (cmap[4 * sg + raw] == cl)
            at += hist[(size_t)t * nkey + 4 * sg + raw];
      }
      if (cl < 3 && at > first)
      {
        h->h_seg_subs[sg].push_back({first, at - first, 4 << cl, -1});
        h->h_colour_fast[sg] += at - first;
      }
    }
  }
  run_threads(
      [&](int t)
      {
        int z0, z1;
        zrange(t, z0, z1);
        int32_t* st = &start[(size_t)t * nkey];
        for (int z = z0; z < z1; ++z)
          if (vseg[z] >= 0)
          {
            const int sg = vseg[z];
            h->h_order[st[4 * sg + cmap[4 * sg + vcls[z]]]++] = z;
          }
      });

  ctm.lap("  colouring: launch order");
  // result ranges of the host pipeline: a range of DOFs can go back to the host after the
  // last stage with a patch that adds into it
  h->h_se_slabs.clear();
  h->h_ev_slabs.clear();
  if (nchunk > 1 && !h->interface_first)
  {
    std::vector<int> nstage(n, -1);
    for (int z = 0; z < n; ++z)
      if (h->h_owned[z] && !h->h_grouped[z])
        nstage[z] = vseg[z] / ncol;
    auto slab_lo = [&](long cnt, int sidx) { return (size_t)(cnt * sidx / nchunk); };
    const int kk = h->k;
    const size_t ncd = (size_t)(kk * kk - kk);
    // last stage that adds into the cells / facets of a slab (host threads: one slab each)
    std::vector<int> cfin(nchunk, 0), ffin(nchunk, 0);
    {
      const int nthr = n > (1 << 16) ? std::max(1, std::min(6, (int)std::thread::hardware_concurrency() / 2)) : 1;
      auto part = [&](int t)
      {
        for (int sidx = t; sidx < nchunk; sidx += nthr)
        {
          const size_t c0 = slab_lo(h->ncell, sidx), c1 = slab_lo(h->ncell, sidx + 1);
          int fin = 0;
          for (size_t c = c0; c < c1; ++c)
            for (int j = 0; j < 3; ++j)
              fin = std::max(fin, nstage[T.cell_node[3 * c + j]]);
          cfin[sidx] = fin;
          const size_t f0 = slab_lo(h->nfct, sidx), f1 = slab_lo(h->nfct, sidx + 1);
          int ff = 0;
          for (size_t f = f0; f < f1; ++f)
            ff = std::max(ff, std::max(nstage[T.fct_node[2 * f]], nstage[T.fct_node[2 * f + 1]]));
          ffin[sidx] = ff;
        }
      };
      std::vector<std::thread> th;
      for (int t = 1; t < nthr; ++t)
        th.emplace_back(part, t);
      part(0);
      for (auto& x : th)
        x.join();
    }
    for (int sidx = 0; sidx < nchunk; ++sidx)
    {
      const size_t c0 = slab_lo(h->ncell, sidx), c1 = slab_lo(h->ncell, sidx + 1);
      h->h_se_slabs.push_back({c0 * h->nrt, (c1 - c0) * h->nrt, cfin[sidx]});
      if (ncd)
        h->h_ev_slabs.push_back({(size_t)h->nfct * kk + c0 * ncd, (c1 - c0) * ncd, cfin[sidx]});
      const size_t f0 = slab_lo(h->nfct, sidx), f1 = slab_lo(h->nfct, sidx + 1);
      h->h_ev_slabs.push_back({f0 * kk, (f1 - f0) * kk, ffin[sidx]});
    }
    ctm.lap("  colouring: result slabs");
  }
  h->d_order.upload(h->h_order.data(), h->h_order.size());
  h->slabs_pending = false;
  h->ordered = true;
}

void append(std::vector<double>& dst, const double* src, size_t n, int& offset)
{
  offset = (int)dst.size();
  dst.insert(dst.end(), src, src + n);
  while (dst.size() % 2)
    dst.push_back(0.0);
}

} // namespace

eqlb_handle::~eqlb_handle()
{
  for (cudaEvent_t e : ev_in)
    cudaEventDestroy(e);
  for (cudaEvent_t e : ev_done)
    cudaEventDestroy(e);
  if (ev_start)
    cudaEventDestroy(ev_start);
  if (s_h2d)
    cudaStreamDestroy(s_h2d);
  if (s_d2h)
    cudaStreamDestroy(s_d2h);
}

// Host copies of the connectivity around the vertices (grouping of boundary patches, re-colouring after a BC set
// has grouped patches): fetched from the device on first use - most handles never need them (the caller's arrays
// are not referenced after eqlb_create).
void ensure_host_topology(eqlb_handle* h)
{
  if (!h->h_node_cell_off.empty())
    return;
  auto fetch = [&](std::vector<int32_t>& dst, const DevBuf<int32_t>& src)
  {
    dst.resize(src.n);
    eqlb_d2h(dst.data(), src.p, src.n * sizeof(int32_t));
  };
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  fetch(h->h_node_cell_off, h->d_node_cell_off);
  fetch(h->h_node_cell, h->d_node_cell);
  fetch(h->h_cell_node, h->d_cell_node);
  fetch(h->h_node_fct_off, h->d_node_fct_off);
  fetch(h->h_node_fct, h->d_node_fct);
  fetch(h->h_fct_node, h->d_fct_node);
  h->topo = {h->h_node_cell_off.data(), h->h_node_cell.data(), h->h_cell_node.data(),
             h->h_node_fct_off.data(), h->h_node_fct.data(), h->h_fct_node.data()};
}

// Shared tail of eqlb_set_bcs / eqlb_set_bcs_poly: grouping of 2-cell traction patches (host, needs the
// facet types and node markers on the host), colouring, record buffers, device patch builder.
static void finish_bcs(eqlb_handle* h, const int8_t* facet_type, const int8_t* node_on_stress_bnd, bool upload_node_markers,
                       StageTimer& tm)
{
    if (node_on_stress_bnd)
    {
      if (upload_node_markers)
        h->d_node_on_bnd.upload(node_on_stress_bnd, h->nnode);
      // grouped boundary patches (se/reconstruction.hpp:170-234, k == 2 only): 2-cell
      // patches on a pure traction boundary are solved together with the adjacent
      // patch, which then imposes weak symmetry on the accumulated global stress
      h->h_grouped.assign(h->nnode, 0);
      h->h_group_off.clear();
      std::vector<int32_t> gorder;
      bool any_marked = false;
      for (int z = 0; z < h->nnode && !any_marked; ++z)
        any_marked = node_on_stress_bnd[z] != 0;
      if ((h->flags & EQLB_FLAG_STRESS) && h->k == 2 && any_marked)
      {
        ensure_host_topology(h);
        auto ncells_of = [&](int z) { return h->h_node_cell_off[z + 1] - h->h_node_cell_off[z]; };
        std::vector<uint8_t> perform(h->nnode, 1);
        for (int z = 0; z < h->nnode; ++z)
        {
          if (!(node_on_stress_bnd[z] && perform[z] && h->h_owned[z]) || ncells_of(z) != 2)
            continue;
          // adjacent_internal_patch (se/Patch.cpp:761-784)
          int inner = -1;
          for (int i = h->h_node_fct_off[z]; i < h->h_node_fct_off[z + 1]; ++i)
          {
            const int32_t fct = h->h_node_fct[i];
            if (facet_type[fct] == EQLB_FCT_INTERNAL)
            {
              inner = (h->h_fct_node[2 * fct] == z) ? h->h_fct_node[2 * fct + 1] : h->h_fct_node[2 * fct];
              break;
            }
          }
          // group_boundary_patches (se/Patch.cpp:60-104)
          std::vector<int32_t> grouped{inner};
          for (int i = h->h_node_cell_off[inner]; i < h->h_node_cell_off[inner + 1]; ++i)
            for (int j = 0; j < 3; ++j)
            {
              const int32_t pnt = h->h_cell_node[3 * (size_t)h->h_node_cell[i] + j];
              if (node_on_stress_bnd[pnt] && std::find(grouped.begin(), grouped.end(), pnt) == grouped.end()
                  && ncells_of(pnt) == 2)
                grouped.push_back(pnt);
            }
          if (grouped.size() < 2)
            continue;
          if (h->h_group_off.empty())
            h->h_group_off.push_back(0);
          for (int32_t q : grouped)
          {
            if (!perform[q] || !h->h_owned[q])
              throw EqlbError(EQLB_ERR_INPUT,
                              "Incompatible mesh! To many patches with 2 cells on neumann boundary.");
            perform[q] = 0;
            h->h_grouped[q] = 1;
            gorder.push_back(q);
          }
          h->h_group_off.push_back((int32_t)gorder.size());
        }
      }
      if (!gorder.empty() || !h->ordered || h->coloured_with_groups)
        colour_patches(h);
      h->coloured_with_groups = !gorder.empty();
      if (!gorder.empty())
        CUDA_CHECK(cudaMemcpy(h->d_order.p, gorder.data(), gorder.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    else
    {
      // the colouring of eqlb_create is still valid unless a previous BC set grouped patches
      h->h_grouped.assign(h->nnode, 0);
      h->h_group_off.clear();
      if (h->coloured_with_groups || !h->ordered)
        colour_patches(h);
      h->coloured_with_groups = false;
    }

    tm.lap("set_bcs: grouping + colouring");
    // patch records (colour-sorted)
    h->pstride = ((size_t)h->nactive + 31) / 32 * 32;
    if (h->pstride == 0)
      h->pstride = 32;
    h->d_pnode.alloc(h->pstride);
    h->d_pncells.alloc(h->pstride);
    h->d_pcell.alloc(h->pstride * h->ncmax);
    h->d_pinfo.alloc(h->pstride * h->ncmax);
    h->d_prhs.alloc(h->pstride * h->nrhs);
    h->d_pcell.zero(h->stream);
    h->d_pinfo.zero(h->stream);
    {
      // lane records: S lanes per patch for the eligible head of every segment, each
      // segment padded to a whole number of warps (zero records: ncells = 0)
      std::vector<int64_t> seginfo;
      int64_t nrec = 0;
      for (int sg = 0; sg < h->nseg; ++sg)
        for (auto& sub : h->h_seg_subs[sg])
        {
          sub.recoff = nrec;
          seginfo.insert(seginfo.end(), {(int64_t)sub.first, (int64_t)sub.count, (int64_t)sub.lanes, nrec});
          nrec += ((int64_t)sub.count * sub.lanes + 127) / 128 * 128;
        }
      h->nsub = (int)(seginfo.size() / 4);
      if (seginfo.empty())
        seginfo.assign(4, 0);
      h->d_prec.alloc((size_t)std::max<int64_t>(nrec, 128));
      h->d_prec.zero(h->stream);
      h->d_seginfo.upload(seginfo.data(), seginfo.size());
    }
    tm.lap("set_bcs: record buffers");
    launch_patch_builder(h, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    tm.lap("set_bcs: patch builder");
    h->bcs_set = true;
}

extern "C"
{

const char* eqlb_last_error(void) { return g_last_error.c_str(); }
const char* eqlb_version(void) { return "eqlb_b200 0.1 (sm_100a)"; }

int eqlb_create(const eqlb_mesh* mesh, const eqlb_tables* t, int nrhs, uint32_t flags, eqlb_handle** out)
{
  return guarded(
      [&]
      {
        if (!mesh || !t || !out)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_create: null argument");
        *out = nullptr;
        // input checks of se/reconstruction.hpp:358-388
        if (nrhs < 1 || nrhs > EQLB_MAXRHS)
          throw EqlbError(EQLB_ERR_INPUT, "Equilibration: Input sizes does not match");
        if (t->p > t->k - 1)
          throw EqlbError(EQLB_ERR_INPUT, "Equilibration: Wrong polynomial degree of the projected RHS");
        if ((flags & EQLB_FLAG_STRESS) && nrhs < 2)
          throw EqlbError(EQLB_ERR_INPUT, "Stress equilibration: Specify all rows of stress tensor");
        if ((flags & EQLB_FLAG_STRESS) && t->k < 2)
          throw EqlbError(EQLB_ERR_INPUT, "Stress equilibration: RT_k with k>1 required!");
        // patch sizes (se/Patch.cpp:337-404)
        int ncmax = 0;
        for (int i = 0; i < mesh->nnode; ++i)
        {
          if (mesh->node_owned && !mesh->node_owned[i])
            continue;
          const int nc = mesh->node_cell_off[i + 1] - mesh->node_cell_off[i];
          if (nc == 1)
            throw EqlbError(EQLB_ERR_INPUT,
                            "Patch around node " + std::to_string(i) + " has only 1 cells.");
          ncmax = std::max(ncmax, nc);
        }
        if (ncmax > EQLB_NCMAX)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_b200: patches with more than " + std::to_string(EQLB_NCMAX)
                                              + " cells are not supported");
        StageTimer tm;
        require_device();
        tm.lap("create: device check");

        std::unique_ptr<eqlb_handle> h(new eqlb_handle());
        CUDA_CHECK(cudaGetDevice(&h->device));
        h->flags = flags;
        h->nrhs = nrhs;
        h->nnode = mesh->nnode;
        h->ncell = mesh->ncell;
        h->nfct = mesh->nfct;
        h->k = t->k;
        h->p = t->p;
        h->nrt = t->nrt;
        h->ndg = t->ndg;
        h->ndg_fct = t->ndg_fct;
        h->ndiv = t->ndiv;
        h->nadd = t->nadd;
        h->nq = t->nq;
        h->nqf = t->nqf;
        h->ncmax = ncmax;
        h->h_owned.assign(mesh->nnode, 1);
        if (mesh->node_owned)
          h->h_owned.assign(mesh->node_owned, mesh->node_owned + mesh->nnode);
        h->nactive = 0;
        for (uint8_t o : h->h_owned)
          h->nactive += o ? 1 : 0;

        const size_t nn = mesh->nnode, nc = mesh->ncell, nf = mesh->nfct;
        // host work that does not touch the device runs beside the uploads: the greedy colouring (reads the
        // caller's topology arrays) and, for stress handles, the topology copies eqlb_set_bcs needs later
        // (boundary-patch grouping, recolouring).  Measured at 1024^2: upload 64-180 ms, colouring 55 ms, copies 72 ms.
        h->h_grouped.assign(nn, 0);
        h->topo = {mesh->node_cell_off, mesh->node_cell, mesh->cell_node, mesh->node_fct_off, mesh->node_fct, mesh->fct_node};
        std::exception_ptr copy_err, dgmap_err;
        eqlb_handle* hp = h.get();
        std::thread copy_thread([] {});  // (topology copies of stress handles are fetched lazily, ensure_host_topology)
        // DG dofmap: identity layout (cell*ndg + i) is the DOLFINx layout; otherwise indirect
        bool dg_identity = true;
        const int ndg_chk = t->ndg;
        std::thread dgmap_thread(
            [mesh, nc, ndg_chk, &dg_identity, &dgmap_err]
            {
              try
              {
                if (mesh->dg_dofmap)
                  for (size_t i = 0; i < nc * (size_t)ndg_chk; ++i)
                    if (mesh->dg_dofmap[i] != (int32_t)i)
                    {
                      dg_identity = false;
                      break;
                    }
              }
              catch (...)
              {
                dgmap_err = std::current_exception();
              }
            });
        struct Joiner
        {
          std::thread &a, &b;
          ~Joiner()
          {
            if (a.joinable())
              a.join();
            if (b.joinable())
              b.join();
          }
        } joiner{copy_thread, dgmap_thread};
        h->d_x.alloc(nn * 3);
        h->d_cell_node.alloc(nc * 3);
        h->d_cell_fct.alloc(nc * 3);
        h->d_fct_node.alloc(nf * 2);
        h->d_fct_cell_off.alloc(nf + 1);
        h->d_fct_cell.alloc(mesh->fct_cell_off[nf]);
        h->d_node_cell_off.alloc(nn + 1);
        h->d_node_cell.alloc(mesh->node_cell_off[nn]);
        h->d_node_fct_off.alloc(nn + 1);
        h->d_node_fct.alloc(mesh->node_fct_off[nn]);
        h->d_fct_perms.alloc(nc * 3);
        tm.lap("create: mesh allocations");
        // connectivity of the colouring first: its kernel runs while the other arrays are uploaded
        h->d_cell_node.upload(mesh->cell_node, nc * 3);
        h->d_node_cell_off.upload(mesh->node_cell_off, nn + 1);
        h->d_node_cell.upload(mesh->node_cell, mesh->node_cell_off[nn]);
        h->d_node_fct_off.upload(mesh->node_fct_off, nn + 1);  // (lane classes of the device launch order)
        static const bool host_col = getenv("EQLB_HOST_COLOURING") && atoi(getenv("EQLB_HOST_COLOURING")) != 0;
        std::vector<uint8_t> skip;
        if (!host_col)
        {
          if (h->nactive < (int)nn)
          {
            skip.resize(nn);
            for (size_t z = 0; z < nn; ++z)
              skip[z] = h->h_owned[z] ? 0 : 1;
          }
          h->colouring_job = device_greedy_colouring_start(h.get(), skip.empty() ? nullptr : skip.data());
        }
        // the other arrays travel on a second host thread while this one waits for the colours and sorts the
        // patches into launch order (the buffers are allocated, the copies go through the staging pool)
        std::exception_ptr upload_err;
        const int dev_id = h->device;
        std::thread upload_thread(
            [hp, mesh, nn, nc, nf, dev_id, &upload_err]
            {
              try
              {
                CUDA_CHECK(cudaSetDevice(dev_id));
                hp->d_x.upload(mesh->x, nn * 3);
                hp->d_cell_fct.upload(mesh->cell_fct, nc * 3);
                hp->d_fct_node.upload(mesh->fct_node, nf * 2);
                hp->d_fct_cell_off.upload(mesh->fct_cell_off, nf + 1);
                hp->d_fct_cell.upload(mesh->fct_cell, mesh->fct_cell_off[nf]);
                hp->d_node_fct.upload(mesh->node_fct, mesh->node_fct_off[nn]);
                hp->d_fct_perms.upload(mesh->fct_perms, nc * 3);
              }
              catch (...)
              {
                upload_err = std::current_exception();
              }
            });
        struct Joiner1
        {
          std::thread& a;
          ~Joiner1()
          {
            if (a.joinable())
              a.join();
          }
        } joiner_up{upload_thread};
        tm.lap("create: colouring connectivity upload");
        // colouring (device) + launch order (host threads) while the topology copies / dofmap check run
        colour_patches(h.get());
        tm.lap("create: colouring + launch order");
        upload_thread.join();
        if (upload_err)
          std::rethrow_exception(upload_err);
        tm.lap("create: wait for the mesh upload");
        dgmap_thread.join();
        if (dgmap_err)
          std::rethrow_exception(dgmap_err);
        h->dg_identity = dg_identity;
        if (!h->dg_identity)
          h->d_dg_dofmap.upload(mesh->dg_dofmap, nc * t->ndg);
        tm.lap("create: dg dofmap check");
        // reference-matrix tables -> one flat device block
        std::vector<double> flat;
        TableView& tv = h->tv;
        tv.k = t->k;
        tv.p = t->p;
        tv.nrt = t->nrt;
        tv.ndg = t->ndg;
        tv.ndg_fct = t->ndg_fct;
        tv.ndiv = t->ndiv;
        tv.nadd = t->nadd;
        const int k = t->k, nrt = t->nrt, ndg = t->ndg, nt = 1 + t->ndiv;
        append(flat, t->rt_mass, (size_t)3 * nrt * nrt, tv.o_rt_mass);
        append(flat, t->fct_mom, (size_t)9 * k * ndg, tv.o_fct_mom);
        append(flat, t->cell_mom_f, (size_t)3 * nt * ndg, tv.o_cell_mom_f);
        append(flat, t->cell_mom_g, (size_t)3 * nt * ndg * 2, tv.o_cell_mom_g);
        append(flat, t->bc_mat, (size_t)9 * k * k, tv.o_bc_mat);
        append(flat, t->trafo, (size_t)k * k, tv.o_trafo);
        append(flat, t->rt_p1, (size_t)nrt * 6, tv.o_rt_p1);
        append(flat, t->dg_mono, (size_t)nt * ndg, tv.o_dg_mono);
        append(flat, t->hat_dg_rt, (size_t)3 * ndg * nrt * 2, tv.o_hat_dg_rt);
        append(flat, t->mono_int, (size_t)nt, tv.o_mono_int);
        h->d_tables.upload(flat.data(), flat.size());
        if (t->rt_basix_fct && t->rt_basix_int)
        {
          std::vector<double> bx(t->rt_basix_fct, t->rt_basix_fct + (size_t)k * k);
          bx.insert(bx.end(), t->rt_basix_int, t->rt_basix_int + (size_t)(k * k - k) * nrt);
          bx.push_back(0.0);
          h->d_basix.upload(bx.data(), bx.size());
        }
        h->npk = t->npk;
        if (t->pk_q && t->pk_gq && t->pk_hq && t->rt_div_q && t->pk_grad_dg && t->pk_to_dg && t->npk > 0)
        {
          // [qwts | rt_q | rt_div_q | pk_q | pk_gq | pk_hq | dg_q | pk_grad_dg | pk_to_dg]
          std::vector<double> pr;
          const int nq = t->nq, npk = t->npk;
          auto put = [&](int slot, const double* src, size_t n)
          {
            h->o_pr[slot] = (int)pr.size();
            pr.insert(pr.end(), src, src + n);
          };
          put(0, t->qwts, nq);
          put(1, t->rt_q, (size_t)nq * nrt * 2);
          put(2, t->rt_div_q, (size_t)nq * nrt);
          put(3, t->pk_q, (size_t)nq * npk);
          put(4, t->pk_gq, (size_t)nq * npk * 2);
          put(5, t->pk_hq, (size_t)nq * npk * 3);
          put(6, t->dg_q, (size_t)3 * nq * ndg);
          put(7, t->pk_grad_dg, (size_t)ndg * npk * 2);
          put(8, t->pk_to_dg, (size_t)ndg * npk);
          h->d_primal.upload(pr.data(), pr.size());
        }
        tv.data = h->d_tables.p;
        tv.ndoubles = (int)flat.size();

        // cell-wise L2 projector into DG_p: P = M^-1 Phi^T W  (ndg x nq), reference cell
        {
          const int nq = t->nq;
          std::vector<double> Mm((size_t)ndg * ndg, 0.0), P((size_t)ndg * nq, 0.0);
          for (int q = 0; q < nq; ++q)
            for (int i = 0; i < ndg; ++i)
              for (int j = 0; j < ndg; ++j)
                Mm[i * ndg + j] += t->qwts[q] * t->dg_q[(size_t)q * ndg + i] * t->dg_q[(size_t)q * ndg + j];
          // Cholesky of the reference mass matrix
          std::vector<double> Lc(Mm);
          for (int j = 0; j < ndg; ++j)
          {
            double d = Lc[j * ndg + j];
            for (int q = 0; q < j; ++q)
              d -= Lc[j * ndg + q] * Lc[j * ndg + q];
            d = std::sqrt(d);
            Lc[j * ndg + j] = d;
            for (int i = j + 1; i < ndg; ++i)
            {
              double s = Lc[i * ndg + j];
              for (int q = 0; q < j; ++q)
                s -= Lc[i * ndg + q] * Lc[j * ndg + q];
              Lc[i * ndg + j] = s / d;
            }
          }
          for (int q = 0; q < nq; ++q)
          {
            std::vector<double> b(ndg);
            for (int i = 0; i < ndg; ++i)
              b[i] = t->qwts[q] * t->dg_q[(size_t)q * ndg + i];
            for (int i = 0; i < ndg; ++i)
            {
              double s = b[i];
              for (int c = 0; c < i; ++c)
                s -= Lc[i * ndg + c] * b[c];
              b[i] = s / Lc[i * ndg + i];
            }
            for (int i = ndg - 1; i >= 0; --i)
            {
              double s = b[i];
              for (int c = i + 1; c < ndg; ++c)
                s -= Lc[c * ndg + i] * b[c];
              b[i] = s / Lc[i * ndg + i];
            }
            for (int i = 0; i < ndg; ++i)
              P[(size_t)i * nq + q] = b[i];
          }
          h->d_proj.upload(P.data(), P.size());
        }

        if (t->k == 1 && t->p == 0)
          build_k1_tables(h.get(), t);
        if (t->k == 2 && t->p == 1)
          build_k2_tables(h.get(), t);
        if (kw_supported(t->k, t->ndg))
          build_kw_tables(h.get(), t);
        tm.lap("create: tables");
        launch_compute_cellJ(h.get());
        tm.lap("create: cell Jacobians");
        copy_thread.join();
        if (copy_err)
          std::rethrow_exception(copy_err);
        // the caller's arrays are not referenced after eqlb_create
        h->topo = {};
        tm.lap("create: wait for topology copies");
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        *out = h.release();
      });
}

void eqlb_destroy(eqlb_handle* h) { delete h; }


int eqlb_set_stream(eqlb_handle* h, void* cuda_stream)
{
  return guarded(
      [&]
      {
        if (!h)
          throw EqlbError(EQLB_ERR_INPUT, "null handle");
        h->stream = (cudaStream_t)cuda_stream;
      });
}

int eqlb_set_bcs(eqlb_handle* h, const int8_t* facet_type, const double* const* bflux, const int8_t* local_fct_id,
                 const int8_t* node_on_stress_bnd)
{
  return guarded(
      [&]
      {
        if (!h || !facet_type)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_set_bcs: null argument");
        // queued patch kernels (EQLB_DEVICE calls return without synchronising) may still read the boundary data
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        StageTimer tm;
        const size_t nf = h->nfct;
        // every boundary facet has to be classified for every RHS (se/Patch.cpp:464-470)
        h->d_facet_type.upload(facet_type, (size_t)h->nrhs * nf);
        const size_t nb = (size_t)h->ncell * h->nrt;
        // the boundary function is only read on cells with a flux-BC facet: without any such facet no
        // [nrhs][ncell*nrt] array is allocated (268 MB + zero fill per RHS at 1024^2, degree 2)
        bool any_dual = false;
        for (size_t i = 0; i < (size_t)h->nrhs * nf && !any_dual; ++i)
          any_dual = facet_type[i] == EQLB_FCT_ESSNT_DUAL;
        h->bflux_dummy = !any_dual;
        h->d_bflux.alloc(any_dual ? (size_t)h->nrhs * nb : 2);
        h->d_bflux.zero(h->stream);
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        for (int r = 0; r < h->nrhs && any_dual; ++r)
          if (bflux && bflux[r])
            CUDA_CHECK(cudaMemcpy(h->d_bflux.p + (size_t)r * nb, bflux[r], nb * sizeof(double), cudaMemcpyHostToDevice));
        tm.lap("set_bcs: facet types + bflux");
        // local_fct_id (`base::BoundaryData::_local_fct_id`): the hot path derives the cell-local id of a boundary
        // facet from the patch maps; a caller-supplied array is validated against the mesh
        h->d_local_fct_id.alloc(nf);
        if (local_fct_id)
        {
          CUDA_CHECK(cudaMemcpy(h->d_local_fct_id.p, local_fct_id, nf, cudaMemcpyHostToDevice));
          if (count_bad_local_fct_ids(h, h->d_local_fct_id.p) > 0)
            throw EqlbError(EQLB_ERR_INPUT, "eqlb_set_bcs: local_fct_id does not match the cell-local id of a flux-BC facet");
        }
        finish_bcs(h, facet_type, node_on_stress_bnd, true, tm);
      });
}

int eqlb_set_bcs_poly(eqlb_handle* h, const int32_t* nprime, const int32_t* const* prime_facets, const int32_t* nbc,
                      const eqlb_fluxbc* const* bcs)
{
  return guarded(
      [&]
      {
        if (!h || !nprime || !prime_facets || !nbc)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_set_bcs_poly: null argument");
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        StageTimer tm;
        const size_t nf = h->nfct, nb = (size_t)h->ncell * h->nrt;
        const bool stress = (h->flags & EQLB_FLAG_STRESS) != 0;
        h->d_facet_type.alloc((size_t)h->nrhs * nf);
        h->d_facet_type.zero(h->stream);
        int64_t nbcf = 0;
        for (int r = 0; r < h->nrhs; ++r)
          for (int b = 0; b < nbc[r]; ++b)
            nbcf += bcs[r][b].nfct;
        h->bflux_dummy = nbcf == 0;
        h->d_bflux.alloc(nbcf ? (size_t)h->nrhs * nb : 2);
        h->d_bflux.zero(h->stream);
        h->d_local_fct_id.alloc(nf);
        h->d_local_fct_id.zero(h->stream);
        DevBuf<int32_t> d_cnt;
        if (stress)
        {
          d_cnt.alloc(h->nnode);
          d_cnt.zero(h->stream);
          h->d_node_on_bnd.alloc(h->nnode);
        }
        std::vector<std::unique_ptr<DevBuf<int32_t>>> keep_i;
        std::vector<std::unique_ptr<DevBuf<double>>> keep_d;
        for (int r = 0; r < h->nrhs; ++r)
        {
          auto dp = std::make_unique<DevBuf<int32_t>>();
          if (nprime[r] > 0)
          {
            if (!prime_facets[r])
              throw EqlbError(EQLB_ERR_INPUT, "eqlb_set_bcs_poly: null facet list");
            dp->upload(prime_facets[r], nprime[r]);
          }
          launch_bc_poly(h, r, nprime[r], dp->p, 0, nullptr, 0, nullptr, nullptr, nullptr);
          keep_i.push_back(std::move(dp));
          for (int b = 0; b < nbc[r]; ++b)
          {
            const eqlb_fluxbc& bc = bcs[r][b];
            if (bc.nfct <= 0)
              continue;
            if (!bc.facets || !bc.coeffs || bc.ncoef < 1)
              throw EqlbError(EQLB_ERR_INPUT, "eqlb_set_bcs_poly: FluxBC without facets or coefficients");
            auto df = std::make_unique<DevBuf<int32_t>>();
            auto dc = std::make_unique<DevBuf<double>>();
            df->upload(bc.facets, bc.nfct);
            dc->upload(bc.coeffs, (size_t)bc.nfct * bc.ncoef);
            // (`base/BoundaryData.cpp:611-619`) nodes of traction facets of the first gdim rows are counted
            launch_bc_poly(h, r, 0, nullptr, bc.nfct, df->p, bc.ncoef, dc->p, h->d_local_fct_id.p, (stress && r < 2) ? d_cnt.p : nullptr);
            keep_i.push_back(std::move(df));
            keep_d.push_back(std::move(dc));
          }
        }
        std::vector<int8_t> ft((size_t)h->nrhs * nf), nob;
        if (stress)
        {
          launch_bc_node_markers(h, d_cnt.p);
          nob.resize(h->nnode);
          CUDA_CHECK(cudaMemcpyAsync(nob.data(), h->d_node_on_bnd.p, h->nnode, cudaMemcpyDeviceToHost, h->stream));
        }
        // the grouping of 2-cell traction patches and the patch classification bookkeeping are host side
        CUDA_CHECK(cudaMemcpyAsync(ft.data(), h->d_facet_type.p, ft.size(), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        tm.lap("set_bcs_poly: boundary data on device");
        finish_bcs(h, ft.data(), stress ? nob.data() : nullptr, false, tm);
      });
}

int eqlb_get_boundary_data(eqlb_handle* h, int8_t* facet_type, double* const* bflux, int8_t* local_fct_id, int8_t* node_on_stress_bnd)
{
  return guarded(
      [&]
      {
        if (!h || !h->bcs_set)
          throw EqlbError(EQLB_ERR_STATE, "eqlb_get_boundary_data: boundary conditions not set");
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        const size_t nf = h->nfct, nb = (size_t)h->ncell * h->nrt;
        if (facet_type)
          CUDA_CHECK(cudaMemcpy(facet_type, h->d_facet_type.p, (size_t)h->nrhs * nf, cudaMemcpyDeviceToHost));
        if (bflux)
          for (int r = 0; r < h->nrhs; ++r)
            if (bflux[r])
            {
              if (h->bflux_dummy)
                std::memset(bflux[r], 0, nb * sizeof(double));
              else
                CUDA_CHECK(cudaMemcpy(bflux[r], h->d_bflux.p + (size_t)r * nb, nb * sizeof(double), cudaMemcpyDeviceToHost));
            }
        if (local_fct_id && h->d_local_fct_id.p)
          CUDA_CHECK(cudaMemcpy(local_fct_id, h->d_local_fct_id.p, nf, cudaMemcpyDeviceToHost));
        if (node_on_stress_bnd && h->d_node_on_bnd.p)
          CUDA_CHECK(cudaMemcpy(node_on_stress_bnd, h->d_node_on_bnd.p, h->nnode, cudaMemcpyDeviceToHost));
      });
}

int eqlb_patch_dims(eqlb_handle* h, int32_t* npatch, int32_t* ncmax, int32_t* ncolours)
{
  return guarded(
      [&]
      {
        if (!h)
          throw EqlbError(EQLB_ERR_INPUT, "null handle");
        if (npatch)
          *npatch = h->nnode;
        if (ncmax)
          *ncmax = h->ncmax;
        if (ncolours)
          *ncolours = h->ncolours;
      });
}

int eqlb_get_patch_maps(eqlb_handle* h, int32_t* ncells, int32_t* cells, int32_t* fcts, int8_t* inodes_local,
                        int8_t* fcts_local, int8_t* type, uint8_t* reversed, uint8_t* reversion, int32_t* colour)
{
  return guarded(
      [&]
      {
        if (!h || !h->bcs_set)
          throw EqlbError(EQLB_ERR_STATE, "eqlb_get_patch_maps: call eqlb_set_bcs first");
        const size_t np = h->nnode, w = h->ncmax + 2;
        DevBuf<int32_t> d_nc, d_cells, d_fcts;
        DevBuf<int8_t> d_inod, d_fl, d_type;
        DevBuf<uint8_t> d_rev, d_reversion;
        d_nc.alloc(np);
        d_cells.alloc(np * w);
        d_fcts.alloc(np * w);
        d_inod.alloc(np * w);
        d_fl.alloc(np * 2 * (h->ncmax + 1));
        d_type.alloc(np * h->nrhs);
        d_rev.alloc(np * h->ncmax * 2);
        d_reversion.alloc(np * h->nrhs);
        CUDA_CHECK(cudaMemsetAsync(d_cells.p, 0xFF, np * w * 4, h->stream));
        CUDA_CHECK(cudaMemsetAsync(d_fcts.p, 0xFF, np * w * 4, h->stream));
        CUDA_CHECK(cudaMemsetAsync(d_inod.p, 0xFF, np * w, h->stream));
        CUDA_CHECK(cudaMemsetAsync(d_fl.p, 0xFF, np * 2 * (h->ncmax + 1), h->stream));
        CUDA_CHECK(cudaMemsetAsync(d_rev.p, 0xFF, np * h->ncmax * 2, h->stream));
        launch_patch_builder(h, d_nc.p, d_cells.p, d_fcts.p, d_inod.p, d_fl.p, d_type.p, d_rev.p, d_reversion.p);
        auto fetch = [&](void* dst, const void* src, size_t bytes)
        {
          if (dst)
            CUDA_CHECK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
        };
        fetch(ncells, d_nc.p, np * 4);
        fetch(cells, d_cells.p, np * w * 4);
        fetch(fcts, d_fcts.p, np * w * 4);
        fetch(inodes_local, d_inod.p, np * w);
        fetch(fcts_local, d_fl.p, np * 2 * (h->ncmax + 1));
        fetch(type, d_type.p, np * h->nrhs);
        fetch(reversed, d_rev.p, np * h->ncmax * 2);
        fetch(reversion, d_reversion.p, np * h->nrhs);
        if (colour)
          std::memcpy(colour, h->h_colour.data(), np * 4);
      });
}

int eqlb_get_se_dofmaps(eqlb_handle* h, int32_t* dofmap, int32_t* projflux_fct, int8_t* bmarkers, int32_t* ndpc_out,
                        int32_t* hzmax_out)
{
  return guarded(
      [&]
      {
        if (!h || !h->bcs_set)
          throw EqlbError(EQLB_ERR_STATE, "eqlb_get_se_dofmaps: call eqlb_set_bcs first");
        const bool stress = (h->flags & EQLB_FLAG_STRESS) != 0;
        const int ndpc = 2 * h->k + h->nadd + h->ndiv + (stress ? 3 : 0);
        const int hzmax = 1 + (h->k - 1) * (h->ncmax + 1) + h->nadd * h->ncmax;
        if (ndpc_out)
          *ndpc_out = ndpc;
        if (hzmax_out)
          *hzmax_out = hzmax;
        const size_t np = h->nnode;
        DevBuf<int32_t> d_dm, d_pf;
        DevBuf<int8_t> d_bm;
        const size_t n_dm = np * 4 * (h->ncmax + 2) * ndpc, n_pf = np * (h->ncmax + 1) * 2 * h->ndg_fct,
                     n_bm = np * h->nrhs * hzmax;
        if (dofmap)
        {
          d_dm.alloc(n_dm);
          CUDA_CHECK(cudaMemsetAsync(d_dm.p, 0xFF, n_dm * 4, h->stream));
        }
        if (projflux_fct)
        {
          d_pf.alloc(n_pf);
          CUDA_CHECK(cudaMemsetAsync(d_pf.p, 0xFF, n_pf * 4, h->stream));
        }
        if (bmarkers)
        {
          d_bm.alloc(n_bm);
          CUDA_CHECK(cudaMemsetAsync(d_bm.p, 0xFF, n_bm, h->stream));
        }
        launch_se_dofmaps(h, d_dm.p, d_pf.p, d_bm.p, ndpc, hzmax);
        if (dofmap)
          CUDA_CHECK(cudaMemcpy(dofmap, d_dm.p, n_dm * 4, cudaMemcpyDeviceToHost));
        if (projflux_fct)
          CUDA_CHECK(cudaMemcpy(projflux_fct, d_pf.p, n_pf * 4, cudaMemcpyDeviceToHost));
        if (bmarkers)
          CUDA_CHECK(cudaMemcpy(bmarkers, d_bm.p, n_bm, cudaMemcpyDeviceToHost));
      });
}

int eqlb_get_ev_dofmaps(eqlb_handle* h, int32_t* ncells, int32_t* cells, int32_t* fcts, int8_t* inodes_local, int32_t* dofs_elmt,
                        int32_t* dofs_patch, int32_t* dofs_global, int32_t* list_patch, int32_t* list_global)
{
  return guarded(
      [&]
      {
        if (!h || !h->bcs_set)
          throw EqlbError(EQLB_ERR_STATE, "eqlb_get_ev_dofmaps: call eqlb_set_bcs first");
        if (h->ncmax > EQLB_NCMAX)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_get_ev_dofmaps: more than 16 cells around a vertex");
        const size_t np = h->nnode, ncm = h->ncmax, nz = h->nrt + h->ndg - h->k;
        const size_t lenf = ncm * (h->nrt - 3 * h->k) + (ncm + 1) * h->k;
        DevBuf<int32_t> b_nc, b_c, b_f, b_e, b_p, b_g, b_lp, b_lg;
        DevBuf<int8_t> b_i;
        auto prep = [&](auto& buf, size_t n)
        {
          buf.alloc(n);
          CUDA_CHECK(cudaMemsetAsync(buf.p, 0xFF, n * sizeof(*buf.p), h->stream));
        };
        prep(b_nc, np);
        prep(b_c, np * ncm);
        prep(b_f, np * (ncm + 1));
        prep(b_i, np * ncm);
        prep(b_e, np * ncm * nz);
        prep(b_p, np * ncm * nz);
        prep(b_g, np * ncm * nz);
        prep(b_lp, np * lenf);
        prep(b_lg, np * lenf);
        launch_ev_dofmaps(h, b_nc.p, b_c.p, b_f.p, b_i.p, b_e.p, b_p.p, b_g.p, b_lp.p, b_lg.p);
        auto fetch = [&](auto* dst, auto& buf)
        {
          if (dst)
            CUDA_CHECK(cudaMemcpy(dst, buf.p, buf.n * sizeof(*buf.p), cudaMemcpyDeviceToHost));
        };
        fetch(ncells, b_nc);
        fetch(cells, b_c);
        fetch(fcts, b_f);
        fetch(inodes_local, b_i);
        fetch(dofs_elmt, b_e);
        fetch(dofs_patch, b_p);
        fetch(dofs_global, b_g);
        fetch(list_patch, b_lp);
        fetch(list_global, b_lg);
      });
}

int eqlb_pin_host(void* ptr, size_t bytes)
{
  return guarded(
      [&]
      {
        if (!ptr || bytes == 0)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_pin_host: null buffer");
        require_device();
        cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
        if (e == cudaErrorHostMemoryAlreadyRegistered)
        {
          cudaGetLastError();
          return;
        }
        CUDA_CHECK(e);
      });
}

int eqlb_unpin_host(void* ptr)
{
  return guarded(
      [&]
      {
        if (!ptr)
          return;
        cudaError_t e = cudaHostUnregister(ptr);
        if (e == cudaErrorHostMemoryNotRegistered)
        {
          cudaGetLastError();
          return;
        }
        CUDA_CHECK(e);
      });
}

int eqlb_set_part(eqlb_handle* h, int part)
{
  return guarded(
      [&]
      {
        if (!h)
          throw EqlbError(EQLB_ERR_INPUT, "null handle");
        if (part != EQLB_PART_ALL && part != EQLB_PART_INTERFACE && part != EQLB_PART_INTERIOR)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_set_part: unknown part");
        if (part != EQLB_PART_ALL && (!h->interface_first || (h->flags & EQLB_FLAG_ATOMIC) || !h->h_group_off.empty()))
          throw EqlbError(EQLB_ERR_STATE, "eqlb_set_part: handle was not created with EQLB_FLAG_INTERFACE_FIRST on a "
                                          "partitioned mesh (or uses atomics / grouped patches)");
        h->part = part;
        h->win_lo = (part == EQLB_PART_INTERIOR) ? h->ncolours : 0;
        h->win_hi = (part == EQLB_PART_INTERFACE) ? h->ncolours : (1 << 30);
      });
}

int eqlb_get_launch_order(eqlb_handle* h, int32_t* order, int32_t* nchunk, int32_t* seg_off)
{
  return guarded(
      [&]
      {
        if (!h || !h->bcs_set)
          throw EqlbError(EQLB_ERR_STATE, "eqlb_get_launch_order: boundary conditions not set");
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (order && h->nactive > 0)
          CUDA_CHECK(cudaMemcpy(order, h->d_order.p, (size_t)h->nactive * sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (nchunk)
          *nchunk = h->nchunk;
        if (seg_off)
          std::memcpy(seg_off, h->h_colour_off.data(), h->h_colour_off.size() * sizeof(int32_t));
      });
}

int eqlb_get_staged_flux(eqlb_handle* h, int r, int is_ev, double** device_ptr, int64_t* n)
{
  return guarded(
      [&]
      {
        if (!h || !device_ptr || r < 0 || r >= h->nrhs)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_get_staged_flux: bad argument");
        const size_t nS = is_ev ? (size_t)h->nfct * h->k + (size_t)h->ncell * (h->k * h->k - h->k) : (size_t)h->ncell * h->nrt;
        if (h->d_stage_sigma.n < nS * (size_t)h->nrhs)
          throw EqlbError(EQLB_ERR_STATE, "eqlb_get_staged_flux: no host-buffer call (EQLB_HOST / EQLB_HOST_ZEROED) has run on this handle");
        *device_ptr = h->d_stage_sigma.p + (size_t)r * nS;
        if (n)
          *n = (int64_t)nS;
      });
}

// Common driver of eqlb_se_run / eqlb_ev_run: device pointers are used in place; host
// pointers are staged, either in one piece or - with EQLB_FLAG_HOST_PIPELINE - stage by
// stage on three streams so that both PCIe directions and the SMs work at the same time.
static void run_equilibration(eqlb_handle* h, bool ev, const double* const* G, const double* const* f,
                              double* const* sigma, double* korn, int memspace)
{
  const char* who = ev ? "eqlb_ev_run" : "eqlb_se_run";
  if (!h || !G || !f || !sigma)
    throw EqlbError(EQLB_ERR_INPUT, std::string(who) + ": null argument");
  if (!h->bcs_set)
    throw EqlbError(EQLB_ERR_STATE, std::string(who) + ": call eqlb_set_bcs first");
  if (memspace != EQLB_HOST && memspace != EQLB_DEVICE && memspace != EQLB_HOST_ZEROED && memspace != EQLB_HOST_IN)
    throw EqlbError(EQLB_ERR_INPUT, std::string(who) + ": unknown memspace");
  for (int r = 0; r < h->nrhs; ++r)
    if (!G[r] || !f[r] || !sigma[r])
      throw EqlbError(EQLB_ERR_INPUT, "Equilibration: Input sizes does not match");
  // Korn constants of a partitioned run are partial sums on halo cells and are not exchanged: all patches only
  if (korn && h->part != EQLB_PART_ALL)
    throw EqlbError(EQLB_ERR_INPUT, std::string(who) + ": Korn constants need EQLB_PART_ALL");
  const int nrhs = h->nrhs;
  const size_t nG = (size_t)h->ncell * h->ndg * 2, nF = (size_t)h->ncell * h->ndg;
  const size_t nS = ev ? (size_t)h->nfct * h->k + (size_t)h->ncell * (h->k * h->k - h->k) : (size_t)h->ncell * h->nrt;
  const double* dG[EQLB_MAXRHS];
  const double* dF[EQLB_MAXRHS];
  double* dS[EQLB_MAXRHS];
  double* dK = korn;
  auto launch = [&]
  {
    if (ev)
      launch_ev(h, dG, dF, dS);
    else
      launch_se(h, dG, dF, dS, dK);
  };
  if (memspace == EQLB_DEVICE)
  {
    for (int r = 0; r < nrhs; ++r)
    {
      dG[r] = G[r];
      dF[r] = f[r];
      dS[r] = sigma[r];
    }
    launch();
    return;
  }
  const bool zeroed = (memspace == EQLB_HOST_ZEROED);
  const bool sigma_dev = (memspace == EQLB_HOST_IN);
  if (sigma_dev && korn)
    throw EqlbError(EQLB_ERR_INPUT, std::string(who) + ": EQLB_HOST_IN does not take Korn constants");
  h->d_stage_G.alloc(nG * nrhs);
  h->d_stage_f.alloc(nF * nrhs);
  if (!sigma_dev)
    h->d_stage_sigma.alloc(nS * nrhs);
  for (int r = 0; r < nrhs; ++r)
  {
    dG[r] = h->d_stage_G.p + r * nG;
    dF[r] = h->d_stage_f.p + r * nF;
    dS[r] = sigma_dev ? sigma[r] : h->d_stage_sigma.p + r * nS;
  }
  if (korn)
  {
    h->d_stage_korn.alloc(h->ncell);
    dK = h->d_stage_korn.p;
  }
  // (grouped boundary patches of the stress path read the accumulated global stress: not staged)
  const bool pipelined = h->nchunk > 1 && !h->interface_first && !(h->flags & EQLB_FLAG_ATOMIC) && h->h_group_off.empty()
                         && !korn && h->nseg == h->nchunk * h->ncolours;
  // large caller buffers in pageable memory (first call on a new mesh, nobody registered them): worker threads
  // stage them through the pinned pool (staged_copy.cu) - 3-5x the speed of cudaMemcpy from pageable memory and
  // no registration cost; the stage pipeline below needs page-locked buffers (eqlb_pin_host)
  auto big_pageable = [](const void* p, size_t bytes) { return bytes >= (size_t(8) << 20) && !host_is_pinned(p); };
  bool pageable = false;
  for (int r = 0; r < nrhs; ++r)
    pageable = pageable || big_pageable(G[r], nG * 8) || big_pageable(f[r], nF * 8) || (!sigma_dev && big_pageable(sigma[r], nS * 8));
  if (pageable)
  {
    CUDA_CHECK(cudaStreamSynchronize(h->stream));  // the copies run on the pool's streams
    for (int r = 0; r < nrhs; ++r)
    {
      eqlb_h2d((void*)dG[r], G[r], nG * 8);
      eqlb_h2d((void*)dF[r], f[r], nF * 8);
      if (zeroed)
        CUDA_CHECK(cudaMemsetAsync(dS[r], 0, nS * 8, h->stream));
      else if (!sigma_dev)
        eqlb_h2d(dS[r], sigma[r], nS * 8);
    }
    if (korn)
      CUDA_CHECK(cudaMemcpyAsync(dK, korn, (size_t)h->ncell * 8, cudaMemcpyHostToDevice, h->stream));
    launch();
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    for (int r = 0; r < nrhs && !sigma_dev; ++r)
      eqlb_d2h(sigma[r], dS[r], nS * 8);
    if (korn)
      CUDA_CHECK(cudaMemcpy(korn, dK, (size_t)h->ncell * 8, cudaMemcpyDeviceToHost));
    return;
  }
  if (!pipelined)
  {
    for (int r = 0; r < nrhs; ++r)
    {
      CUDA_CHECK(cudaMemcpyAsync((void*)dG[r], G[r], nG * 8, cudaMemcpyHostToDevice, h->stream));
      CUDA_CHECK(cudaMemcpyAsync((void*)dF[r], f[r], nF * 8, cudaMemcpyHostToDevice, h->stream));
      if (zeroed)
        CUDA_CHECK(cudaMemsetAsync(dS[r], 0, nS * 8, h->stream));
      else if (!sigma_dev)
        CUDA_CHECK(cudaMemcpyAsync(dS[r], sigma[r], nS * 8, cudaMemcpyHostToDevice, h->stream));
    }
    if (korn)
      CUDA_CHECK(cudaMemcpyAsync(dK, korn, (size_t)h->ncell * 8, cudaMemcpyHostToDevice, h->stream));
    launch();
    for (int r = 0; r < nrhs && !sigma_dev; ++r)
      CUDA_CHECK(cudaMemcpyAsync(sigma[r], dS[r], nS * 8, cudaMemcpyDeviceToHost, h->stream));
    if (korn)
      CUDA_CHECK(cudaMemcpyAsync(korn, dK, (size_t)h->ncell * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return;
  }

  // ---- staged: copy-in stream, compute stream (the caller's), copy-out stream ----
  ensure_slabs(h);
  const int nst = h->nchunk;
  if (!h->s_h2d)
  {
    CUDA_CHECK(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
  }
  while ((int)h->ev_in.size() < nst)
  {
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    h->ev_in.push_back(a);
    h->ev_done.push_back(b);
  }
  // everything queued on the caller's stream so far happens before the first copy
  CUDA_CHECK(cudaEventRecord(h->ev_start, h->stream));
  CUDA_CHECK(cudaStreamWaitEvent(h->s_h2d, h->ev_start, 0));
  for (int r = 0; r < nrhs; ++r)
  {
    if (zeroed)
      CUDA_CHECK(cudaMemsetAsync(dS[r], 0, nS * 8, h->stream));
    else if (!sigma_dev)
      CUDA_CHECK(cudaMemcpyAsync(dS[r], sigma[r], nS * 8, cudaMemcpyHostToDevice, h->s_h2d));
  }
  auto cell_lo = [&](int sidx) { return (size_t)((long)h->ncell * sidx / nst); };
  for (int sidx = 0; sidx < nst; ++sidx)
  {
    const size_t c0 = cell_lo(sidx), nc = cell_lo(sidx + 1) - c0;
    for (int r = 0; r < nrhs && nc; ++r)
    {
      const size_t og = c0 * h->ndg * 2, of = c0 * h->ndg;
      CUDA_CHECK(cudaMemcpyAsync((void*)(dG[r] + og), G[r] + og, nc * h->ndg * 16, cudaMemcpyHostToDevice, h->s_h2d));
      CUDA_CHECK(cudaMemcpyAsync((void*)(dF[r] + of), f[r] + of, nc * h->ndg * 8, cudaMemcpyHostToDevice, h->s_h2d));
    }
    CUDA_CHECK(cudaEventRecord(h->ev_in[sidx], h->s_h2d));
  }
  const auto& slabs = ev ? h->h_ev_slabs : h->h_se_slabs;
  try
  {
    for (int sidx = 0; sidx < nst; ++sidx)
    {
      CUDA_CHECK(cudaStreamWaitEvent(h->stream, h->ev_in[sidx], 0));
      h->win_lo = sidx * h->ncolours;
      h->win_hi = (sidx + 1) * h->ncolours;
      launch();
      CUDA_CHECK(cudaEventRecord(h->ev_done[sidx], h->stream));
      CUDA_CHECK(cudaStreamWaitEvent(h->s_d2h, h->ev_done[sidx], 0));
      if (sigma_dev)
        continue;
      for (const auto& sl : slabs)
        if (sl.final_stage == sidx && sl.len)
          for (int r = 0; r < nrhs; ++r)
            CUDA_CHECK(cudaMemcpyAsync(sigma[r] + sl.off, dS[r] + sl.off, sl.len * 8, cudaMemcpyDeviceToHost, h->s_d2h));
    }
  }
  catch (...)
  {
    h->win_lo = 0;
    h->win_hi = 1 << 30;
    throw;
  }
  h->win_lo = 0;
  h->win_hi = 1 << 30;
  if (sigma_dev)
  {
    CUDA_CHECK(cudaStreamSynchronize(h->s_h2d));  // the host inputs may be reused; kernels stay queued
    return;
  }
  CUDA_CHECK(cudaStreamSynchronize(h->s_d2h));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

int eqlb_se_run(eqlb_handle* h, const double* const* G, const double* const* f, double* const* sigma, double* korn,
                int memspace)
{
  return guarded([&] { run_equilibration(h, false, G, f, sigma, korn, memspace); });
}

int eqlb_ev_run(eqlb_handle* h, const double* const* G, const double* const* f, double* const* sigma, int memspace)
{
  return guarded([&] { run_equilibration(h, true, G, f, sigma, nullptr, memspace); });
}

int eqlb_local_project(eqlb_handle* h, int nfun, const double* const* qvals, double* const* out, int memspace)
{
  return guarded(
      [&]
      {
        if (!h || !qvals || !out || nfun < 1)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_local_project: bad argument");
        if (!h->dg_identity)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_local_project: the DG_p functions must use the DOLFINx layout cell*ndg + i (eqlb_mesh.dg_dofmap)");
        const size_t nin = (size_t)h->ncell * h->nq, nout = (size_t)h->ncell * h->ndg;
        std::vector<const double*> dq(nfun);
        std::vector<double*> dout(nfun);
        DevBuf<double> stage_in, stage_out;
        if (memspace == EQLB_DEVICE)
        {
          for (int i = 0; i < nfun; ++i)
          {
            dq[i] = qvals[i];
            dout[i] = out[i];
          }
        }
        else
        {
          stage_in.alloc(nin * nfun);
          stage_out.alloc(nout * nfun);
          for (int i = 0; i < nfun; ++i)
          {
            CUDA_CHECK(cudaMemcpyAsync(stage_in.p + i * nin, qvals[i], nin * 8, cudaMemcpyHostToDevice, h->stream));
            dq[i] = stage_in.p + i * nin;
            dout[i] = stage_out.p + i * nout;
          }
        }
        launch_project(h, nfun, dq.data(), dout.data());
        if (memspace != EQLB_DEVICE)
        {
          for (int i = 0; i < nfun; ++i)
            CUDA_CHECK(cudaMemcpyAsync(out[i], dout[i], nout * 8, cudaMemcpyDeviceToHost, h->stream));
          CUDA_CHECK(cudaStreamSynchronize(h->stream));
        }
      });
}

int eqlb_flux_l2norm(eqlb_handle* h, int nfun, const double* const* sigma, double* const* eta2, int memspace)
{
  return guarded(
      [&]
      {
        if (!h || !sigma || !eta2 || nfun < 1)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_flux_l2norm: bad argument");
        if (memspace != EQLB_HOST && memspace != EQLB_DEVICE)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_flux_l2norm: memspace must be EQLB_HOST or EQLB_DEVICE");
        const size_t nin = (size_t)h->ncell * h->nrt, nout = (size_t)h->ncell;
        std::vector<const double*> ds(nfun);
        std::vector<double*> dout(nfun);
        DevBuf<double> stage_in, stage_out;
        if (memspace == EQLB_DEVICE)
        {
          for (int i = 0; i < nfun; ++i)
          {
            ds[i] = sigma[i];
            dout[i] = eta2[i];
          }
        }
        else
        {
          stage_in.alloc(nin * nfun);
          stage_out.alloc(nout * nfun);
          for (int i = 0; i < nfun; ++i)
          {
            CUDA_CHECK(cudaMemcpyAsync(stage_in.p + i * nin, sigma[i], nin * 8, cudaMemcpyHostToDevice, h->stream));
            ds[i] = stage_in.p + i * nin;
            dout[i] = stage_out.p + i * nout;
          }
        }
        launch_flux_norm(h, nfun, ds.data(), dout.data());
        if (memspace != EQLB_DEVICE)
        {
          for (int i = 0; i < nfun; ++i)
            CUDA_CHECK(cudaMemcpyAsync(eta2[i], dout[i], nout * 8, cudaMemcpyDeviceToHost, h->stream));
          CUDA_CHECK(cudaStreamSynchronize(h->stream));
        }
      });
}

int eqlb_ev_to_basix_rt(eqlb_handle* h, int nfun, const double* const* sigma_hier, double* const* sigma_basix, int memspace)
{
  return guarded(
      [&]
      {
        if (!h || !sigma_hier || !sigma_basix || nfun < 1)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_ev_to_basix_rt: bad argument");
        if (memspace != EQLB_HOST && memspace != EQLB_DEVICE)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_ev_to_basix_rt: memspace must be EQLB_HOST or EQLB_DEVICE");
        const size_t n = (size_t)h->nfct * h->k + (size_t)h->ncell * (h->k * h->k - h->k);
        std::vector<const double*> din(nfun);
        std::vector<double*> dout(nfun);
        DevBuf<double> st_in, st_out;
        for (int i = 0; i < nfun; ++i)
          if (sigma_hier[i] == sigma_basix[i])
            throw EqlbError(EQLB_ERR_INPUT, "eqlb_ev_to_basix_rt: conversion is out of place");
        if (memspace == EQLB_DEVICE)
          for (int i = 0; i < nfun; ++i)
          {
            din[i] = sigma_hier[i];
            dout[i] = sigma_basix[i];
          }
        else
        {
          st_in.alloc(n * nfun);
          st_out.alloc(n * nfun);
          for (int i = 0; i < nfun; ++i)
          {
            CUDA_CHECK(cudaMemcpyAsync(st_in.p + i * n, sigma_hier[i], n * 8, cudaMemcpyHostToDevice, h->stream));
            din[i] = st_in.p + i * n;
            dout[i] = st_out.p + i * n;
          }
        }
        launch_ev_to_basix(h, nfun, din.data(), dout.data());
        if (memspace != EQLB_DEVICE)
        {
          for (int i = 0; i < nfun; ++i)
            CUDA_CHECK(cudaMemcpyAsync(sigma_basix[i], dout[i], n * 8, cudaMemcpyDeviceToHost, h->stream));
          CUDA_CHECK(cudaStreamSynchronize(h->stream));
        }
      });
}

int64_t eqlb_launch_count(eqlb_handle* h) { return h ? h->launches : 0; }

} // extern "C"

namespace
{
// device views of caller vectors: the pointers themselves (EQLB_DEVICE) or staged copies (host)
struct Staged
{
  std::vector<std::unique_ptr<DevBuf<double>>> bufs;
  std::vector<const double*> in(eqlb_handle* h, int n, const double* const* v, size_t len, bool host)
  {
    std::vector<const double*> out(n, nullptr);
    for (int i = 0; i < n; ++i)
    {
      if (!v || !v[i])
        continue;
      if (!host)
      {
        out[i] = v[i];
        continue;
      }
      auto b = std::make_unique<DevBuf<double>>();
      b->alloc(len);
      CUDA_CHECK(cudaMemcpyAsync(b->p, v[i], len * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      out[i] = b->p;
      bufs.push_back(std::move(b));
    }
    return out;
  }
  std::vector<double*> out(eqlb_handle* h, int n, double* const* v, size_t len, bool host, bool upload)
  {
    std::vector<double*> o(n, nullptr);
    for (int i = 0; i < n; ++i)
    {
      if (!v || !v[i])
        continue;
      if (!host)
      {
        o[i] = v[i];
        continue;
      }
      auto b = std::make_unique<DevBuf<double>>();
      b->alloc(len);
      if (upload)
        CUDA_CHECK(cudaMemcpyAsync(b->p, v[i], len * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      else
        b->zero(h->stream);
      o[i] = b->p;
      bufs.push_back(std::move(b));
    }
    return o;
  }
  static void back(eqlb_handle* h, int n, double* const* host, const std::vector<double*>& dev, size_t len)
  {
    for (int i = 0; i < n; ++i)
      if (host && host[i] && dev[i])
        CUDA_CHECK(cudaMemcpyAsync(host[i], dev[i], len * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  }
};

void check_memspace3(int memspace, const char* who)
{
  if (memspace != EQLB_HOST && memspace != EQLB_DEVICE && memspace != EQLB_HOST_ZEROED)
    throw EqlbError(EQLB_ERR_INPUT, std::string(who) + ": memspace must be EQLB_HOST, EQLB_DEVICE or EQLB_HOST_ZEROED");
}

// projection + equilibration: G, F live in device scratch only
void run_primal(eqlb_handle* h, bool ev, const double* const* uh, const double* const* fh, double* const* sigma, double* korn,
                int memspace)
{
  if (!h || !uh || !fh || !sigma)
    throw EqlbError(EQLB_ERR_INPUT, "eqlb_*_run_primal: null argument");
  if (!h->dg_identity)
    throw EqlbError(EQLB_ERR_INPUT, "eqlb_*_run_primal: the DG_p functions must use the DOLFINx layout cell*ndg + i (eqlb_mesh.dg_dofmap)");
  if (!h->bcs_set)
    throw EqlbError(EQLB_ERR_STATE, "boundary conditions not set (eqlb_set_bcs)");
  check_memspace3(memspace, "eqlb_*_run_primal");
  const bool host = memspace != EQLB_DEVICE;
  const int n = h->nrhs;
  const size_t nG = (size_t)h->ncell * h->ndg * 2, nF = (size_t)h->ncell * h->ndg;
  const size_t nS = ev ? (size_t)h->nfct * h->k + (size_t)h->ncell * (h->k * h->k - h->k) : (size_t)h->ncell * h->nrt;
  Staged st;
  auto du = st.in(h, n, uh, (size_t)h->pk_ndofs, host);
  auto df = st.in(h, n, fh, (size_t)h->pk_ndofs, host);
  h->d_stage_G.alloc(nG * n);
  h->d_stage_f.alloc(nF * n);
  std::vector<double*> G(n), F(n);
  std::vector<const double*> Gc(n), Fc(n);
  for (int i = 0; i < n; ++i)
  {
    G[i] = h->d_stage_G.p + i * nG;
    F[i] = h->d_stage_f.p + i * nF;
    Gc[i] = G[i];
    Fc[i] = F[i];
  }
  launch_primal_project(h, n, du.data(), df.data(), G.data(), F.data());
  auto dS = st.out(h, n, sigma, nS, host, memspace == EQLB_HOST);
  DevBuf<double> dk;
  double* dkorn = korn;
  if (korn && host)
  {
    dk.alloc(h->ncell);
    CUDA_CHECK(cudaMemcpyAsync(dk.p, korn, (size_t)h->ncell * 8, cudaMemcpyHostToDevice, h->stream));
    dkorn = dk.p;
  }
  if (ev)
    launch_ev(h, Gc.data(), Fc.data(), dS.data());
  else
    launch_se(h, Gc.data(), Fc.data(), dS.data(), dkorn);
  if (host)
  {
    if (korn)
      CUDA_CHECK(cudaMemcpyAsync(korn, dkorn, (size_t)h->ncell * 8, cudaMemcpyDeviceToHost, h->stream));
    Staged::back(h, n, sigma, dS, nS);
  }
}
} // namespace

extern "C"
{
int eqlb_set_primal_space(eqlb_handle* h, const int32_t* pk_dofmap, int64_t ndofs)
{
  return guarded(
      [&]
      {
        if (!h || !pk_dofmap || ndofs < 1)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_set_primal_space: bad argument");
        if (h->npk < 1 || !h->d_primal.p)
          throw EqlbError(EQLB_ERR_STATE, "eqlb_set_primal_space: the tables carry no primal-space data (pk_*)");
        const size_t n = (size_t)h->ncell * h->npk;
        for (size_t i = 0; i < n; ++i)
          if (pk_dofmap[i] < 0 || pk_dofmap[i] >= ndofs)
            throw EqlbError(EQLB_ERR_INPUT, "eqlb_set_primal_space: dofmap entry out of range");
        h->d_pk_dofmap.upload(pk_dofmap, n);
        h->pk_ndofs = ndofs;
      });
}

int eqlb_project_primal(eqlb_handle* h, int nfun, const double* const* uh, const double* const* fh, double* const* G,
                        double* const* F, int memspace)
{
  return guarded(
      [&]
      {
        if (!h || nfun < 1)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_project_primal: bad argument");
        if (!h->dg_identity)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_project_primal: the DG_p functions must use the DOLFINx layout cell*ndg + i (eqlb_mesh.dg_dofmap)");
        if (memspace != EQLB_HOST && memspace != EQLB_DEVICE)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_project_primal: memspace must be EQLB_HOST or EQLB_DEVICE");
        const bool host = memspace == EQLB_HOST;
        const size_t nG = (size_t)h->ncell * h->ndg * 2, nF = (size_t)h->ncell * h->ndg;
        Staged st;
        auto du = st.in(h, nfun, uh, (size_t)h->pk_ndofs, host);
        auto df = st.in(h, nfun, fh, (size_t)h->pk_ndofs, host);
        auto dG = st.out(h, nfun, G, nG, host, false);
        auto dF = st.out(h, nfun, F, nF, host, false);
        for (int i = 0; i < nfun; ++i)
          if ((du[i] && !dG[i]) || (df[i] && !dF[i]))
            throw EqlbError(EQLB_ERR_INPUT, "eqlb_project_primal: output missing for a given input");
        launch_primal_project(h, nfun, du.data(), df.data(), dG.data(), dF.data());
        if (host)
        {
          Staged::back(h, nfun, G, dG, nG);
          Staged::back(h, nfun, F, dF, nF);
        }
      });
}

int eqlb_ev_run_primal(eqlb_handle* h, const double* const* uh, const double* const* fh, double* const* sigma, int memspace)
{
  return guarded([&] { run_primal(h, true, uh, fh, sigma, nullptr, memspace); });
}

int eqlb_se_run_primal(eqlb_handle* h, const double* const* uh, const double* const* fh, double* const* sigma, double* korn,
                       int memspace)
{
  return guarded([&] { run_primal(h, false, uh, fh, sigma, korn, memspace); });
}

int eqlb_estimate_poisson(eqlb_handle* h, int nfun, const double* const* sigma, const double* const* uh,
                          const double* const* fh, double* const* eta_sig2, double* const* eta_osc2, int is_ev, int memspace)
{
  return guarded(
      [&]
      {
        if (!h || nfun < 1 || !sigma || !eta_sig2 || !eta_osc2)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_estimate_poisson: bad argument");
        if (!h->dg_identity)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_estimate_poisson: the DG_p functions must use the DOLFINx layout cell*ndg + i (eqlb_mesh.dg_dofmap)");
        if (memspace != EQLB_HOST && memspace != EQLB_DEVICE)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_estimate_poisson: memspace must be EQLB_HOST or EQLB_DEVICE");
        const bool host = memspace == EQLB_HOST;
        const size_t nS = is_ev ? (size_t)h->nfct * h->k + (size_t)h->ncell * (h->k * h->k - h->k) : (size_t)h->ncell * h->nrt;
        Staged st;
        auto dS = st.in(h, nfun, sigma, nS, host);
        auto du = st.in(h, nfun, uh, (size_t)h->pk_ndofs, host);
        auto df = st.in(h, nfun, fh, (size_t)h->pk_ndofs, host);
        auto e1 = st.out(h, nfun, eta_sig2, h->ncell, host, false);
        auto e2 = st.out(h, nfun, eta_osc2, h->ncell, host, false);
        for (int i = 0; i < nfun; ++i)
          if (!dS[i] || !e1[i] || !e2[i])
            throw EqlbError(EQLB_ERR_INPUT, "eqlb_estimate_poisson: null vector");
        launch_estimate_poisson(h, nfun, dS.data(), uh ? du.data() : nullptr, fh ? df.data() : nullptr, e1.data(), e2.data(), is_ev);
        if (host)
        {
          Staged::back(h, nfun, eta_sig2, e1, h->ncell);
          Staged::back(h, nfun, eta_osc2, e2, h->ncell);
        }
      });
}

int eqlb_estimate_elasticity(eqlb_handle* h, const double* const* dsig, const double* const* sigma_h, const double* const* fh,
                             const double* korn, double pi_1, double* const* eta, int memspace)
{
  return guarded(
      [&]
      {
        if (!h || !dsig || !eta)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_estimate_elasticity: bad argument");
        if (!h->dg_identity)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_estimate_elasticity: the DG_p functions must use the DOLFINx layout cell*ndg + i (eqlb_mesh.dg_dofmap)");
        if (memspace != EQLB_HOST && memspace != EQLB_DEVICE)
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_estimate_elasticity: memspace must be EQLB_HOST or EQLB_DEVICE");
        const bool host = memspace == EQLB_HOST;
        Staged st;
        auto dS = st.in(h, 2, dsig, (size_t)h->ncell * h->nrt, host);
        auto dH = st.in(h, 2, sigma_h, (size_t)h->ncell * h->ndg * 2, host);
        auto df = st.in(h, 2, fh, (size_t)h->pk_ndofs, host);
        const double* kp[1] = {korn};
        auto dk = st.in(h, 1, korn ? kp : nullptr, h->ncell, host);
        auto de = st.out(h, 3, eta, h->ncell, host, false);
        if (!dS[0] || !dS[1] || !de[0] || !de[1] || !de[2])
          throw EqlbError(EQLB_ERR_INPUT, "eqlb_estimate_elasticity: null vector");
        launch_estimate_elasticity(h, dS.data(), sigma_h ? dH.data() : nullptr, fh ? df.data() : nullptr, dk[0], pi_1, de.data());
        if (host)
          Staged::back(h, 3, eta, de, h->ncell);
      });
}
} // extern "C"
