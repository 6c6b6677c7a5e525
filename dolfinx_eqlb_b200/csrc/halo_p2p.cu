// Halo sum of a partitioned equilibration over NVLink peer memory, without NCCL on the data
// path.  One process per GPU; every rank owns one communication buffer (cudaMalloc, exported
// with a CUDA IPC handle and mapped by its neighbours):
//
//   [flags: 2 parities x EQLB_HALO_MAXNEIGH uint64][send area of neighbour slot 0 | 1 | ...]
//
// One kernel launch per exchange on every rank:
//   1. pack: the values of the shared DOFs are gathered into the rank's own send areas;
//   2. the last CTA to finish packing publishes the epoch in the flag slot it owns inside each
//      NEIGHBOUR's buffer (store over NVLink after a system-wide fence);
//   3. every CTA waits until the neighbour's epoch has arrived in the local flag slot and adds
//      the neighbour's send area (read over NVLink) to the shared DOFs, neighbours in ascending
//      rank order -> the same bits as the NCCL send/recv + index_add implementation.
// Send areas and flags are double buffered by epoch parity: a rank can only overwrite parity p
// after it has consumed its neighbours' epoch e+1, which they publish after consuming epoch e.
#include <algorithm>
#include <cstring>
#include <memory>

#include "eqlb_internal.cuh"

#define EQLB_HALO_MAXNEIGH 8

struct eqlb_halo
{
  int nneigh = 0, nrhs_max = 0;
  unsigned long long epoch = 0;
  char* buf = nullptr;  // own communication buffer (device)
  size_t buf_bytes = 0;
  int64_t count[EQLB_HALO_MAXNEIGH];        // shared DOFs per neighbour
  size_t send_off[EQLB_HALO_MAXNEIGH];      // byte offset of the send area for neighbour n in the OWN buffer
  size_t peer_send_off[EQLB_HALO_MAXNEIGH]; // byte offset of the area addressed to this rank in neighbour n's buffer
  int peer_slot[EQLB_HALO_MAXNEIGH];        // flag slot this rank owns in neighbour n's buffer
  char* peer_buf[EQLB_HALO_MAXNEIGH] = {};  // mapped neighbour buffers (null until eqlb_halo_connect)
  DevBuf<int64_t> idx[EQLB_HALO_MAXNEIGH];  // local DOF indices shared with neighbour n (ordered by global id)
  DevBuf<unsigned long long> done;          // [0] CTA counter of the barriers (monotone over launches), [1] error word
  int grid = 1;                             // fixed launch grid
  ~eqlb_halo()
  {
    for (int n = 0; n < nneigh; ++n)
      if (peer_buf[n])
        cudaIpcCloseMemHandle(peer_buf[n]);
    if (buf)
      cudaFree(buf);
  }
};

namespace
{

struct HaloArgs
{
  int nneigh, nrhs;
  unsigned long long epoch;
  double* x[EQLB_MAXRHS];
  const int64_t* idx[EQLB_HALO_MAXNEIGH];
  int64_t count[EQLB_HALO_MAXNEIGH];
  double* send[EQLB_HALO_MAXNEIGH];                      // own send area (current parity)
  const double* recv[EQLB_HALO_MAXNEIGH];                // neighbour's send area addressed to this rank
  unsigned long long* peer_flag[EQLB_HALO_MAXNEIGH];     // flag slot in the neighbour's buffer
  const unsigned long long* my_flag[EQLB_HALO_MAXNEIGH]; // flag slot the neighbour writes in the own buffer
};

// Spins are bounded: a neighbour that never arrives (it failed, or its process died) must not hang this GPU.
// After about 4 s (clock64 at ~2 GHz) the waiting thread records the error in done[1] and carries on; the
// result of that exchange is then incomplete and eqlb_halo_status reports it.
constexpr long long EQLB_HALO_SPIN_CYCLES = 8000000000ll;

__global__ void __launch_bounds__(256) halo_push_add_kernel(HaloArgs a, unsigned long long* done)
{
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  // 1. pack
  for (int n = 0; n < a.nneigh; ++n)
    for (int64_t i = tid; i < a.count[n]; i += nth)
    {
      const int64_t d = a.idx[n][i];
      for (int r = 0; r < a.nrhs; ++r)
        a.send[n][(int64_t)r * a.count[n] + i] = a.x[r][d];
    }
  // 2. publish: make the packed values visible system wide; the last CTA raises the flags.
  //    `done` counts CTA arrivals over all launches (grid size and number of barriers per
  //    launch are fixed per handle).  Barrier 0: no CTA starts adding before every CTA of this
  //    rank has packed; barrier n > 0: neighbour n is added after neighbour n-1 everywhere
  //    (a DOF can be shared with two neighbours).
  const unsigned long long base = (a.epoch - 1) * (unsigned long long)a.nneigh * gridDim.x;
  auto grid_barrier = [&](int kbar, bool raise_flags)
  {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
      const unsigned long long target = base + (unsigned long long)(kbar + 1) * gridDim.x;
      const unsigned long long prev = atomicAdd(done, 1ull);
      if (prev + 1 == target)
      {
        if (raise_flags)
        {
          for (int n = 0; n < a.nneigh; ++n)
            *reinterpret_cast<volatile unsigned long long*>(a.peer_flag[n]) = a.epoch;
          __threadfence_system();
        }
      }
      else
      {
        const volatile unsigned long long* dn = done;
        const long long t0 = clock64();
        while (*dn < target)
        {
          __nanosleep(50);
          if (clock64() - t0 > EQLB_HALO_SPIN_CYCLES)
          {
            atomicMax(done + 1, 1ull);  // grid barrier timed out (grid not co-resident?)
            break;
          }
        }
      }
    }
    __syncthreads();
  };
  grid_barrier(0, true);
  // 3. consume, neighbours in ascending rank order
  for (int n = 0; n < a.nneigh; ++n)
  {
    if (n > 0)
      grid_barrier(n, false);
    if (threadIdx.x == 0)
    {
      const volatile unsigned long long* f = a.my_flag[n];
      const long long t0 = clock64();
      while (*f < a.epoch)
      {
        __nanosleep(100);
        if (clock64() - t0 > EQLB_HALO_SPIN_CYCLES)
        {
          atomicMax(done + 1, 2ull);  // neighbour n never published this epoch
          break;
        }
      }
      __threadfence_system();
    }
    __syncthreads();
    for (int64_t i = tid; i < a.count[n]; i += nth)
    {
      const int64_t d = a.idx[n][i];
      for (int r = 0; r < a.nrhs; ++r)
        a.x[r][d] += __ldcv(a.recv[n] + (int64_t)r * a.count[n] + i);
    }
  }
}

size_t align256(size_t v) { return (v + 255) / 256 * 256; }
constexpr size_t FLAG_BYTES = 2 * EQLB_HALO_MAXNEIGH * sizeof(unsigned long long);

} // namespace

extern "C"
{

int eqlb_halo_create(int nneigh, const int64_t* counts, const int64_t* const* idx, int nrhs_max, eqlb_halo** out,
                     unsigned char* ipc_handle_out /*[64]*/, int64_t* send_off_out /*[nneigh]*/)
{
  try
  {
    if (!out || !ipc_handle_out || nneigh < 0 || nneigh > EQLB_HALO_MAXNEIGH || nrhs_max < 1 || nrhs_max > EQLB_MAXRHS)
      throw EqlbError(EQLB_ERR_INPUT, "eqlb_halo_create: bad argument");
    std::unique_ptr<eqlb_halo> h(new eqlb_halo());
    h->nneigh = nneigh;
    h->nrhs_max = nrhs_max;
    size_t off = align256(FLAG_BYTES);
    for (int n = 0; n < nneigh; ++n)
    {
      h->count[n] = counts[n];
      h->send_off[n] = off;
      if (send_off_out)
        send_off_out[n] = (int64_t)off;
      off += align256(2 * (size_t)nrhs_max * counts[n] * sizeof(double));  // two parities
      h->idx[n].upload(idx[n], counts[n]);
      h->peer_buf[n] = nullptr;
    }
    h->buf_bytes = std::max<size_t>(off, 256);
    CUDA_CHECK(cudaMalloc(&h->buf, h->buf_bytes));
    CUDA_CHECK(cudaMemset(h->buf, 0, h->buf_bytes));
    h->done.alloc(2);
    CUDA_CHECK(cudaMemset(h->done.p, 0, 2 * sizeof(unsigned long long)));
    {
      int64_t maxc = 0;
      for (int n = 0; n < nneigh; ++n)
        maxc = std::max(maxc, counts[n]);
      int nsm = 148, dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
      // the grid barrier needs every CTA resident at once: cooperative launch, grid bounded by the occupancy
      int per_sm = 0;
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, halo_push_add_kernel, 256, 0));
      int coop = 0;
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
      if (!coop || per_sm < 1)
        throw EqlbError(EQLB_ERR_CUDA, "eqlb_halo_create: device cannot launch the halo kernel cooperatively");
      h->grid = (int)std::max<int64_t>(1, std::min<int64_t>((maxc + 255) / 256, (int64_t)nsm * std::min(per_sm, 1)));
    }
    cudaIpcMemHandle_t ipc;
    CUDA_CHECK(cudaIpcGetMemHandle(&ipc, h->buf));
    static_assert(sizeof(ipc) == 64, "CUDA IPC handle size");
    std::memcpy(ipc_handle_out, &ipc, 64);
    CUDA_CHECK(cudaDeviceSynchronize());
    *out = h.release();
    return EQLB_OK;
  }
  catch (const EqlbError& e)
  {
    eqlb_set_error(e.what());
    return e.code;
  }
  catch (const std::exception& e)
  {
    eqlb_set_error(e.what());
    return EQLB_ERR_CUDA;
  }
}

int eqlb_halo_connect(eqlb_halo* h, int n, const unsigned char* peer_ipc_handle /*[64]*/, int64_t peer_send_off, int peer_slot)
{
  try
  {
    if (!h || n < 0 || n >= h->nneigh || !peer_ipc_handle || peer_slot < 0 || peer_slot >= EQLB_HALO_MAXNEIGH)
      throw EqlbError(EQLB_ERR_INPUT, "eqlb_halo_connect: bad argument");
    cudaIpcMemHandle_t ipc;
    std::memcpy(&ipc, peer_ipc_handle, 64);
    void* p = nullptr;
    CUDA_CHECK(cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->peer_buf[n] = static_cast<char*>(p);
    h->peer_send_off[n] = (size_t)peer_send_off;
    h->peer_slot[n] = peer_slot;
    return EQLB_OK;
  }
  catch (const EqlbError& e)
  {
    eqlb_set_error(e.what());
    return e.code;
  }
  catch (const std::exception& e)
  {
    eqlb_set_error(e.what());
    return EQLB_ERR_CUDA;
  }
}

int eqlb_halo_apply(eqlb_halo* h, double* const* x, int nrhs, void* cuda_stream)
{
  try
  {
    if (!h || !x || nrhs < 1 || nrhs > h->nrhs_max)
      throw EqlbError(EQLB_ERR_INPUT, "eqlb_halo_apply: bad argument");
    if (h->nneigh == 0)
      return EQLB_OK;
    HaloArgs a{};
    a.nneigh = h->nneigh;
    a.nrhs = nrhs;
    a.epoch = h->epoch + 1;  // committed below, once the launch has succeeded
    const int par = (int)(a.epoch & 1);
    for (int r = 0; r < nrhs; ++r)
      a.x[r] = x[r];
    for (int n = 0; n < h->nneigh; ++n)
    {
      if (!h->peer_buf[n])
        throw EqlbError(EQLB_ERR_STATE, "eqlb_halo_apply: neighbour not connected");
      const size_t par_bytes = (size_t)h->nrhs_max * h->count[n] * sizeof(double);
      a.idx[n] = h->idx[n].p;
      a.count[n] = h->count[n];
      a.send[n] = reinterpret_cast<double*>(h->buf + h->send_off[n] + par * par_bytes);
      a.recv[n] = reinterpret_cast<const double*>(h->peer_buf[n] + h->peer_send_off[n] + par * par_bytes);
      a.peer_flag[n] = reinterpret_cast<unsigned long long*>(h->peer_buf[n]) + par * EQLB_HALO_MAXNEIGH + h->peer_slot[n];
      a.my_flag[n] = reinterpret_cast<const unsigned long long*>(h->buf) + par * EQLB_HALO_MAXNEIGH + n;
    }
    unsigned long long* donep = h->done.p;
    void* params[] = {&a, &donep};
    CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)halo_push_add_kernel, dim3(h->grid), dim3(256), params, 0,
                                           (cudaStream_t)cuda_stream));
    h->epoch = a.epoch;
    return EQLB_OK;
  }
  catch (const EqlbError& e)
  {
    eqlb_set_error(e.what());
    return e.code;
  }
  catch (const std::exception& e)
  {
    eqlb_set_error(e.what());
    return EQLB_ERR_CUDA;
  }
}

int eqlb_halo_status(eqlb_halo* h, void* cuda_stream)
{
  try
  {
    if (!h)
      throw EqlbError(EQLB_ERR_INPUT, "eqlb_halo_status: null handle");
    unsigned long long err = 0;
    CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)cuda_stream));
    CUDA_CHECK(cudaMemcpy(&err, h->done.p + 1, sizeof(err), cudaMemcpyDeviceToHost));
    if (err == 1)
      throw EqlbError(EQLB_ERR_CUDA, "halo exchange: grid barrier timed out (kernel not co-resident)");
    if (err)
      throw EqlbError(EQLB_ERR_CUDA, "halo exchange: a neighbour did not publish its values in time");
    return EQLB_OK;
  }
  catch (const EqlbError& e)
  {
    eqlb_set_error(e.what());
    return e.code;
  }
  catch (const std::exception& e)
  {
    eqlb_set_error(e.what());
    return EQLB_ERR_CUDA;
  }
}

// The caller has to make sure (collective barrier + device synchronisation on every rank) that no neighbour
// still reads this rank's buffer: `P2PHaloExchange.close()` does.
void eqlb_halo_destroy(eqlb_halo* h) { delete h; }

} // extern "C"
