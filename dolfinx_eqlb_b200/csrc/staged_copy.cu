// Pageable host memory <-> device at (close to) pinned speed.
//
// The reference's data live in ordinary host memory (numpy / PETSc arrays).  `cudaMemcpy` from pageable memory
// goes through the driver's single staging buffer: 3-6 GB/s measured for the 404 MB of mesh arrays of a
// 1024^2 crossed mesh (70-140 ms, the largest item of `eqlb_create`); page-locking the caller's arrays
// (`cudaHostRegister`) costs about as much as one such copy and only pays for repeated calls.  Here worker
// threads copy chunks into a process-wide pool of pinned buffers (memcpy of several cores in parallel) while
// the DMA engine drains the previous chunks: a first call on a new mesh no longer pays for registration.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "eqlb_internal.cuh"

namespace
{

constexpr size_t CHUNK = 4u << 20;  // bytes per pinned buffer
constexpr int NBUF = 2;             // buffers per worker (memcpy of chunk i+1 overlaps the DMA of chunk i)

struct Worker
{
  char* buf[NBUF] = {nullptr, nullptr};
  cudaEvent_t ev[NBUF] = {nullptr, nullptr};
  cudaStream_t stream = nullptr;
};

struct Pool
{
  std::mutex mtx;  // one staged copy at a time per process
  int device = -1;
  std::vector<Worker> workers;

  void release()
  {
    for (auto& w : workers)
    {
      for (int b = 0; b < NBUF; ++b)
      {
        if (w.buf[b])
          cudaFreeHost(w.buf[b]);
        if (w.ev[b])
          cudaEventDestroy(w.ev[b]);
      }
      if (w.stream)
        cudaStreamDestroy(w.stream);
    }
    workers.clear();
    device = -1;
  }

  void prepare(int dev)
  {
    if (device == dev && !workers.empty())
      return;
    release();
    CUDA_CHECK(cudaSetDevice(dev));
    static const int nenv = getenv("EQLB_COPY_THREADS") ? atoi(getenv("EQLB_COPY_THREADS")) : 0;
    const int hw = (int)std::thread::hardware_concurrency();
    const int nw = nenv > 0 ? std::min(nenv, 32) : std::max(2, std::min(6, hw / 2));  // 4 threads reach 45 GB/s on the B200 hosts
    workers.resize(nw);
    for (auto& w : workers)
    {
      for (int b = 0; b < NBUF; ++b)
      {
        CUDA_CHECK(cudaHostAlloc((void**)&w.buf[b], CHUNK, cudaHostAllocPortable));
        CUDA_CHECK(cudaEventCreateWithFlags(&w.ev[b], cudaEventDisableTiming));
      }
      CUDA_CHECK(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    }
    device = dev;
  }
};

Pool& pool()
{
  static Pool p;  // buffers live until process exit (freed by the driver with the context)
  return p;
}

// to_device: host -> device; else device -> host
void staged(void* dev, void* host, size_t bytes, int device, bool to_device)
{
  Pool& P = pool();
  std::lock_guard<std::mutex> lock(P.mtx);
  P.prepare(device);
  const size_t nchunk = (bytes + CHUNK - 1) / CHUNK;
  const int nw = (int)std::min<size_t>(P.workers.size(), nchunk);
  std::vector<cudaError_t> err(nw, cudaSuccess);
  auto work = [&](int t)
  {
    Worker& w = P.workers[t];
    cudaError_t e = cudaSetDevice(device);
    // chunk c of this worker: c = t, t + nw, ...; buffer b alternates
    int b = 0;
    size_t prev_c = (size_t)-1;
    int prev_b = 0;
    for (size_t c = t; c < nchunk && e == cudaSuccess; c += nw, b ^= 1)
    {
      const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
      if (to_device)
      {
        e = cudaEventSynchronize(w.ev[b]);  // the DMA that last read this buffer is done (no-op on a fresh event)
        if (e != cudaSuccess)
          break;
        std::memcpy(w.buf[b], (const char*)host + off, len);
        e = cudaMemcpyAsync((char*)dev + off, w.buf[b], len, cudaMemcpyHostToDevice, w.stream);
        if (e == cudaSuccess)
          e = cudaEventRecord(w.ev[b], w.stream);
      }
      else
      {
        e = cudaMemcpyAsync(w.buf[b], (const char*)dev + off, len, cudaMemcpyDeviceToHost, w.stream);
        if (e == cudaSuccess)
          e = cudaEventRecord(w.ev[b], w.stream);
        if (prev_c != (size_t)-1 && e == cudaSuccess)
        {
          // drain the previous chunk while this one is in flight
          e = cudaEventSynchronize(w.ev[prev_b]);
          const size_t poff = prev_c * CHUNK, plen = std::min(CHUNK, bytes - poff);
          if (e == cudaSuccess)
            std::memcpy((char*)host + poff, w.buf[prev_b], plen);
        }
        prev_c = c;
        prev_b = b;
      }
    }
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(w.stream);
    if (!to_device && prev_c != (size_t)-1 && e == cudaSuccess)
    {
      const size_t poff = prev_c * CHUNK, plen = std::min(CHUNK, bytes - poff);
      std::memcpy((char*)host + poff, w.buf[prev_b], plen);
    }
    err[t] = e;
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nw; ++t)
    th.emplace_back(work, t);
  work(0);
  for (auto& x : th)
    x.join();
  for (cudaError_t e : err)
    CUDA_CHECK(e);
}

} // namespace

bool host_is_pinned(const void* p)
{
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess)
  {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

// Blocking copies; the device side must not be in use by work queued on other streams (callers synchronise).
void eqlb_h2d(void* dst_dev, const void* src_host, size_t bytes)
{
  if (bytes == 0)
    return;
  static const bool off = getenv("EQLB_STAGED_COPY") && atoi(getenv("EQLB_STAGED_COPY")) == 0;
  if (off || bytes < 2 * CHUNK || host_is_pinned(src_host))
  {
    CUDA_CHECK(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
    return;
  }
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  staged(dst_dev, const_cast<void*>(src_host), bytes, dev, true);
}

void eqlb_d2h(void* dst_host, const void* src_dev, size_t bytes)
{
  if (bytes == 0)
    return;
  static const bool off = getenv("EQLB_STAGED_COPY") && atoi(getenv("EQLB_STAGED_COPY")) == 0;
  if (off || bytes < 2 * CHUNK || host_is_pinned(dst_host))
  {
    CUDA_CHECK(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
    return;
  }
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  staged(const_cast<void*>(src_dev), dst_host, bytes, dev, false);
}

// ---------------------------------------------------------------------------------------------------
// Device allocations from the default stream-ordered memory pool with an unlimited release threshold.
// cudaMalloc / cudaFree of the 50-270 MB blocks of a handle cost 1-2 ms each (20-30 ms per eqlb_create +
// eqlb_set_bcs at 1024^2); blocks freed into the pool are handed out again without a trip to the driver.
// Every allocation is complete (stream synchronised) when the call returns, every free waits for the device
// like cudaFree does, so the buffers can be used on any stream.  EQLB_ASYNC_ALLOC=0: plain cudaMalloc/cudaFree.
// ---------------------------------------------------------------------------------------------------
namespace
{
struct AllocState
{
  bool pooled = false;
  cudaStream_t stream = nullptr;
  AllocState()
  {
    const char* e = getenv("EQLB_ASYNC_ALLOC");
    if (e && atoi(e) == 0)
      return;
    int dev = 0, ok = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&ok, cudaDevAttrMemoryPoolsSupported, dev) != cudaSuccess || !ok)
    {
      cudaGetLastError();
      return;
    }
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess)
    {
      cudaGetLastError();
      return;
    }
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess)
    {
      cudaGetLastError();
      return;
    }
    pooled = true;
  }
};
AllocState& alloc_state()
{
  static AllocState s;
  return s;
}
} // namespace

void* eqlb_dev_alloc(size_t bytes)
{
  void* p = nullptr;
  AllocState& A = alloc_state();
  if (A.pooled)
  {
    CUDA_CHECK(cudaMallocAsync(&p, bytes, A.stream));
    CUDA_CHECK(cudaStreamSynchronize(A.stream));
  }
  else
    CUDA_CHECK(cudaMalloc(&p, bytes));
  return p;
}

void eqlb_dev_free(void* p)
{
  if (!p)
    return;
  AllocState& A = alloc_state();
  if (A.pooled)
  {
    cudaDeviceSynchronize();  // like cudaFree: nothing on any stream still uses the block
    cudaFreeAsync(p, A.stream);
  }
  else
    cudaFree(p);
}
