// capi.cpp -- the C ABI of libhbayes_cuda (include/hbayes_cuda.h): handle management, program cache,
// host<->device staging.  No arithmetic happens here.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hbayes_cuda.h"
#include "engine.hpp"
#include "jit.hpp"
#include "model.hpp"
#include "nccl_dyn.hpp"

using namespace hb;

namespace {

thread_local std::string g_error;

template <class F>
int guarded(F&& f) {
  try {
    f();
    return HB_OK;
  } catch (const Error& e) {
    g_error = e.what();
    return e.code;
  } catch (const std::bad_alloc&) {
    g_error = "host allocation failed";
    return HB_E_OOM;
  } catch (const std::exception& e) {
    g_error = e.what();
    return HB_E_ARG;
  }
}

char* dup_json(const std::string& s) {
  char* p = (char*)malloc(s.size() + 1);
  if (!p) throw std::bad_alloc();
  memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

void cuda_ok(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw Error(e == cudaErrorMemoryAllocation ? HB_E_OOM : HB_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

int pick_device(const int32_t* devices, int32_t n_devices) {
  if (!devices && n_devices == HB_STRUCTURE_ONLY) return -1;  // host structure only: every compute call fails
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    throw Error(HB_E_CUDA, "no CUDA device: libhbayes_cuda has no CPU fallback");
  int dev = 0;
  if (devices && n_devices > 0) dev = devices[0];
  if (dev < 0 || dev >= count) throw Error(HB_E_ARG, "device ordinal out of range");
  for (int i = 1; devices && i < n_devices; ++i) {
    if (devices[i] < 0 || devices[i] >= count) throw Error(HB_E_ARG, "device ordinal out of range");
    for (int j = 0; j < i; ++j)
      if (devices[j] == devices[i]) throw Error(HB_E_ARG, "a device is listed twice");
  }
  return dev;
}

struct Compiled {
  std::unique_ptr<Program> prog;
  Executor* ex = nullptr;
  bool stale = false;  // the network's CPT values changed after `ex` uploaded them (changeFactor)
  bool warm = false;       // loaded from a saved handle that had generated kernels: rebuilt (from the saved cubin) with the executor
  uint64_t last_used = 0;  // Session::tick of the last call that ran it (executors are evicted least recently used first)
  ~Compiled() {
    if (ex) executor_destroy(ex);
  }
};

// What a junction-tree handle and a variable-elimination handle share: the network, the GPU, the cache
// of compiled programs (one per observed-variable list) and grow-only staging buffers.
struct Session {
  Network net;
  int device = 0;
  cudaStream_t stream = nullptr;
  int64_t workspace = 0;
  std::map<std::vector<int>, std::unique_ptr<Compiled>> cache;
  int8_t* d_ev = nullptr;
  double* d_out = nullptr;
  size_t ev_cap = 0, out_cap = 0;
  // single-row calls: pinned host buffers mapped into the device address space (h_* host side, d_* device side)
  void *h_one_ev = nullptr, *h_one_out = nullptr, *d_one_ev = nullptr, *d_one_out = nullptr;
  size_t one_ev_cap = 0, one_out_cap = 0;
  void reserve_mapped(size_t ev_bytes, size_t out_bytes) {
    auto grow = [](void*& h, void*& d, size_t& cap, size_t bytes) {
      if (bytes <= cap) return;
      if (h) cudaFreeHost(h), h = nullptr;
      bytes = std::max<size_t>(bytes * 2, 256);
      cuda_ok(cudaHostAlloc(&h, bytes, cudaHostAllocMapped), "cudaHostAlloc(single-row staging)");
      cuda_ok(cudaHostGetDevicePointer(&d, h, 0), "cudaHostGetDevicePointer");
      cap = bytes;
    };
    grow(h_one_ev, d_one_ev, one_ev_cap, ev_bytes);
    grow(h_one_out, d_one_out, one_out_cap, out_bytes);
  }
  int64_t stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool timing = false;
  Executor* last_ex = nullptr;  // executor of the last batch call (for hb_*_last_timing)
  Executor* last_post_ex = nullptr;  // executor of the last hb_jt_posteriors_batch call (for hb_jt_normalise_batch)
  uint64_t tick = 0, call_start = 0;  // executors used since call_start belong to the running API call: never evicted
  bool fma = false;  // 1e-12 mode for the kernels generated from now on (hb_jt_set_precision)
  // Device-resident programs kept alive per handle (each owns a row arena, index maps, graphs and generated kernels).
  // Mixed observed sets (grouped batches, EM with missing data, mpe) create one per set: beyond this many, the least
  // recently used executors are destroyed (their Programs stay compiled on the host).  HB_MAX_EXECUTORS overrides.
  int max_executors = 16;
  // host-buffer batches are cut into chunks: the device->host copy of chunk i (copy_stream) overlaps the
  // kernels of chunk i+1 (stream)
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> chunk_done;
  cudaEvent_t rest_uploaded = nullptr;
  int64_t chunk_rows = 1 << 18;
  // Pageable caller buffers (a numpy array, a Haskell mallocForeignPtrBytes block): the handle owns a ring of page-locked
  // bounce buffers; the copy engine fills one while host threads drain the others into the caller's matrix (Bounce below).
  // A direct cudaMemcpyAsync to pageable memory is staged by the driver on the calling thread at ~15 GB/s.
  std::vector<void*> ring;
  std::vector<cudaEvent_t> ring_copied;
  size_t ring_bytes = 0;
  void* h_ev_stage = nullptr;  // page-locked copy of a pageable evidence matrix (grow-only)
  size_t ev_stage_cap = 0;
  int copy_thread_share = 1;  // handles that run at the same time on this host (devices of one multi-GPU handle)
  // split clique: variables fixed to this handle's slice; batch results are then left unnormalised
  std::vector<std::pair<int, int>> fixed;
  // split clique with the exchange inside the library (hb_jt_compile_split_nccl): this rank's communicator; programs are
  // compiled with reduce points (CompileOptions::reduce_groups) and batch results come back complete and normalised
  void* comm = nullptr;
  int nccl_rank = 0, nccl_world = 1;
  // grow-only device scratch of hb_jt_em_step (result matrix of a chunk, group staging, evidence, row indices):
  // freeing gigabytes per call costs more than the step itself
  struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
    void* ensure(size_t bytes) {
      if (bytes > cap) {
        if (p) cudaFree(p), p = nullptr, cap = 0;
        cuda_ok(cudaMalloc(&p, bytes), "cudaMalloc(EM scratch)");
        cap = bytes;
      }
      return p;
    }
    ~Scratch() {
      if (p) cudaFree(p);
    }
  } em_mat, em_tmp, em_ev, em_rows;

  ~Session() {
    if (device >= 0) cudaSetDevice(device);
    cache.clear();
    if (comm) {
      try {
        nccl_comm_destroy(comm);
      } catch (...) {
      }
    }
    if (d_ev) cudaFree(d_ev);
    if (d_out) cudaFree(d_out);
    if (h_one_ev) cudaFreeHost(h_one_ev);
    if (h_one_out) cudaFreeHost(h_one_out);
    for (cudaEvent_t e : chunk_done) cudaEventDestroy(e);
    if (rest_uploaded) cudaEventDestroy(rest_uploaded);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    for (void* b : ring) cudaFreeHost(b);
    for (cudaEvent_t e : ring_copied) cudaEventDestroy(e);
    if (h_ev_stage) cudaFreeHost(h_ev_stage);
  }
  // the device-resident form of a compiled program, created on first use
  void evict(Compiled* keep, size_t leave) {  // destroys least recently used executors until at most `leave` others remain
    for (;;) {
      size_t live = 0;
      Compiled* lru = nullptr;
      for (auto& kv : cache) {
        Compiled* x = kv.second.get();
        if (!x->ex || x == keep) continue;
        ++live;
        if (x->last_used > call_start) continue;  // in use by the running call
        if (!lru || x->last_used < lru->last_used) lru = x;
      }
      if (live <= leave || !lru) return;
      if (last_ex == lru->ex) last_ex = nullptr;
      if (last_post_ex == lru->ex) last_post_ex = nullptr;
      cuda_ok(cudaSetDevice(device), "cudaSetDevice");
      cudaDeviceSynchronize();
      executor_destroy(lru->ex);
      lru->ex = nullptr;
    }
  }
  Executor* executor(Compiled& c) {
    if (device < 0) throw Error(HB_E_CUDA, "handle was compiled with HB_STRUCTURE_ONLY: no device, and there is no CPU fallback");
    c.last_used = ++tick;
    if (!c.ex) {
      if (const char* env = getenv("HB_MAX_EXECUTORS")) max_executors = std::max(1, atoi(env));
      evict(&c, (size_t)max_executors - 1);
      try {
        c.ex = executor_create(*c.prog, device, workspace);
      } catch (const Error& e) {
        if (e.code != HB_E_OOM) throw;
        cudaGetLastError();
        evict(&c, 0);  // out of device memory: drop every other device-resident program of this handle and retry once
        c.ex = executor_create(*c.prog, device, workspace);
      }
      executor_set_fma(c.ex, fma);
      executor_set_comm(c.ex, comm);
      c.stale = false;
      if (c.warm) executor_specialise(c.ex);  // the cubin came with the file: no compilation
    }
    if (c.stale) executor_refresh_potentials(c.ex), c.stale = false;
    executor_set_timing(c.ex, timing);
    last_ex = c.ex;
    return c.ex;
  }
  void reserve(size_t ev_bytes, size_t out_bytes) {
    if (device < 0) throw Error(HB_E_CUDA, "handle was compiled with HB_STRUCTURE_ONLY: no device, and there is no CPU fallback");
    cuda_ok(cudaSetDevice(device), "cudaSetDevice");
    if (ev_bytes > ev_cap) {
      if (d_ev) cudaFree(d_ev), d_ev = nullptr;
      cuda_ok(cudaMalloc(&d_ev, ev_bytes), "cudaMalloc(evidence staging)");
      ev_cap = ev_bytes;
    }
    if (out_bytes > out_cap) {
      if (d_out) cudaFree(d_out), d_out = nullptr;
      cuda_ok(cudaMalloc(&d_out, out_bytes), "cudaMalloc(output staging)");
      out_cap = out_bytes;
    }
  }
  void note(const Compiled& c, const RunStats& rs, bool accumulate) {
    if (!accumulate) memset(stats, 0, sizeof stats);
    stats[0] += rs.launches;
    stats[1] += rs.tiles;
    stats[2] = rs.rows_per_tile;
    stats[3] = rs.arena_bytes;
    stats[4] = c.prog->row_entries;
    stats[5] = c.prog->stat_prod;
    stats[6] = c.prog->stat_row_read;
    stats[7] = c.prog->stat_row_write;
  }
};

std::vector<int> observed_of_row(const int8_t* row, int n_vars) {
  std::vector<int> obs;
  for (int v = 0; v < n_vars; ++v)
    if (row[v] >= 0) obs.push_back(v);
  return obs;
}

// true when `p` is ordinary pageable host memory (neither cudaHostAlloc'ed nor cudaHostRegister'ed nor device memory)
bool is_pageable(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();  // older drivers report plain malloc memory as an error
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

// Device -> pageable host memory through the handle's ring of page-locked buffers.  The caller enqueues pieces (at most
// ring_bytes each) in order with `copy`; piece i goes to ring[i % NB] on the copy stream as soon as the host threads have
// drained piece i - NB; W host threads each copy their 1/W of every piece into the destination once its event has fired.
// So the PCIe transfer of a piece overlaps the host copies of the pieces before it, and the host copy runs on W cores.
struct Bounce {
  static constexpr int NB = 4;
  static constexpr size_t PIECE = (size_t)16 << 20;
  Session& s;
  int W = 1;
  struct Piece {
    char* dst;
    size_t bytes;
  };
  std::vector<Piece> pieces;  // reserved up front: workers read entries below `issued` only
  std::atomic<int> issued{0};
  std::atomic<bool> abort{false};
  std::unique_ptr<std::atomic<int>[]> drained;
  std::vector<std::thread> workers;
  int n_issued = 0;

  Bounce(Session& session, size_t max_pieces) : s(session) {
    if (s.ring.empty()) {
      for (int i = 0; i < NB; ++i) {
        void* b = nullptr;
        cuda_ok(cudaHostAlloc(&b, PIECE, cudaHostAllocDefault), "cudaHostAlloc(bounce ring)");
        s.ring.push_back(b);
        cudaEvent_t e;
        cuda_ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
        s.ring_copied.push_back(e);
      }
      s.ring_bytes = PIECE;
    }
    const int hw = (int)std::thread::hardware_concurrency();
    W = std::max(1, std::min(8, hw / 2 / std::max(1, s.copy_thread_share)));
    if (const char* env = getenv("HB_COPY_THREADS")) W = std::max(1, std::min(64, atoi(env)));
    pieces.resize(max_pieces);
    drained.reset(new std::atomic<int>[max_pieces]);
    for (size_t i = 0; i < max_pieces; ++i) drained[i].store(0, std::memory_order_relaxed);
    for (int j = 0; j < W; ++j) workers.emplace_back([this, j] { drain(j); });
  }
  static size_t pieces_for(size_t bytes) { return (bytes + PIECE - 1) / PIECE; }
  void drain(int j) {
    cudaSetDevice(s.device);
    for (size_t i = 0; i < pieces.size(); ++i) {
      while (issued.load(std::memory_order_acquire) <= (int)i) {
        if (abort.load(std::memory_order_relaxed)) return;
        std::this_thread::yield();
      }
      if (cudaEventSynchronize(s.ring_copied[i % NB]) != cudaSuccess) {
        abort.store(true);
        return;
      }
      const Piece& p = pieces[i];
      const size_t a = p.bytes * (size_t)j / (size_t)W / 64 * 64, b = j + 1 == W ? p.bytes : p.bytes * (size_t)(j + 1) / (size_t)W / 64 * 64;
      if (b > a) memcpy(p.dst + a, (const char*)s.ring[i % NB] + a, b - a);
      drained[i].fetch_add(1, std::memory_order_release);
    }
  }
  void wait_drained(int i) {
    while (drained[i].load(std::memory_order_acquire) < W) {
      if (abort.load(std::memory_order_relaxed)) throw Error(HB_E_CUDA, "device-to-host copy failed");
      std::this_thread::yield();
    }
  }
  // enqueue src[0, bytes) -> dst on the copy stream (which already waits for the kernels that produce src)
  void copy(char* dst, const char* src, size_t bytes) {
    for (size_t off = 0; off < bytes; off += PIECE) {
      const size_t m = std::min(PIECE, bytes - off);
      const int i = n_issued;
      if (i >= NB) wait_drained(i - NB);
      cuda_ok(cudaMemcpyAsync(s.ring[i % NB], src + off, m, cudaMemcpyDeviceToHost, s.copy_stream), "cudaMemcpy(out)");
      cuda_ok(cudaEventRecord(s.ring_copied[i % NB], s.copy_stream), "cudaEventRecord");
      pieces[(size_t)i] = Piece{dst + off, m};
      issued.store(++n_issued, std::memory_order_release);
    }
  }
  void finish() {  // every enqueued piece is in the caller's buffer when this returns
    for (int i = std::max(0, n_issued - NB); i < n_issued; ++i) wait_drained(i);
    stop();
  }
  void stop() {
    abort.store(true);
    for (std::thread& t : workers)
      if (t.joinable()) t.join();
    workers.clear();
  }
  ~Bounce() { stop(); }
};

// Runs `get(observed)`'s program over an evidence matrix.  Host buffers: rows are assumed to share the
// observed set of row 0 (the device checks every row); if they do not, they are grouped by observed set
// on the host and each group runs with its own program.
template <class GetProgram>
void run_matrix(Session& s, GetProgram&& get, int64_t n_rows, const int8_t* evidence, double* out, int32_t flags) {
  if (n_rows < 0 || (n_rows > 0 && (!evidence || !out))) throw Error(HB_E_ARG, "null evidence or output buffer");
  if (n_rows == 0) return;
  const int n_vars = s.net.n();
  if (s.device < 0) throw Error(HB_E_CUDA, "handle was compiled with HB_STRUCTURE_ONLY: no device, and there is no CPU fallback");
  cuda_ok(cudaSetDevice(s.device), "cudaSetDevice");
  s.call_start = s.tick;
  RunStats rs;
  if (flags & HB_DEVICE_PTRS) {
    std::vector<int8_t> first(n_vars);
    cuda_ok(cudaMemcpyAsync(first.data(), evidence, n_vars, cudaMemcpyDeviceToHost, s.stream), "cudaMemcpy(first row)");
    cuda_ok(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
    if (s.comm) {  // split clique over NCCL: the rows leave the split variables unobserved; this rank's states go into a copy
      for (auto& f : s.fixed) first[f.first] = (int8_t)f.second;
      Compiled& c = get(observed_of_row(first.data(), n_vars));
      s.reserve((size_t)n_rows * n_vars, 0);
      cuda_ok(cudaMemcpyAsync(s.d_ev, evidence, (size_t)n_rows * n_vars, cudaMemcpyDeviceToDevice, s.stream), "cudaMemcpy(evidence)");
      executor_fill_fixed(s.executor(c), s.d_ev, n_rows, s.stream);
      executor_run(s.executor(c), n_rows, s.d_ev, out, s.stream, &rs, true);
      s.note(c, rs, false);
      executor_check(s.executor(c), s.stream);
      return;
    }
    Compiled& c = get(observed_of_row(first.data(), n_vars));
    executor_run(s.executor(c), n_rows, evidence, out, s.stream, &rs, s.fixed.empty());
    s.note(c, rs, false);
    executor_check(s.executor(c), s.stream);
    return;
  }
  if (s.comm) {  // host buffers, split clique over NCCL: one upload, one pass (every rank must take the same path: no regrouping)
    std::vector<int8_t> first(evidence, evidence + n_vars);
    for (auto& f : s.fixed) first[f.first] = (int8_t)f.second;
    Compiled& c = get(observed_of_row(first.data(), n_vars));
    const int64_t w = c.prog->out_width;
    s.reserve((size_t)n_rows * n_vars, (size_t)n_rows * w * sizeof(double));
    cuda_ok(cudaMemcpyAsync(s.d_ev, evidence, (size_t)n_rows * n_vars, cudaMemcpyHostToDevice, s.stream), "cudaMemcpy(evidence)");
    executor_fill_fixed(s.executor(c), s.d_ev, n_rows, s.stream);
    executor_run(s.executor(c), n_rows, s.d_ev, s.d_out, s.stream, &rs, true);
    s.note(c, rs, false);
    cuda_ok(cudaMemcpyAsync(out, s.d_out, (size_t)n_rows * w * sizeof(double), cudaMemcpyDeviceToHost, s.stream), "cudaMemcpy(out)");
    executor_check(s.executor(c), s.stream);
    return;
  }
  Compiled& c0 = get(observed_of_row(evidence, n_vars));
  const int64_t width = c0.prog->out_width;
  {
    // Zero-copy: when the program runs on the on-chip kernel (one launch, every result row written once, in full 128-byte
    // lines) and both caller buffers are page-locked, the kernel reads the evidence and writes the posteriors straight
    // through the buffers' device mappings -- no staging copy in HBM, the PCIe transfer starts with the first tile and the
    // call takes max(compute, transfer).  Measured on C2 (2^20 rows): 18.3 ms against 18.0 ms for the staged, chunked path
    // below (SM-issued PCIe writes reach ~49 GB/s, the copy engine ~50 in chunks and 57 in one piece), so it is opt-in
    // (HB_ZERO_COPY=1): what it saves is the 0.9 GB staging buffer in HBM, not time.
    const bool zero_copy = getenv("HB_ZERO_COPY") && atoi(getenv("HB_ZERO_COPY")) != 0;
    Executor* ex = s.executor(c0);
    cudaPointerAttributes ae{}, ao{};
    if (zero_copy && executor_on_chip(ex) && cudaPointerGetAttributes(&ae, evidence) == cudaSuccess && cudaPointerGetAttributes(&ao, out) == cudaSuccess &&
        ae.type == cudaMemoryTypeHost && ao.type == cudaMemoryTypeHost && ae.devicePointer && ao.devicePointer) {
      executor_run(ex, n_rows, (const int8_t*)ae.devicePointer, (double*)ao.devicePointer, s.stream, &rs, s.fixed.empty());
      s.note(c0, rs, false);
      try {
        executor_check(ex, s.stream);  // synchronises the stream: the results are in the caller's buffer
        return;
      } catch (const Error& e) {
        if (e.code != HB_E_ARG) throw;  // rows with another observed set: fall through to the grouping path
      }
    }
    cudaGetLastError();  // cudaPointerGetAttributes on plain malloc memory may leave an error state on older drivers
  }
  s.reserve((size_t)n_rows * n_vars, (size_t)n_rows * width * sizeof(double));
  if (!s.copy_stream) cuda_ok(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
  // Chunk schedule: the device->host copy of a chunk overlaps the kernels of the next one, so only the copy of
  // the LAST chunk is exposed, and copying can only start after the FIRST chunk's kernels -- so both ends are
  // small: 10/30/30/20/10 % of a large batch.
  if (const char* env = getenv("HB_CHUNK_ROWS")) s.chunk_rows = std::max<int64_t>(64, atoll(env));
  std::vector<int64_t> cuts;  // chunk end rows
  if (executor_on_chip(s.executor(c0)) && !getenv("HB_CHUNK_ROWS")) {
    // one launch per chunk and a kernel 2.5 x faster than the transfer: small chunks keep the copy engine fed from the start
    // (measured on C2, 2^20 rows: 16.9 ms with 32 K-row chunks, 17.1 ms with 128 K, 18.0 ms with the 10/30/30/20/10 % cut)
    for (int64_t r = 32768; r < n_rows; r += 32768) cuts.push_back(r);
  } else if (n_rows >= 4 * s.chunk_rows / 2 && !getenv("HB_CHUNK_ROWS")) {
    for (double f : {0.1, 0.4, 0.7, 0.9}) cuts.push_back((int64_t)(n_rows * f) / 64 * 64);
  } else {
    const int64_t n_chunks = std::max<int64_t>(1, (n_rows + s.chunk_rows - 1) / s.chunk_rows);
    const int64_t per = ((n_rows + n_chunks - 1) / n_chunks + 63) / 64 * 64;
    for (int64_t r = per; r < n_rows; r += per) cuts.push_back(r);
  }
  cuts.push_back(n_rows);
  Executor* ex0 = s.executor(c0);
  // pageable caller buffers go through page-locked staging owned by the handle (HB_BOUNCE=0: plain cudaMemcpyAsync)
  const bool bounce_on = !(getenv("HB_BOUNCE") && atoi(getenv("HB_BOUNCE")) == 0);
  const size_t out_bytes = (size_t)n_rows * width * sizeof(double), ev_bytes = (size_t)n_rows * n_vars;
  const bool out_bounce = bounce_on && out_bytes >= ((size_t)4 << 20) && is_pageable(out);
  const bool ev_bounce = bounce_on && ev_bytes >= ((size_t)1 << 20) && is_pageable(evidence);
  std::unique_ptr<Bounce> bounce;
  if (out_bounce) {
    size_t np = 0;
    int64_t b0 = 0;
    for (int64_t cut : cuts) np += Bounce::pieces_for((size_t)std::max<int64_t>(0, cut - b0) * width * sizeof(double)), b0 = cut;
    bounce.reset(new Bounce(s, np));
  }
  if (ev_bounce && ev_bytes > s.ev_stage_cap) {
    if (s.h_ev_stage) cudaFreeHost(s.h_ev_stage), s.h_ev_stage = nullptr, s.ev_stage_cap = 0;
    cuda_ok(cudaHostAlloc(&s.h_ev_stage, ev_bytes, cudaHostAllocDefault), "cudaHostAlloc(evidence staging)");
    s.ev_stage_cap = ev_bytes;
  }
  // evidence: the first chunk's rows on the compute stream, the rest on the (still idle) copy stream meanwhile
  const int64_t first_rows = ev_bounce ? n_rows : (cuts[0] > 0 ? cuts[0] : n_rows);
  if (!ev_bounce) cuda_ok(cudaMemcpyAsync(s.d_ev, evidence, (size_t)first_rows * n_vars, cudaMemcpyHostToDevice, s.stream), "cudaMemcpy(evidence)");
  if (first_rows < n_rows) {
    if (!s.rest_uploaded) cuda_ok(cudaEventCreateWithFlags(&s.rest_uploaded, cudaEventDisableTiming), "cudaEventCreate");
    cuda_ok(cudaMemcpyAsync(s.d_ev + first_rows * n_vars, evidence + first_rows * n_vars, (size_t)(n_rows - first_rows) * n_vars,
                            cudaMemcpyHostToDevice, s.copy_stream), "cudaMemcpy(evidence)");
    cuda_ok(cudaEventRecord(s.rest_uploaded, s.copy_stream), "cudaEventRecord");
  }
  bool first_chunk = true;
  int64_t r0 = 0;
  for (size_t c = 0; c < cuts.size(); r0 = cuts[c], ++c) {
    const int64_t m = cuts[c] - r0;
    if (m <= 0) continue;
    if (r0 == first_rows && first_rows < n_rows) cuda_ok(cudaStreamWaitEvent(s.stream, s.rest_uploaded, 0), "cudaStreamWaitEvent");
    if (ev_bounce) {  // this chunk's rows: host copy into the page-locked stage (overlaps the GPU work already queued), then upload
      int8_t* const stage = (int8_t*)s.h_ev_stage + r0 * n_vars;
      memcpy(stage, evidence + r0 * n_vars, (size_t)m * n_vars);
      cuda_ok(cudaMemcpyAsync(s.d_ev + r0 * n_vars, stage, (size_t)m * n_vars, cudaMemcpyHostToDevice, s.stream), "cudaMemcpy(evidence)");
    }
    executor_run(ex0, m, s.d_ev + r0 * n_vars, s.d_out + r0 * width, s.stream, &rs, s.fixed.empty());
    s.note(c0, rs, !first_chunk);
    first_chunk = false;
    if (s.chunk_done.size() <= c) {
      cudaEvent_t e;
      cuda_ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
      s.chunk_done.push_back(e);
    }
    cuda_ok(cudaEventRecord(s.chunk_done[c], s.stream), "cudaEventRecord");
    cuda_ok(cudaStreamWaitEvent(s.copy_stream, s.chunk_done[c], 0), "cudaStreamWaitEvent");
    if (bounce)
      bounce->copy((char*)(out + r0 * width), (const char*)(s.d_out + r0 * width), (size_t)m * width * sizeof(double));
    else
      cuda_ok(cudaMemcpyAsync(out + r0 * width, s.d_out + r0 * width, (size_t)m * width * sizeof(double), cudaMemcpyDeviceToHost, s.copy_stream),
              "cudaMemcpy(out)");
  }
  bool uniform = true;
  try {
    executor_check(ex0, s.stream);
  } catch (const Error& e) {
    if (e.code != HB_E_ARG) throw;
    uniform = false;
  }
  if (bounce) bounce->finish(), bounce.reset();
  cuda_ok(cudaStreamSynchronize(s.copy_stream), "cudaStreamSynchronize");
  if (uniform) return;
  // mixed observed sets: group rows on the host
  std::map<std::vector<int>, std::vector<int64_t>> groups;
  for (int64_t r = 0; r < n_rows; ++r) groups[observed_of_row(evidence + r * n_vars, n_vars)].push_back(r);
  std::vector<int8_t> ev;
  std::vector<double> res;
  bool first = true;
  for (auto& kv : groups) {
    Compiled& c = get(kv.first);
    const int64_t m = (int64_t)kv.second.size();
    ev.resize((size_t)m * n_vars);
    for (int64_t i = 0; i < m; ++i) memcpy(ev.data() + i * n_vars, evidence + kv.second[i] * n_vars, n_vars);
    cuda_ok(cudaMemcpyAsync(s.d_ev, ev.data(), ev.size(), cudaMemcpyHostToDevice, s.stream), "cudaMemcpy(evidence)");
    executor_run(s.executor(c), m, s.d_ev, s.d_out, s.stream, &rs, s.fixed.empty());
    s.note(c, rs, !first);
    first = false;
    executor_check(s.executor(c), s.stream);  // states out of range -> HB_E_ARG
    res.resize((size_t)m * width);
    cuda_ok(cudaMemcpyAsync(res.data(), s.d_out, res.size() * sizeof(double), cudaMemcpyDeviceToHost, s.stream), "cudaMemcpy(out)");
    cuda_ok(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
    for (int64_t i = 0; i < m; ++i) memcpy(out + kv.second[i] * width, res.data() + i * width, width * sizeof(double));
  }
}

}  // namespace

struct hb_net {
  Network net;
};

struct hb_jt {
  Session s;
  std::unique_ptr<Tree> tree;
  // devices[1..] of the compile call: one replica of the handle per further GPU (same network, same tree, own programs
  // and device state).  Host-buffer batch calls shard their rows over self + peers, one host thread per GPU, no
  // communication between them; everything else runs on this handle's own device.
  std::vector<std::unique_ptr<hb_jt>> peers;
  std::vector<int> ev_vars, ev_states;  // current changeEvidence list
  std::vector<double> posts;            // posterior [v] for every v under the current evidence
  bool posts_stale = false;             // a changeFactor happened since `posts` was computed
  std::vector<int64_t> post_off;

  // program for changeEvidence `observed` (in that order) + the given queries (empty = posterior [v] for all v)
  Compiled& program(const std::vector<int>& observed, const std::vector<int>& query) {
    std::vector<int> key = observed;
    key.push_back(-1);
    key.insert(key.end(), query.begin(), query.end());
    key.push_back(-1);
    for (auto& f : s.fixed) key.push_back(f.first), key.push_back(f.second);
    auto it = s.cache.find(key);
    if (it != s.cache.end()) return *it->second;
    std::vector<Query> qs;  // `query`: empty = posterior [v] for every v; else queries separated by -1
    if (query.empty())
      for (int v = 0; v < s.net.n(); ++v) qs.push_back(Query{{v}});
    else {
      qs.push_back(Query{});
      for (int v : query) {
        if (v == -1)
          qs.push_back(Query{});
        else
          qs.back().vars.push_back(v);
      }
    }
    auto c = std::make_unique<Compiled>();
    CompileOptions opt;
    opt.fixed = s.fixed;
    opt.reduce_groups = s.comm != nullptr || s.nccl_world < 0;  // nccl_world < 0: structure-only handle asking for the split program
    c->prog.reset(compile_jt(s.net, *tree, observed, qs, opt));
    return *(s.cache[key] = std::move(c));
  }
  // One evidence row through the program (the stateful single-row API).  Latency path: the row and the result live in
  // pinned host memory mapped into the device's address space, so the call is one launch and one synchronisation -- the
  // kernels read the few evidence bytes and write the few result doubles over PCIe themselves.  The evidence list was
  // validated on the host (check_evidence_list) and the program was compiled for exactly its variables, so the device-side
  // consistency flag cannot be raised here and is not read back.
  void run_one(Compiled& c, std::vector<double>& out) {
    if (!s.fixed.empty()) throw Error(HB_E_UNSUPPORTED, "the single-row calls are not available on a split handle: use the batch call");
    s.call_start = s.tick;
    const size_t n_vars = (size_t)s.net.n(), width = (size_t)c.prog->out_width;
    Executor* ex = s.executor(c);  // sets the device
    s.reserve_mapped(n_vars, width * sizeof(double));
    int8_t* row = (int8_t*)s.h_one_ev;
    memset(row, -1, n_vars);
    for (size_t i = 0; i < ev_vars.size(); ++i) row[ev_vars[i]] = (int8_t)ev_states[i];
    RunStats rs;
    executor_run(ex, 1, (const int8_t*)s.d_one_ev, (double*)s.d_one_out, s.stream, &rs);
    s.note(c, rs, false);
    cuda_ok(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
    out.assign((const double*)s.h_one_out, (const double*)s.h_one_out + width);
  }
};

struct hb_ve {
  Session s;
  std::vector<int> p, r;
  bool mpe = false;  // compiled with hb_ve_compile_mpe: r is maximised instead of kept
  Compiled& program(const std::vector<int>& observed) {
    auto it = s.cache.find(observed);
    if (it != s.cache.end()) return *it->second;
    auto c = std::make_unique<Compiled>();
    c->prog.reset(mpe ? compile_mpe(s.net, p, r, observed) : compile_ve(s.net, p, r, observed));
    return *(s.cache[observed] = std::move(c));
  }
};

static void check_evidence_list(const Network& net, int32_t n, const int32_t* vars, const int32_t* states) {
  if (n < 0 || (n > 0 && (!vars || !states))) throw Error(HB_E_ARG, "null evidence list");
  for (int i = 0; i < n; ++i) {
    if (vars[i] < 0 || vars[i] >= net.n()) throw Error(HB_E_ARG, "evidence on unknown vertex");
    if (states[i] < 0 || states[i] >= net.dims[vars[i]]) throw Error(HB_E_ARG, "evidence state out of range");
  }
}

namespace {
struct DevBuf {  // a device allocation that lives for one call
  void* p = nullptr;
  explicit DevBuf(size_t bytes) { cuda_ok(cudaMalloc(&p, std::max<size_t>(bytes, 8)), "cudaMalloc(EM buffers)"); }
  ~DevBuf() { cudaFree(p); }
  DevBuf(const DevBuf&) = delete;
  template <class T>
  T* as() const { return (T*)p; }
};
}  // namespace

// changeFactor on the handle's private copy of the network (Bayes/Factor.hs:66-71): the CPT whose variable list
// equals `scope` gets the new values; every device-resident program of the handle re-uploads its potentials.
static bool change_factor(Session& s, int32_t n_scope, const int32_t* scope, const double* values) {
  if (n_scope <= 0 || !scope || !values) throw Error(HB_E_ARG, "null or empty factor");
  Network& net = s.net;
  const std::vector<int> sc(scope, scope + n_scope);
  for (int x : sc)
    if (x < 0 || x >= net.n()) throw Error(HB_E_ARG, "factor on unknown vertex");
  int hit = -1;
  for (int v = 0; v < net.n() && hit < 0; ++v)
    if (net.scope[v] == sc) hit = v;
  if (hit < 0) return false;
  if (net.procedural(hit)) throw Error(HB_E_UNSUPPORTED, "changeFactor on a procedural CPT");
  net.values[hit].assign(values, values + net.cpt_entries(hit));
  for (auto& kv : s.cache) kv.second->stale = kv.second->ex != nullptr;  // re-uploaded on next use (Session::executor)
  return true;
}

extern "C" {

const char* hb_last_error(void) { return g_error.c_str(); }
void hb_string_free(char* s) { free(s); }
int hb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int hb_net_create(int32_t n_vars, const int32_t* dims, const int64_t* cpt_off, const int32_t* cpt_scope,
                  const int64_t* val_off, const double* cpt_vals, hb_net** out) {
  return guarded([&] {
    if (!out) throw Error(HB_E_ARG, "null out");
    std::unique_ptr<Network> n(network_create(n_vars, dims, cpt_off, cpt_scope, val_off, cpt_vals));
    *out = new hb_net{std::move(*n)};
  });
}
int hb_net_create_procedural(int32_t n_vars, const int32_t* dims, const int64_t* cpt_off, const int32_t* cpt_scope,
                             const int64_t* val_off, const double* cpt_vals, const uint64_t* proc_seed, hb_net** out) {
  return guarded([&] {
    if (!out || !proc_seed) throw Error(HB_E_ARG, "null argument");
    std::unique_ptr<Network> n(network_create(n_vars, dims, cpt_off, cpt_scope, val_off, cpt_vals, proc_seed));
    *out = new hb_net{std::move(*n)};
  });
}
int hb_net_cpt_values(const hb_net* net, int32_t v, int64_t offset, int64_t count, double* out) {
  return guarded([&] {
    if (!net || !out || v < 0 || v >= net->net.n() || offset < 0 || count < 0 || offset + count > net->net.cpt_entries(v))
      throw Error(HB_E_ARG, "bad CPT range");
    for (int64_t i = 0; i < count; ++i) out[i] = net->net.cpt_value(v, offset + i);
  });
}
int hb_net_load_hugin(const char* path, hb_net** out) {
  return guarded([&] {
    if (!out) throw Error(HB_E_ARG, "null out");
    std::unique_ptr<Network> n(network_load_hugin(path));
    *out = new hb_net{std::move(*n)};
  });
}
int hb_net_n_vars(const hb_net* net) { return net ? net->net.n() : HB_E_ARG; }
int hb_net_dims(const hb_net* net, int32_t* dims_out) {
  return guarded([&] {
    if (!net || !dims_out) throw Error(HB_E_ARG, "null argument");
    for (int v = 0; v < net->net.n(); ++v) dims_out[v] = net->net.dims[v];
  });
}
int hb_net_describe_json(const hb_net* net, char** out_json) {
  return guarded([&] {
    if (!net || !out_json) throw Error(HB_E_ARG, "null argument");
    *out_json = dup_json(describe_network(net->net));
  });
}
void hb_net_free(hb_net* net) { delete net; }

int hb_ve_order(const hb_net* net, int32_t metric, int32_t* order_out, int32_t* n_out) {
  return guarded([&] {
    if (!net || !order_out || !n_out) throw Error(HB_E_ARG, "null argument");
    std::vector<int> o = ve_elimination_order(net->net, metric);
    for (size_t i = 0; i < o.size(); ++i) order_out[i] = o[i];
    *n_out = (int32_t)o.size();
  });
}

int hb_ve_degree_order(const hb_net* net, const int32_t* order, int32_t n, int32_t* width) {
  return guarded([&] {
    if (!net || !width || n < 0 || (n > 0 && !order)) throw Error(HB_E_ARG, "null argument");
    *width = ve_degree_order(net->net, std::vector<int>(order, order + n));
  });
}

static int jt_compile(const hb_net* net, int32_t cmp, const int32_t* order, const int32_t* devices, int32_t n_devices, hb_jt** out) {
  return guarded([&] {
    if (!net || !out) throw Error(HB_E_ARG, "null argument");
    auto jt = std::make_unique<hb_jt>();
    jt->s.net = net->net;
    jt->tree.reset(tree_build(jt->s.net, cmp, order));  // host work first: structure errors surface without a GPU
    jt->s.device = pick_device(devices, n_devices);
    jt->post_off.assign(1, 0);
    for (int d : jt->s.net.dims) jt->post_off.push_back(jt->post_off.back() + d);
    Compiled& prior = jt->program({}, {});
    if (jt->s.device >= 0) jt->run_one(prior, jt->posts);  // createJunctionTree calibrates the tree (distribute . collect)
    for (int i = 1; jt->s.device >= 0 && devices && i < n_devices; ++i) {
      auto peer = std::make_unique<hb_jt>();
      peer->s.net = jt->s.net;
      peer->tree = std::make_unique<Tree>(*jt->tree);
      peer->s.device = devices[i];
      peer->post_off = jt->post_off;
      jt->peers.push_back(std::move(peer));
    }
    jt->s.copy_thread_share = (int)(1 + jt->peers.size());
    for (auto& peer : jt->peers) peer->s.copy_thread_share = jt->s.copy_thread_share;
    *out = jt.release();
  });
}
int hb_jt_compile(const hb_net* net, int32_t cmp, const int32_t* devices, int32_t n_devices, hb_jt** out) {
  return jt_compile(net, cmp, nullptr, devices, n_devices, out);
}
int hb_jt_compile_split(const hb_net* net, int32_t cmp, const int32_t* devices, int32_t n_devices, int32_t n_fixed,
                        const int32_t* vars, const int32_t* states, hb_jt** out) {
  return guarded([&] {
    if (!net || !out) throw Error(HB_E_ARG, "null argument");
    check_evidence_list(net->net, n_fixed, vars, states);
    auto jt = std::make_unique<hb_jt>();
    jt->s.net = net->net;
    jt->tree.reset(tree_build(jt->s.net, cmp, nullptr));
    jt->s.device = pick_device(devices, n_devices);
    jt->post_off.assign(1, 0);
    for (int d : jt->s.net.dims) jt->post_off.push_back(jt->post_off.back() + d);
    for (int i = 0; i < n_fixed; ++i) jt->s.fixed.push_back({vars[i], states[i]});
    *out = jt.release();  // no calibration: the unsplit tables may not fit one GPU
  });
}
int hb_nccl_unique_id(void* out128) {
  return guarded([&] {
    if (!out128) throw Error(HB_E_ARG, "null out");
    nccl_unique_id(out128);
  });
}
int hb_jt_compile_split_nccl(const hb_net* net, int32_t cmp, int32_t device, int32_t n_fixed, const int32_t* vars, const int32_t* states,
                             int32_t rank, int32_t world, const void* unique_id128, hb_jt** out) {
  return guarded([&] {
    if (!net || !out) throw Error(HB_E_ARG, "null argument");
    check_evidence_list(net->net, n_fixed, vars, states);
    if (world < 1 || rank < 0 || rank >= world) throw Error(HB_E_ARG, "rank outside the world");
    auto jt = std::make_unique<hb_jt>();
    jt->s.net = net->net;
    jt->tree.reset(tree_build(jt->s.net, cmp, nullptr));
    jt->post_off.assign(1, 0);
    for (int d : jt->s.net.dims) jt->post_off.push_back(jt->post_off.back() + d);
    for (int i = 0; i < n_fixed; ++i) jt->s.fixed.push_back({vars[i], states[i]});
    jt->s.nccl_rank = rank;
    if (device == HB_STRUCTURE_ONLY) {  // parity hook: the split program (reduce points) without a GPU or a communicator
      jt->s.device = -1;
      jt->s.nccl_world = -world;
      *out = jt.release();
      return;
    }
    if (!unique_id128) throw Error(HB_E_ARG, "null NCCL unique id");
    int32_t dev_list[1] = {device};
    jt->s.device = pick_device(dev_list, 1);
    jt->s.nccl_world = world;
    cuda_ok(cudaSetDevice(jt->s.device), "cudaSetDevice");
    jt->s.comm = nccl_comm_init(world, rank, unique_id128);  // collective: every rank of the world is in this call now
    *out = jt.release();
  });
}
int hb_jt_compile_with_order(const hb_net* net, const int32_t* elim_order, const int32_t* devices, int32_t n_devices, hb_jt** out) {
  if (!elim_order) {
    g_error = "null elimination order";
    return HB_E_ARG;
  }
  return jt_compile(net, 0, elim_order, devices, n_devices, out);
}
// ---- persistence: network + junction-tree structure in one little-endian binary file -----------------------
namespace {
const char JT_MAGIC[8] = {'H', 'B', 'J', 'T', 'r', '0', '2', '\0'};     // r02 = r01 + compiled programs and their cubins
const char JT_MAGIC_R01[8] = {'H', 'B', 'J', 'T', 'r', '0', '1', '\0'};

struct Writer {
  FILE* f;
  void i64(int64_t v) {
    if (fwrite(&v, sizeof v, 1, f) != 1) throw Error(HB_E_IO, "write failed");
  }
  void ints(const std::vector<int>& v) {
    i64((int64_t)v.size());
    for (int x : v) i64(x);
  }
  void lists(const std::vector<std::vector<int>>& v) {
    i64((int64_t)v.size());
    for (auto& x : v) ints(x);
  }
};
struct Reader {
  FILE* f;
  int64_t size = 0;  // of the file: no count may promise more data than the file still holds
  explicit Reader(FILE* file) : f(file) {
    const long at = ftell(f);
    fseek(f, 0, SEEK_END);
    size = ftell(f);
    fseek(f, at, SEEK_SET);
  }
  int64_t i64() {
    int64_t v;
    if (fread(&v, sizeof v, 1, f) != 1) throw Error(HB_E_IO, "truncated file");
    return v;
  }
  int64_t count(int64_t limit) {
    const int64_t n = i64();
    if (n < 0 || n > limit || n > size - (int64_t)ftell(f)) throw Error(HB_E_IO, "corrupt file");
    return n;
  }
  std::vector<int> ints() {
    std::vector<int> v((size_t)count(1 << 28));
    for (int& x : v) x = (int)i64();
    return v;
  }
  std::vector<std::vector<int>> lists() {
    std::vector<std::vector<int>> v((size_t)count(1 << 24));
    for (auto& x : v) x = ints();
    return v;
  }
};
}  // namespace

// A loaded tree is used as raw indices by the schedule compiler: every id is range-checked and the structure must be the
// one createJunctionTree would hand over (clusters sorted, factors covering every vertex once inside their cluster,
// separators joining a parent and a child whose intersection they are, one root, no cycle).
static void validate_tree(const Network& net, const Tree& t) {
  auto bad = [](const char* what) { throw Error(HB_E_IO, std::string("corrupt file: ") + what); };
  const int n = net.n(), nc = (int)t.clusters.size(), nsep = t.n_seps();
  if (nc < 1 || nsep != nc - 1) bad("cluster / separator count");
  if ((int)t.parent_sep.size() != nc || (int)t.children.size() != nc || (int)t.factors.size() != nc || (int)t.by_rank.size() != nc ||
      (int)t.sep_child.size() != nsep || (int)t.sep_vars.size() != nsep || t.root < 0 || t.root >= nc)
    bad("table sizes");
  auto sorted_ids = [&](const std::vector<int>& v) {
    for (size_t i = 0; i < v.size(); ++i)
      if (v[i] < 0 || v[i] >= n || (i && v[i] <= v[i - 1])) return false;
    return true;
  };
  if ((int)t.elim_order.size() != n) bad("elimination order");
  {
    std::vector<char> seen(n, 0);
    for (int v : t.elim_order) {
      if (v < 0 || v >= n || seen[v]) bad("elimination order");
      seen[v] = 1;
    }
  }
  for (const auto& c : t.clusters)
    if (c.empty() || !sorted_ids(c)) bad("cluster");
  {
    std::vector<char> seen(nc, 0);
    for (int c : t.by_rank) {
      if (c < 0 || c >= nc || seen[c]) bad("cluster rank");
      seen[c] = 1;
    }
    for (int i = 1; i < nc; ++i)
      if (!(t.clusters[t.by_rank[i - 1]] < t.clusters[t.by_rank[i]])) bad("cluster rank order");
  }
  auto subset = [](const std::vector<int>& a, const std::vector<int>& b) { return std::includes(b.begin(), b.end(), a.begin(), a.end()); };
  std::vector<char> has_factor(n, 0);
  for (int c = 0; c < nc; ++c)
    for (int v : t.factors[c]) {
      if (v < 0 || v >= n || has_factor[v]) bad("factor assignment");
      has_factor[v] = 1;
      std::vector<int> fam = net.scope[v];
      std::sort(fam.begin(), fam.end());
      if (!subset(fam, t.clusters[c])) bad("factor outside its cluster");
    }
  for (int v = 0; v < n; ++v)
    if (!has_factor[v]) bad("vertex without a cluster");
  if (t.parent_sep[t.root] != -1) bad("root");
  std::vector<int> child_of(nsep, -1);
  for (int c = 0; c < nc; ++c) {
    const int ps = t.parent_sep[c];
    if (c != t.root) {
      if (ps < 1 || ps > nsep || t.sep_child[ps - 1] != c) bad("parent separator");
    }
    for (int sp : t.children[c]) {
      if (sp < 1 || sp > nsep || t.sep_parent[sp - 1] != c || child_of[sp - 1] >= 0) bad("children list");
      child_of[sp - 1] = c;
    }
  }
  for (int sp = 0; sp < nsep; ++sp) {
    const int a = t.sep_parent[sp], b = t.sep_child[sp];
    if (a < 0 || a >= nc || b < 0 || b >= nc || a == b || child_of[sp] != a || t.parent_sep[b] != sp + 1) bad("separator ends");
    if (!sorted_ids(t.sep_vars[sp]) || !subset(t.sep_vars[sp], t.clusters[a]) || !subset(t.sep_vars[sp], t.clusters[b])) bad("separator variables");
  }
  for (int c = 0; c < nc; ++c) {  // every cluster reaches the root: no cycle
    int at = c, hops = 0;
    while (at != t.root) {
      at = t.sep_parent[t.parent_sep[at] - 1];
      if (++hops > nc) bad("cycle in the tree");
    }
  }
}

int hb_jt_save(const hb_jt* jt, const char* path) {
  return guarded([&] {
    if (!jt || !path) throw Error(HB_E_ARG, "null argument");
    FILE* f = fopen(path, "wb");
    if (!f) throw Error(HB_E_IO, std::string("cannot create ") + path);
    std::unique_ptr<FILE, int (*)(FILE*)> guard(f, fclose);
    if (fwrite(JT_MAGIC, 1, 8, f) != 8) throw Error(HB_E_IO, "write failed");
    Writer w{f};
    const Network& net = jt->s.net;
    const Tree& t = *jt->tree;
    w.ints(net.dims);
    w.lists(net.scope);
    for (int v = 0; v < net.n(); ++v) {
      w.i64(net.procedural(v) ? (int64_t)net.proc_seed[v] : 0);
      w.i64((int64_t)net.values[v].size());
      if (!net.values[v].empty() && fwrite(net.values[v].data(), sizeof(double), net.values[v].size(), f) != net.values[v].size())
        throw Error(HB_E_IO, "write failed");
      w.i64((int64_t)net.names[v].size());
      if (fwrite(net.names[v].data(), 1, net.names[v].size(), f) != net.names[v].size()) throw Error(HB_E_IO, "write failed");
    }
    w.ints(t.elim_order);
    w.lists(t.clusters);
    w.i64(t.root);
    w.ints(t.parent_sep);
    w.ints(t.sep_parent);
    w.ints(t.sep_child);
    w.lists(t.sep_vars);
    w.lists(t.children);
    w.lists(t.factors);
    w.ints(t.by_rank);
    // compiled programs: the (observed list, queries) each was compiled for -- the step list and its index maps are a
    // deterministic function of network + tree + this key and are rebuilt in milliseconds -- and the cubin of the
    // kernels generated for it (seconds of NVRTC), keyed by the hash of their source
    std::vector<int> fixed;
    for (auto& f : jt->s.fixed) fixed.push_back(f.first), fixed.push_back(f.second);
    w.ints(fixed);
    std::vector<std::pair<std::string, std::shared_ptr<const std::string>>> cubins;
    w.i64((int64_t)jt->s.cache.size());
    for (auto& kv : jt->s.cache) {
      w.ints(kv.first);
      std::string key = kv.second->ex ? executor_jit_key(kv.second->ex) : "";
      std::shared_ptr<const std::string> cubin = key.empty() ? nullptr : jit_registry_get(key);
      if (!cubin) key.clear();
      w.i64((int64_t)key.size());
      if (!key.empty() && fwrite(key.data(), 1, key.size(), f) != key.size()) throw Error(HB_E_IO, "write failed");
      if (cubin) cubins.push_back({key, cubin});
    }
    w.i64((int64_t)cubins.size());
    for (auto& c : cubins) {
      w.i64((int64_t)c.first.size());
      w.i64((int64_t)c.second->size());
      if (fwrite(c.first.data(), 1, c.first.size(), f) != c.first.size() || fwrite(c.second->data(), 1, c.second->size(), f) != c.second->size())
        throw Error(HB_E_IO, "write failed");
    }
  });
}

int hb_jt_load(const char* path, const int32_t* devices, int32_t n_devices, hb_jt** out) {
  return guarded([&] {
    if (!path || !out) throw Error(HB_E_ARG, "null argument");
    FILE* f = fopen(path, "rb");
    if (!f) throw Error(HB_E_IO, std::string("cannot open ") + path);
    std::unique_ptr<FILE, int (*)(FILE*)> guard(f, fclose);
    char magic[8];
    if (fread(magic, 1, 8, f) != 8 || (memcmp(magic, JT_MAGIC, 8) != 0 && memcmp(magic, JT_MAGIC_R01, 8) != 0))
      throw Error(HB_E_IO, "not a libhbayes_cuda junction-tree file");
    const bool has_programs = memcmp(magic, JT_MAGIC, 8) == 0;
    Reader r(f);
    auto jt = std::make_unique<hb_jt>();
    Network& net = jt->s.net;
    net.dims = r.ints();
    net.scope = r.lists();
    const int n = net.n();
    if ((int)net.scope.size() != n) throw Error(HB_E_IO, "corrupt file");
    net.values.resize(n);
    net.proc_seed.assign(n, 0);
    for (int v = 0; v < n; ++v) {
      net.proc_seed[v] = (uint64_t)r.i64();
      net.values[v].resize((size_t)r.count(((int64_t)1 << 34) / 8));
      if (!net.values[v].empty() && fread(net.values[v].data(), sizeof(double), net.values[v].size(), f) != net.values[v].size())
        throw Error(HB_E_IO, "truncated file");
      std::string name((size_t)r.count(1 << 16), ' ');
      if (!name.empty() && fread(&name[0], 1, name.size(), f) != name.size()) throw Error(HB_E_IO, "truncated file");
      net.names.push_back(name);
    }
    // the same consistency checks as hb_net_create
    {
      std::vector<int32_t> dims(net.dims.begin(), net.dims.end()), scopes;
      std::vector<int64_t> coff(1, 0), voff(1, 0);
      std::vector<double> vals(1, 0.0);
      for (int v = 0; v < n; ++v) {
        scopes.insert(scopes.end(), net.scope[v].begin(), net.scope[v].end());
        coff.push_back((int64_t)scopes.size());
        voff.push_back(voff.back() + (int64_t)net.values[v].size());
      }
      std::vector<double> flat;
      for (int v = 0; v < n; ++v) flat.insert(flat.end(), net.values[v].begin(), net.values[v].end());
      flat.push_back(0.0);
      std::unique_ptr<Network> checked(network_create(n, dims.data(), coff.data(), scopes.data(), voff.data(), flat.data(), net.proc_seed.data()));
    }
    auto t = std::make_unique<Tree>();
    t->elim_order = r.ints();
    t->clusters = r.lists();
    t->root = (int)r.i64();
    t->parent_sep = r.ints();
    t->sep_parent = r.ints();
    t->sep_child = r.ints();
    t->sep_vars = r.lists();
    t->children = r.lists();
    t->factors = r.lists();
    t->by_rank = r.ints();
    const size_t nc = t->clusters.size();
    if (t->parent_sep.size() != nc || t->children.size() != nc || t->factors.size() != nc || t->by_rank.size() != nc ||
        t->sep_child.size() != t->sep_parent.size() || t->sep_vars.size() != t->sep_parent.size() || t->root < 0 || (size_t)t->root >= nc)
      throw Error(HB_E_IO, "corrupt file");
    validate_tree(net, *t);
    jt->tree = std::move(t);
    jt->s.device = pick_device(devices, n_devices);
    jt->post_off.assign(1, 0);
    for (int d : net.dims) jt->post_off.push_back(jt->post_off.back() + d);
    if (has_programs) {
      const std::vector<int> fixed = r.ints();
      if (fixed.size() % 2) throw Error(HB_E_IO, "corrupt file: split list");
      for (size_t i = 0; i < fixed.size(); i += 2) {
        if (fixed[i] < 0 || fixed[i] >= n || fixed[i + 1] < 0 || fixed[i + 1] >= net.dims[fixed[i]]) throw Error(HB_E_IO, "corrupt file: split list");
        jt->s.fixed.push_back({fixed[i], fixed[i + 1]});
      }
      const int64_t n_prog = r.count(1 << 20);
      std::vector<std::pair<std::vector<int>, std::string>> programs;
      for (int64_t i = 0; i < n_prog; ++i) {
        std::vector<int> key = r.ints();
        std::string ck((size_t)r.count(64), ' ');
        if (!ck.empty() && fread(&ck[0], 1, ck.size(), f) != ck.size()) throw Error(HB_E_IO, "truncated file");
        programs.push_back({std::move(key), std::move(ck)});
      }
      const int64_t n_cubin = r.count(1 << 20);
      for (int64_t i = 0; i < n_cubin; ++i) {
        std::string ck((size_t)r.count(64), ' ');
        std::string blob((size_t)r.count((int64_t)1 << 30), '\0');
        if (fread(&ck[0], 1, ck.size(), f) != ck.size() || fread(&blob[0], 1, blob.size(), f) != blob.size()) throw Error(HB_E_IO, "truncated file");
        // a cubin is only ever used under the key of the source this process generates itself: a damaged or foreign blob
        // under a wrong key is never looked up, one under the right key fails cuModuleLoadData and is ignored
        if (blob.size() > 64 && blob.compare(0, 4, "\x7f" "ELF") == 0) jit_registry_put(ck, blob);
      }
      for (auto& pr : programs) {  // key = observed..., -1, queries (separated by -1)..., -1, fixed pairs
        const std::vector<int>& key = pr.first;
        if (key.size() < 2 + fixed.size()) throw Error(HB_E_IO, "corrupt file: program key");
        const size_t q_end = key.size() - fixed.size() - 1;
        size_t o_end = 0;
        while (o_end < q_end && key[o_end] != -1) ++o_end;
        if (o_end >= q_end || key[q_end] != -1) throw Error(HB_E_IO, "corrupt file: program key");
        const std::vector<int> observed(key.begin(), key.begin() + o_end), query(key.begin() + o_end + 1, key.begin() + q_end);
        try {
          Compiled& c = jt->program(observed, query);
          c.warm = !pr.second.empty();
        } catch (const Error& e) {
          throw Error(HB_E_IO, std::string("corrupt file: saved program does not compile: ") + e.what());
        }
      }
    }
    if (jt->s.fixed.empty()) {
      Compiled& prior = jt->program({}, {});
      if (jt->s.device >= 0) jt->run_one(prior, jt->posts);
    }
    *out = jt.release();
  });
}

int hb_jt_describe_json(const hb_jt* jt, char** out_json) {
  return guarded([&] {
    if (!jt || !out_json) throw Error(HB_E_ARG, "null argument");
    *out_json = dup_json(describe_tree(*jt->tree));
  });
}
int hb_jt_program_json(hb_jt* jt, int32_t n, const int32_t* vars, char** out_json) {
  return guarded([&] {
    if (!jt || !out_json || n < 0 || (n > 0 && !vars)) throw Error(HB_E_ARG, "null argument");
    *out_json = dup_json(describe_program(*jt->program(std::vector<int>(vars, vars + n), {}).prog));
  });
}
int hb_jt_jit_source(hb_jt* jt, int32_t n_observed, const int32_t* observed, char** out_source) {
  return guarded([&] {
    if (!jt || !out_source || n_observed < 0 || (n_observed > 0 && !observed)) throw Error(HB_E_ARG, "null argument");
    const Program& P = *jt->program(std::vector<int>(observed, observed + n_observed), {}).prog;
    std::vector<int> steps;
    for (size_t si = 0; si < P.steps.size(); ++si)
      if (jit_eligible(P, P.steps[si])) steps.push_back((int)si);
    *out_source = dup_json(jit_source(P, steps));
  });
}
int hb_jt_tree_source(hb_jt* jt, int32_t n_observed, const int32_t* observed, char** out_source) {
  return guarded([&] {
    if (!jt || !out_source || n_observed < 0 || (n_observed > 0 && !observed)) throw Error(HB_E_ARG, "null argument");
    const Program& P = *jt->program(std::vector<int>(observed, observed + n_observed), {}).prog;
    TreePlan plan;
    std::string why;
    if (!tree_plan(P, plan, &why)) throw Error(HB_E_UNSUPPORTED, "no on-chip kernel for this program: " + why);
    *out_source = dup_json(tree_source(P, plan));
  });
}
int hb_jt_set_evidence(hb_jt* jt, int32_t n, const int32_t* vars, const int32_t* states) {
  return guarded([&] {
    if (!jt) throw Error(HB_E_ARG, "null handle");
    check_evidence_list(jt->s.net, n, vars, states);
    std::vector<int> v(vars, vars + n), st(states, states + n);
    Compiled& c = jt->program(v, {});
    jt->ev_vars = v;
    jt->ev_states = st;
    jt->run_one(c, jt->posts);
    jt->posts_stale = false;
  });
}
int hb_jt_posterior(hb_jt* jt, int32_t n_q, const int32_t* q_vars, int32_t* out_scope, double* out, int64_t out_len,
                    int64_t* n_written) {
  return guarded([&] {
    if (!jt || n_q <= 0 || !q_vars || !out) throw Error(HB_E_ARG, "null or empty query");
    const Network& net = jt->s.net;
    if (n_q == 1) {
      const int v = q_vars[0];
      if (v < 0 || v >= net.n()) throw Error(HB_E_ARG, "query on unknown vertex");
      if (jt->s.device < 0) throw Error(HB_E_CUDA, "handle was compiled with HB_STRUCTURE_ONLY: no device, and there is no CPU fallback");
      if (!jt->s.fixed.empty()) throw Error(HB_E_UNSUPPORTED, "the single-row calls are not available on a split handle: use the batch call");
      // a handle that was never calibrated (or whose CPTs changed) answers after a calibration under the current evidence
      if (jt->posts_stale || (int64_t)jt->posts.size() != jt->post_off.back()) jt->run_one(jt->program(jt->ev_vars, {}), jt->posts), jt->posts_stale = false;
      if (out_len < net.dims[v]) throw Error(HB_E_ARG, "output buffer too small");
      for (int i = 0; i < net.dims[v]; ++i) out[i] = jt->posts[jt->post_off[v] + i];
      if (out_scope) out_scope[0] = v;
      if (n_written) *n_written = net.dims[v];
      return;
    }
    Compiled& c = jt->program(jt->ev_vars, std::vector<int>(q_vars, q_vars + n_q));
    if (out_len < c.prog->out_width) throw Error(HB_E_ARG, "output buffer too small");
    std::vector<double> res;
    jt->run_one(c, res);
    std::copy(res.begin(), res.end(), out);
    const Group& g = c.prog->groups.back();
    if (out_scope)
      for (size_t i = 0; i < g.out_scope.size(); ++i) out_scope[i] = g.out_scope[i];
    if (n_written) *n_written = c.prog->out_width;
  });
}
// rows below which a host-buffer batch is not worth spreading over the handle's other GPUs
static constexpr int64_t kMinRowsPerDevice = 4096;

int hb_jt_posteriors_batch(hb_jt* jt, int64_t n_rows, const int8_t* evidence, double* out, int32_t flags) {
  return guarded([&] {
    if (!jt) throw Error(HB_E_ARG, "null handle");
    const int64_t width = jt->s.net.width();
    const int n_vars = jt->s.net.n();
    const size_t n_dev = 1 + jt->peers.size();
    if (n_dev > 1 && !(flags & HB_DEVICE_PTRS) && n_rows >= (int64_t)n_dev * kMinRowsPerDevice && evidence && out) {
      // evidence rows are independent: contiguous slices, one host thread and one GPU each, no communication
      std::vector<int> codes(n_dev, HB_OK);
      std::vector<std::string> messages(n_dev);
      std::vector<std::thread> threads;
      for (size_t d = 0; d < n_dev; ++d) {
        const int64_t lo = n_rows * (int64_t)d / (int64_t)n_dev, hi = n_rows * (int64_t)(d + 1) / (int64_t)n_dev;
        hb_jt* h = d == 0 ? jt : jt->peers[d - 1].get();
        threads.emplace_back([=, &codes, &messages] {
          codes[d] = guarded([&] {
            run_matrix(h->s, [&](const std::vector<int>& obs) -> Compiled& { return h->program(obs, {}); }, hi - lo, evidence + lo * n_vars,
                       out + lo * width, flags);
            h->s.last_post_ex = h->s.last_ex;
          });
          if (codes[d] != HB_OK) messages[d] = g_error;  // thread-local: carried back to the caller's thread
        });
      }
      for (std::thread& t : threads) t.join();
      for (size_t d = 0; d < n_dev; ++d)
        if (codes[d] != HB_OK) throw Error(codes[d], "device " + std::to_string(d == 0 ? jt->s.device : jt->peers[d - 1]->s.device) + ": " + messages[d]);
      return;
    }
    run_matrix(jt->s, [&](const std::vector<int>& obs) -> Compiled& { return jt->program(obs, {}); }, n_rows, evidence, out, flags);
    jt->s.last_post_ex = jt->s.last_ex;  // every program of this call has the all-posteriors layout
  });
}
int hb_jt_n_devices(const hb_jt* jt) { return jt ? (int)(1 + jt->peers.size()) : HB_E_ARG; }

int hb_host_alloc(int64_t bytes, void** out) {
  return guarded([&] {
    if (!out || bytes < 0) throw Error(HB_E_ARG, "null out or negative size");
    void* p = nullptr;
    cuda_ok(cudaHostAlloc(&p, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocPortable), "cudaHostAlloc");
    *out = p;
  });
}
void hb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
int hb_host_register(void* p, int64_t bytes) {
  return guarded([&] {
    if (!p || bytes <= 0) throw Error(HB_E_ARG, "null buffer");
    cuda_ok(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable), "cudaHostRegister");
  });
}
int hb_host_unregister(void* p) {
  return guarded([&] {
    if (!p) throw Error(HB_E_ARG, "null buffer");
    cuda_ok(cudaHostUnregister(p), "cudaHostUnregister");
  });
}
int hb_jt_queries_batch(hb_jt* jt, int32_t n_queries, const int32_t* query_off, const int32_t* query_vars, int64_t n_rows,
                        const int8_t* evidence, double* out, int32_t flags, int32_t* out_scopes) {
  return guarded([&] {
    if (!jt || n_queries <= 0 || !query_off || !query_vars) throw Error(HB_E_ARG, "null or empty query list");
    std::vector<int> flat;
    for (int q = 0; q < n_queries; ++q) {
      if (query_off[q + 1] <= query_off[q]) throw Error(HB_E_ARG, "empty query");
      if (q) flat.push_back(-1);
      flat.insert(flat.end(), query_vars + query_off[q], query_vars + query_off[q + 1]);
    }
    const Program* last = nullptr;
    run_matrix(jt->s, [&](const std::vector<int>& obs) -> Compiled& {
      Compiled& c = jt->program(obs, flat);
      last = c.prog.get();
      return c;
    }, n_rows, evidence, out, flags);
    if (out_scopes && last) {  // result variable order of every query (rule A.5 / A.10), concatenated like query_vars
      int at = 0;
      for (const Group& g : last->groups)
        if (g.kind == 2)
          for (int v : g.out_scope) out_scopes[at++] = v;
    }
  });
}
int hb_jt_queries_width(hb_jt* jt, int32_t n_queries, const int32_t* query_off, const int32_t* query_vars, int64_t* width) {
  return guarded([&] {
    if (!jt || n_queries <= 0 || !query_off || !query_vars || !width) throw Error(HB_E_ARG, "null or empty query list");
    int64_t w = 0;
    for (int q = 0; q < n_queries; ++q) {
      int64_t p = 1;
      for (int i = query_off[q]; i < query_off[q + 1]; ++i) {
        if (query_vars[i] < 0 || query_vars[i] >= jt->s.net.n()) throw Error(HB_E_ARG, "query on unknown vertex");
        p *= jt->s.net.dims[query_vars[i]];
      }
      w += p;
    }
    *width = w;
  });
}
int hb_jt_change_factor(hb_jt* jt, int32_t n_scope, const int32_t* scope, const double* values, int32_t* replaced) {
  return guarded([&] {
    if (!jt) throw Error(HB_E_ARG, "null handle");
    const bool hit = change_factor(jt->s, n_scope, scope, values);
    if (replaced) *replaced = hit ? 1 : 0;
    if (hit) jt->posts_stale = true;  // recalibrated under the kept evidence when the next posterior is asked for
    for (auto& peer : jt->peers) change_factor(peer->s, n_scope, scope, values);
  });
}

int hb_jt_em_step(hb_jt* jt, int64_t n_rows, const int8_t* evidence, double* out_values, int32_t* out_scopes, int32_t flags) {
  return guarded([&] {
    if (!jt || !evidence || !out_values || n_rows <= 0) throw Error(HB_E_ARG, "null argument or no samples (foldl1 of an empty list)");
    Session& s = jt->s;
    const Network& net = s.net;
    const int n_vars = net.n();
    if (s.device < 0) throw Error(HB_E_CUDA, "handle was compiled with HB_STRUCTURE_ONLY: no device, and there is no CPU fallback");
    if (!s.fixed.empty()) throw Error(HB_E_UNSUPPORTED, "hb_jt_em_step on a split handle");
    cuda_ok(cudaSetDevice(s.device), "cudaSetDevice");
    s.call_start = s.tick;
    const bool tracing = getenv("HB_EM_TRACE") != nullptr;  // host-side phase times on stderr (synchronises)
    auto t_last = std::chrono::steady_clock::now();
    auto trace = [&](const char* what) {
      if (!tracing) return;
      cudaStreamSynchronize(s.stream);
      auto now = std::chrono::steady_clock::now();
      fprintf(stderr, "[hb_jt_em_step] %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
      t_last = now;
    };
    // the samples on the host (they decide the programs), grouped by observed set in order of first appearance
    std::vector<int8_t> host_ev;
    const int8_t* dev_evidence = (flags & HB_DEVICE_PTRS) ? evidence : nullptr;
    if (flags & HB_DEVICE_PTRS) {
      host_ev.resize((size_t)n_rows * n_vars);
      cuda_ok(cudaMemcpyAsync(host_ev.data(), evidence, host_ev.size(), cudaMemcpyDeviceToHost, s.stream), "cudaMemcpy(evidence)");
      cuda_ok(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
      evidence = host_ev.data();
    }
    std::map<std::vector<int>, int> group_of;
    std::vector<std::vector<int>> group_obs;
    std::vector<int> row_group((size_t)n_rows);
    const int8_t* rep = nullptr;  // a row of the previous row's group: most batches have one observed set
    for (int64_t r = 0; r < n_rows; ++r) {
      const int8_t* row = evidence + r * n_vars;
      bool same = rep != nullptr;
      int v = 0;
      for (; same && v + 8 <= n_vars; v += 8) {  // observed <=> sign bit clear: compare the sign bits of 8 entries at once
        uint64_t a, b;
        memcpy(&a, row + v, 8), memcpy(&b, rep + v, 8);
        same = ((a ^ b) & 0x8080808080808080ull) == 0;
      }
      for (; same && v < n_vars; ++v) same = (row[v] >= 0) == (rep[v] >= 0);
      if (same) {
        row_group[(size_t)r] = row_group[(size_t)r - 1];
        continue;
      }
      std::vector<int> obs = observed_of_row(row, n_vars);
      auto it = group_of.find(obs);
      if (it == group_of.end()) it = group_of.emplace(obs, (int)group_obs.size()).first, group_obs.push_back(obs);
      row_group[(size_t)r] = it->second;
      rep = row;
    }
    trace("group rows");
    // queries: posterior jt' (factorVariables cpt) and posterior jt' (tail ...) per vertex (EM.hs:60-66)
    std::vector<int> flat, q_vertex, q_is_parents;
    for (int v = 0; v < n_vars; ++v) {
      if (!flat.empty()) flat.push_back(-1);
      flat.insert(flat.end(), net.scope[v].begin(), net.scope[v].end());
      q_vertex.push_back(v), q_is_parents.push_back(0);
      if (net.scope[v].size() > 1) {
        flat.push_back(-1);
        flat.insert(flat.end(), net.scope[v].begin() + 1, net.scope[v].end());
        q_vertex.push_back(v), q_is_parents.push_back(1);
      }
    }
    const int n_q = (int)q_vertex.size();
    // result layouts per group; the accumulators keep the layout of the first sample (cptSum: foldr1 union, acc first)
    struct Layout {
      std::vector<std::vector<int>> scope;  // per query
      std::vector<int64_t> col;             // per query: first column
    };
    auto layout_of = [&](const Program& P) {
      Layout L;
      int64_t at = 0;
      for (const Group& g : P.groups)
        if (g.kind == 2) {
          L.scope.push_back(g.out_scope);
          L.col.push_back(at);
          int64_t w = 1;
          for (int x : g.out_scope) w *= net.dims[x];
          at += w;
        }
      if ((int)L.scope.size() != n_q || at != P.out_width) throw Error(HB_E_ARG, "internal: EM query list does not match the program");
      return L;
    };
    std::vector<Compiled*> progs;
    std::vector<Layout> layouts;
    for (const auto& obs : group_obs) {
      progs.push_back(&jt->program(obs, flat));
      layouts.push_back(layout_of(*progs.back()->prog));
    }
    const Layout& L0 = layouts[0];
    const int64_t width = progs[0]->prog->out_width;
    // column of group g's layout -> column of the first sample's layout
    auto column_map = [&](const Layout& L) {
      std::vector<uint32_t> m((size_t)width);
      for (int q = 0; q < n_q; ++q) {
        const std::vector<int>& from = L.scope[q];
        const std::vector<int>& to = L0.scope[q];
        std::vector<int64_t> tstride(from.size());
        for (size_t i = 0; i < from.size(); ++i) {
          int64_t st = 1;
          size_t j = to.size();
          while (j-- > 0 && to[j] != from[i]) st *= net.dims[to[j]];
          tstride[i] = st;
        }
        int64_t n = 1;
        for (int x : from) n *= net.dims[x];
        for (int64_t e = 0; e < n; ++e) {
          int64_t rem = e, dst = 0;
          for (size_t i = from.size(); i-- > 0;) dst += (rem % net.dims[from[i]]) * tstride[i], rem /= net.dims[from[i]];
          m[(size_t)(L.col[q] + e)] = (uint32_t)(L0.col[q] + dst);
        }
      }
      return m;
    };
    // rows per chunk: the result matrix of a chunk stays below 1 GiB
    int64_t chunk = std::max<int64_t>(64, (((int64_t)1 << 30) / (width * 8)) / 64 * 64);
    if (const char* env = getenv("HB_EM_CHUNK_ROWS")) chunk = std::max<int64_t>(1, atoll(env));
    chunk = std::min(chunk, n_rows);
    const bool mixed = group_obs.size() > 1;
    DevBuf d_acc((size_t)width * 8);
    double* d_mat = (double*)s.em_mat.ensure((size_t)chunk * width * 8);
    int8_t* d_ev = (int8_t*)s.em_ev.ensure((size_t)chunk * n_vars);
    double* d_tmp = mixed ? (double*)s.em_tmp.ensure((size_t)chunk * width * 8) : nullptr;
    int64_t* d_rows = mixed ? (int64_t*)s.em_rows.ensure((size_t)chunk * 8) : nullptr;
    std::vector<uint32_t*> d_colmap(group_obs.size(), nullptr);
    std::vector<std::unique_ptr<DevBuf>> colmap_bufs;
    if (mixed)
      for (size_t g = 0; g < group_obs.size(); ++g) {
        std::vector<uint32_t> m = column_map(layouts[g]);
        colmap_bufs.push_back(std::make_unique<DevBuf>(m.size() * 4));
        d_colmap[g] = colmap_bufs.back()->as<uint32_t>();
        cuda_ok(cudaMemcpyAsync(d_colmap[g], m.data(), m.size() * 4, cudaMemcpyHostToDevice, s.stream), "cudaMemcpy(column map)");
        cuda_ok(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");  // m is a local
      }
    cuda_ok(cudaMemsetAsync(d_acc.p, 0, (size_t)width * 8, s.stream), "cudaMemset");
    trace("programs + buffers");
    RunStats rs;
    bool first = true;
    std::vector<int8_t> ev_rows;
    std::vector<int64_t> idx;
    for (int64_t r0 = 0; r0 < n_rows; r0 += chunk) {
      const int64_t m = std::min(chunk, n_rows - r0);
      if (!mixed) {
        if (!dev_evidence)
          cuda_ok(cudaMemcpyAsync(d_ev, evidence + r0 * n_vars, (size_t)m * n_vars, cudaMemcpyHostToDevice, s.stream), "cudaMemcpy(evidence)");
        executor_run(s.executor(*progs[0]), m, dev_evidence ? dev_evidence + r0 * n_vars : d_ev, d_mat, s.stream, &rs);
        s.note(*progs[0], rs, !first);
        first = false;
      } else {
        for (size_t g = 0; g < group_obs.size(); ++g) {
          ev_rows.clear();
          idx.clear();
          for (int64_t r = r0; r < r0 + m; ++r)
            if (row_group[(size_t)r] == (int)g) {
              ev_rows.insert(ev_rows.end(), evidence + r * n_vars, evidence + (r + 1) * n_vars);
              idx.push_back(r - r0);
            }
          if (idx.empty()) continue;
          const int64_t k = (int64_t)idx.size();
          cuda_ok(cudaMemcpyAsync(d_ev, ev_rows.data(), ev_rows.size(), cudaMemcpyHostToDevice, s.stream), "cudaMemcpy(evidence)");
          cuda_ok(cudaMemcpyAsync(d_rows, idx.data(), idx.size() * 8, cudaMemcpyHostToDevice, s.stream), "cudaMemcpy(row index)");
          executor_run(s.executor(*progs[g]), k, d_ev, d_tmp, s.stream, &rs);
          s.note(*progs[g], rs, !first);
          first = false;
          em_scatter(d_mat, d_tmp, d_rows, d_colmap[g], k, width, s.stream);
          cuda_ok(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");  // ev_rows / idx are reused
        }
      }
      trace("tree pass (chunk)");
      em_accumulate(d_acc.as<double>(), d_mat, m, width, s.stream);
      s.stats[0] += mixed ? 1 + (int64_t)group_obs.size() : 1;
      trace("sample-order sum");
    }
    for (Compiled* c : progs) executor_check(s.executor(*c), s.stream);
    // cptDivide: learnt table of v over (parents in the order of their posterior) ++ [v]
    std::vector<uint32_t> a_col;
    std::vector<int32_t> b_col;
    int scope_at = 0;
    for (int q = 0; q < n_q; ++q) {
      if (q_is_parents[q]) continue;
      const int v = q_vertex[q];
      const bool has_parents = q + 1 < n_q && q_is_parents[q + 1];
      std::vector<int> ref_scope = has_parents ? L0.scope[q + 1] : std::vector<int>{};
      ref_scope.push_back(v);
      const std::vector<int>& asc = L0.scope[q];
      std::vector<int64_t> astride(ref_scope.size());
      for (size_t i = 0; i < ref_scope.size(); ++i) {
        int64_t st = 1;
        size_t j = asc.size();
        while (j-- > 0 && asc[j] != ref_scope[i]) st *= net.dims[asc[j]];
        astride[i] = st;
      }
      int64_t n = 1;
      for (int x : ref_scope) n *= net.dims[x];
      for (int64_t e = 0; e < n; ++e) {
        int64_t rem = e, src = 0;
        for (size_t i = ref_scope.size(); i-- > 0;) src += (rem % net.dims[ref_scope[i]]) * astride[i], rem /= net.dims[ref_scope[i]];
        a_col.push_back((uint32_t)(L0.col[q] + src));
        b_col.push_back(has_parents ? (int32_t)(L0.col[q + 1] + e / net.dims[v]) : -1);
      }
      if (out_scopes)
        for (int x : ref_scope) out_scopes[scope_at++] = x;
    }
    const int64_t n_entries = (int64_t)a_col.size();
    DevBuf d_a((size_t)n_entries * 4), d_b((size_t)n_entries * 4), d_new((size_t)n_entries * 8);
    cuda_ok(cudaMemcpyAsync(d_a.p, a_col.data(), (size_t)n_entries * 4, cudaMemcpyHostToDevice, s.stream), "cudaMemcpy");
    cuda_ok(cudaMemcpyAsync(d_b.p, b_col.data(), (size_t)n_entries * 4, cudaMemcpyHostToDevice, s.stream), "cudaMemcpy");
    em_divide(d_new.as<double>(), d_acc.as<double>(), d_a.as<uint32_t>(), d_b.as<int32_t>(), n_entries, (double)n_rows, s.stream);
    s.stats[0] += 1;
    cuda_ok(cudaMemcpyAsync(out_values, d_new.p, (size_t)n_entries * 8, cudaMemcpyDeviceToHost, s.stream), "cudaMemcpy(out)");
    cuda_ok(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
    trace("divide + copy out");
  });
}
int hb_jt_set_split(hb_jt* jt, int32_t n, const int32_t* vars, const int32_t* states) {
  return guarded([&] {
    if (!jt) throw Error(HB_E_ARG, "null handle");
    check_evidence_list(jt->s.net, n, vars, states);
    jt->s.fixed.clear();
    for (int i = 0; i < n; ++i) jt->s.fixed.push_back({vars[i], states[i]});
    jt->s.cache.clear();
    jt->s.last_ex = jt->s.last_post_ex = nullptr;
  });
}
int hb_jt_normalise_batch(hb_jt* jt, int64_t n_rows, double* out, int32_t flags) {
  return guarded([&] {
    if (!jt || !out || n_rows < 0) throw Error(HB_E_ARG, "null argument");
    Session& s = jt->s;
    if (!s.last_post_ex) throw Error(HB_E_ARG, "hb_jt_posteriors_batch has not run on this handle yet");
    const int64_t width = s.net.width();
    if (flags & HB_DEVICE_PTRS) {
      executor_normalise(s.last_post_ex, n_rows, out, s.stream);
      return;
    }
    s.reserve(0, (size_t)n_rows * width * sizeof(double));
    cuda_ok(cudaMemcpyAsync(s.d_out, out, (size_t)n_rows * width * sizeof(double), cudaMemcpyHostToDevice, s.stream), "cudaMemcpy");
    executor_normalise(s.last_post_ex, n_rows, s.d_out, s.stream);
    cuda_ok(cudaMemcpyAsync(out, s.d_out, (size_t)n_rows * width * sizeof(double), cudaMemcpyDeviceToHost, s.stream), "cudaMemcpy");
    cuda_ok(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
  });
}
int hb_jt_specialise(hb_jt* jt, int32_t n_observed, const int32_t* observed, int32_t* n_kernels) {
  return guarded([&] {
    if (!jt || n_observed < 0 || (n_observed > 0 && !observed)) throw Error(HB_E_ARG, "null argument");
    const std::vector<int> obs(observed, observed + n_observed);
    Compiled& c = jt->program(obs, {});
    const int n = executor_specialise(jt->s.executor(c));
    if (n_kernels) *n_kernels = n;
    for (auto& peer : jt->peers) {  // the cubin is in this process' registry by now: the peers only load it
      Compiled& pc = peer->program(obs, {});
      executor_specialise(peer->s.executor(pc));
    }
  });
}
int hb_jt_specialised_kernels(const hb_jt* jt, int32_t* n) {
  if (!jt || !n) return HB_E_ARG;
  *n = jt->s.last_ex ? executor_specialised_kernels(jt->s.last_ex) : 0;
  return HB_OK;
}
int hb_jt_kernel_info(const hb_jt* jt, int64_t* out12) {
  if (!jt || !out12) return HB_E_ARG;
  for (int i = 0; i < 12; ++i) out12[i] = 0;
  out12[11] = -1;
  if (!jt->s.last_ex) return HB_OK;
  executor_tree_info(jt->s.last_ex, out12);
  out12[10] = executor_on_chip(jt->s.last_ex) ? 1 : 0;
  out12[11] = executor_jit_origin(jt->s.last_ex);
  return HB_OK;
}
int hb_jt_tree_profile(hb_jt* jt, int64_t* out64) {
  return guarded([&] {
    if (!jt || !out64 || !jt->s.last_ex) throw Error(HB_E_ARG, "no batch has run on this handle yet");
    executor_tree_profile(jt->s.last_ex, jt->s.stream, out64);
  });
}
// precision mode of the kernels GENERATED from now on; device-resident programs are dropped so that they are rebuilt
static int set_precision(Session& s, int32_t mode) {
  return guarded([&] {
    if (mode != HB_PRECISION_EXACT && mode != HB_PRECISION_CONTRACT) throw Error(HB_E_ARG, "unknown precision mode");
    const bool fma = mode == HB_PRECISION_CONTRACT;
    if (fma == s.fma) return;
    s.fma = fma;
    if (s.device >= 0) {
      cuda_ok(cudaSetDevice(s.device), "cudaSetDevice");
      cudaDeviceSynchronize();
    }
    for (auto& kv : s.cache)
      if (kv.second->ex) executor_destroy(kv.second->ex), kv.second->ex = nullptr;
    s.last_ex = s.last_post_ex = nullptr;
  });
}
int hb_jt_set_precision(hb_jt* jt, int32_t mode) {
  if (!jt) return HB_E_ARG;
  for (auto& peer : jt->peers) set_precision(peer->s, mode);
  return set_precision(jt->s, mode);
}
int hb_ve_set_precision(hb_ve* ve, int32_t mode) { return ve ? set_precision(ve->s, mode) : HB_E_ARG; }
int hb_jt_set_stream(hb_jt* jt, void* cuda_stream) {
  if (!jt) return HB_E_ARG;
  jt->s.stream = (cudaStream_t)cuda_stream;
  return HB_OK;
}
int hb_jt_set_workspace(hb_jt* jt, int64_t bytes) {
  if (!jt || bytes < 0) return HB_E_ARG;
  jt->s.workspace = bytes;
  jt->s.cache.clear();
  jt->s.last_ex = jt->s.last_post_ex = nullptr;
  for (auto& peer : jt->peers) peer->s.workspace = bytes, peer->s.cache.clear(), peer->s.last_ex = peer->s.last_post_ex = nullptr;
  return HB_OK;
}
int hb_jt_set_timing(hb_jt* jt, int32_t on) {
  if (!jt) return HB_E_ARG;
  jt->s.timing = on != 0;
  return HB_OK;
}
int hb_jt_last_timing(hb_jt* jt, double* out4) {
  return guarded([&] {
    if (!jt || !out4 || !jt->s.last_ex) throw Error(HB_E_ARG, "no timed run");
    executor_last_timing(jt->s.last_ex, jt->s.stream, out4);
  });
}
int hb_jt_last_stats(const hb_jt* jt, int64_t* out8) {
  if (!jt || !out8) return HB_E_ARG;
  memcpy(out8, jt->s.stats, sizeof jt->s.stats);
  return HB_OK;
}
void hb_jt_free(hb_jt* jt) { delete jt; }

int hb_factor_marginalize(int32_t n_vars, const int32_t* dims, int32_t k, const int32_t* scope_off, const int32_t* scopes,
                          const int64_t* val_off, const double* vals, int32_t sum_var, int32_t device, int32_t* out_scope,
                          int32_t* n_scope, double* out, int64_t out_len, int64_t* n_written) {
  return guarded([&] {
    if (n_vars <= 0 || !dims || k < 0 || !scope_off || !val_off || !out || !n_written || (k > 0 && (!scopes || !vals)))
      throw Error(HB_E_ARG, "null or empty argument");
    std::vector<FreeFactor> fs((size_t)k);
    bool mentions = sum_var < 0;
    for (int i = 0; i < k; ++i) {
      fs[i].scope.assign(scopes + scope_off[i], scopes + scope_off[i + 1]);
      fs[i].values.assign(vals + val_off[i], vals + val_off[i + 1]);
      for (int v : fs[i].scope) mentions = mentions || v == sum_var;
    }
    if (!mentions) throw Error(HB_E_ARG, "the variable to sum out occurs in no factor");
    int32_t dev_list[1] = {device};
    const int dev = pick_device(dev_list, 1);
    std::unique_ptr<FactorOp> op(compile_factor_op(std::vector<int>(dims, dims + n_vars), fs, sum_var));
    const Program& P = *op->prog;
    if (out_len < P.out_width) throw Error(HB_E_ARG, "output buffer too small");
    std::unique_ptr<Executor, void (*)(Executor*)> ex(executor_create(P, dev, 64 << 20), executor_destroy);
    std::vector<int8_t> row((size_t)n_vars, -1);
    int8_t* d_ev = nullptr;
    double* d_out = nullptr;
    cuda_ok(cudaMalloc(&d_ev, row.size()), "cudaMalloc");
    cuda_ok(cudaMalloc(&d_out, (size_t)P.out_width * sizeof(double)), "cudaMalloc");
    try {
      cuda_ok(cudaMemcpy(d_ev, row.data(), row.size(), cudaMemcpyHostToDevice), "cudaMemcpy");
      executor_run(ex.get(), 1, d_ev, d_out, nullptr, nullptr, /*normalise=*/false);
      executor_check(ex.get(), nullptr);
      cuda_ok(cudaMemcpy(out, d_out, (size_t)P.out_width * sizeof(double), cudaMemcpyDeviceToHost), "cudaMemcpy");
    } catch (...) {
      cudaFree(d_ev);
      cudaFree(d_out);
      throw;
    }
    cudaFree(d_ev);
    cudaFree(d_out);
    const Group& g = P.groups.back();
    if (out_scope)
      for (size_t i = 0; i < g.out_scope.size(); ++i) out_scope[i] = g.out_scope[i];
    if (n_scope) *n_scope = (int32_t)g.out_scope.size();
    *n_written = P.out_width;
  });
}

int hb_ve_compile(const hb_net* net, const int32_t* p, int32_t np, const int32_t* r, int32_t nr, const int32_t* devices,
                  int32_t n_devices, hb_ve** out) {
  return guarded([&] {
    if (!net || !out || np < 0 || nr <= 0 || (np > 0 && !p) || !r) throw Error(HB_E_ARG, "null or empty argument");
    auto ve = std::make_unique<hb_ve>();
    ve->s.net = net->net;
    ve->p.assign(p, p + np);
    ve->r.assign(r, r + nr);
    for (int v : ve->p)
      if (v < 0 || v >= net->net.n()) throw Error(HB_E_ARG, "elimination order mentions an unknown vertex");
    for (int v : ve->r)
      if (v < 0 || v >= net->net.n()) throw Error(HB_E_ARG, "query mentions an unknown vertex");
    ve->s.device = pick_device(devices, n_devices);
    *out = ve.release();
  });
}
int hb_ve_compile_mpe(const hb_net* net, const int32_t* p, int32_t np, const int32_t* r, int32_t nr, const int32_t* devices,
                      int32_t n_devices, hb_ve** out) {
  const int rc = hb_ve_compile(net, p, np, r, nr, devices, n_devices, out);
  if (rc == HB_OK) (*out)->mpe = true;
  return rc;
}
int hb_ve_mpe_order(hb_ve* ve, int32_t n_obs, const int32_t* observed, int32_t* order_out, int32_t* n_out) {
  return guarded([&] {
    if (!ve || !order_out || !n_out || n_obs < 0 || (n_obs > 0 && !observed)) throw Error(HB_E_ARG, "null argument");
    if (!ve->mpe) throw Error(HB_E_ARG, "not an mpe handle (hb_ve_compile_mpe)");
    const Program& P = *ve->program(std::vector<int>(observed, observed + n_obs)).prog;
    for (size_t i = 0; i < P.mpe_order.size(); ++i) order_out[i] = P.mpe_order[i];
    *n_out = (int32_t)P.mpe_order.size();
  });
}
// one pass of an mpe program over rows that share one observed set; device buffers
static void mpe_rows(hb_ve* ve, Compiled& c, int64_t n_rows, const int8_t* d_ev, int8_t* d_states, double* d_value, bool accumulate) {
  Session& s = ve->s;
  RunStats rs;
  executor_run(s.executor(c), n_rows, d_ev, d_value, s.stream, &rs, /*normalise=*/false, d_states);
  s.note(c, rs, accumulate);
}
int hb_ve_mpe_batch(hb_ve* ve, int64_t n_rows, const int8_t* evidence, int8_t* out_states, double* out_value, int32_t flags) {
  return guarded([&] {
    if (!ve || n_rows < 0 || (n_rows > 0 && (!evidence || !out_states))) throw Error(HB_E_ARG, "null argument");
    if (!ve->mpe) throw Error(HB_E_ARG, "not an mpe handle (hb_ve_compile_mpe)");
    if (n_rows == 0) return;
    Session& s = ve->s;
    const int n_vars = s.net.n();
    const int64_t nr = (int64_t)ve->r.size();
    if (s.device < 0) throw Error(HB_E_CUDA, "handle was compiled with HB_STRUCTURE_ONLY: no device, and there is no CPU fallback");
    cuda_ok(cudaSetDevice(s.device), "cudaSetDevice");
    s.call_start = s.tick;
    if (flags & HB_DEVICE_PTRS) {
      std::vector<int8_t> first(n_vars);
      cuda_ok(cudaMemcpyAsync(first.data(), evidence, n_vars, cudaMemcpyDeviceToHost, s.stream), "cudaMemcpy(first row)");
      cuda_ok(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
      Compiled& c = ve->program(observed_of_row(first.data(), n_vars));
      double* d_value = out_value ? out_value : (double*)s.em_mat.ensure((size_t)n_rows * 8);
      mpe_rows(ve, c, n_rows, evidence, out_states, d_value, false);
      executor_check(s.executor(c), s.stream);
      return;
    }
    // host buffers: rows grouped by observed set (one program each)
    std::map<std::vector<int>, std::vector<int64_t>> groups;
    for (int64_t r = 0; r < n_rows; ++r) groups[observed_of_row(evidence + r * n_vars, n_vars)].push_back(r);
    int8_t* d_ev = (int8_t*)s.em_ev.ensure((size_t)n_rows * n_vars);
    int8_t* d_states = (int8_t*)s.em_rows.ensure((size_t)n_rows * nr);
    double* d_value = (double*)s.em_mat.ensure((size_t)n_rows * 8);
    std::vector<int8_t> ev, st;
    std::vector<double> val;
    bool first = true;
    for (auto& kv : groups) {
      Compiled& c = ve->program(kv.first);
      const int64_t m = (int64_t)kv.second.size();
      const bool whole = m == n_rows;
      const int8_t* src = evidence;
      if (!whole) {
        ev.resize((size_t)m * n_vars);
        for (int64_t i = 0; i < m; ++i) memcpy(ev.data() + i * n_vars, evidence + kv.second[i] * n_vars, n_vars);
        src = ev.data();
      }
      cuda_ok(cudaMemcpyAsync(d_ev, src, (size_t)m * n_vars, cudaMemcpyHostToDevice, s.stream), "cudaMemcpy(evidence)");
      mpe_rows(ve, c, m, d_ev, d_states, d_value, !first);
      first = false;
      executor_check(s.executor(c), s.stream);
      if (whole) {
        cuda_ok(cudaMemcpyAsync(out_states, d_states, (size_t)m * nr, cudaMemcpyDeviceToHost, s.stream), "cudaMemcpy(states)");
        if (out_value) cuda_ok(cudaMemcpyAsync(out_value, d_value, (size_t)m * 8, cudaMemcpyDeviceToHost, s.stream), "cudaMemcpy(value)");
        cuda_ok(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
      } else {
        st.resize((size_t)m * nr), val.resize((size_t)m);
        cuda_ok(cudaMemcpyAsync(st.data(), d_states, st.size(), cudaMemcpyDeviceToHost, s.stream), "cudaMemcpy(states)");
        cuda_ok(cudaMemcpyAsync(val.data(), d_value, val.size() * 8, cudaMemcpyDeviceToHost, s.stream), "cudaMemcpy(value)");
        cuda_ok(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
        for (int64_t i = 0; i < m; ++i) {
          memcpy(out_states + kv.second[i] * nr, st.data() + i * nr, (size_t)nr);
          if (out_value) out_value[kv.second[i]] = val[(size_t)i];
        }
      }
    }
  });
}
int hb_ve_mpe(hb_ve* ve, int32_t n, const int32_t* vars, const int32_t* states, int8_t* out_states, double* out_value,
              int32_t* out_order, int32_t* n_order) {
  return guarded([&] {
    if (!ve || !out_states) throw Error(HB_E_ARG, "null argument");
    if (!ve->mpe) throw Error(HB_E_ARG, "not an mpe handle (hb_ve_compile_mpe)");
    check_evidence_list(ve->s.net, n, vars, states);
    Session& s = ve->s;
    if (s.device < 0) throw Error(HB_E_CUDA, "handle was compiled with HB_STRUCTURE_ONLY: no device, and there is no CPU fallback");
    cuda_ok(cudaSetDevice(s.device), "cudaSetDevice");
    s.call_start = s.tick;
    Compiled& c = ve->program(std::vector<int>(vars, vars + n));  // the assignment in the caller's order
    std::vector<int8_t> row(s.net.n(), -1);
    for (int i = 0; i < n; ++i) row[vars[i]] = (int8_t)states[i];
    const size_t nr = ve->r.size();
    int8_t* d_ev = (int8_t*)s.em_ev.ensure(row.size());
    int8_t* d_states = (int8_t*)s.em_rows.ensure(nr);
    double* d_value = (double*)s.em_mat.ensure(8);
    cuda_ok(cudaMemcpyAsync(d_ev, row.data(), row.size(), cudaMemcpyHostToDevice, s.stream), "cudaMemcpy(evidence)");
    mpe_rows(ve, c, 1, d_ev, d_states, d_value, false);
    executor_check(s.executor(c), s.stream);
    cuda_ok(cudaMemcpyAsync(out_states, d_states, nr, cudaMemcpyDeviceToHost, s.stream), "cudaMemcpy(states)");
    double value = 0.0;
    cuda_ok(cudaMemcpyAsync(&value, d_value, 8, cudaMemcpyDeviceToHost, s.stream), "cudaMemcpy(value)");
    cuda_ok(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
    if (out_value) *out_value = value;
    if (out_order)
      for (size_t i = 0; i < c.prog->mpe_order.size(); ++i) out_order[i] = c.prog->mpe_order[i];
    if (n_order) *n_order = (int32_t)c.prog->mpe_order.size();
  });
}
int hb_ve_result_scope(hb_ve* ve, int32_t n_obs, const int32_t* observed, int32_t* scope_out, int32_t* n_out) {
  return guarded([&] {
    if (!ve || !scope_out || !n_out || n_obs < 0 || (n_obs > 0 && !observed)) throw Error(HB_E_ARG, "null argument");
    const Group& g = ve->program(std::vector<int>(observed, observed + n_obs)).prog->groups.back();
    for (size_t i = 0; i < g.out_scope.size(); ++i) scope_out[i] = g.out_scope[i];
    *n_out = (int32_t)g.out_scope.size();
  });
}
int hb_ve_program_json(hb_ve* ve, int32_t n_obs, const int32_t* observed, char** out_json) {
  return guarded([&] {
    if (!ve || !out_json || n_obs < 0 || (n_obs > 0 && !observed)) throw Error(HB_E_ARG, "null argument");
    *out_json = dup_json(describe_program(*ve->program(std::vector<int>(observed, observed + n_obs)).prog));
  });
}
int hb_ve_posterior(hb_ve* ve, int32_t n, const int32_t* vars, const int32_t* states, int32_t* out_scope, double* out,
                    int64_t out_len, int64_t* n_written) {
  return guarded([&] {
    if (!ve || !out) throw Error(HB_E_ARG, "null argument");
    if (ve->mpe) throw Error(HB_E_ARG, "an mpe handle answers hb_ve_mpe / hb_ve_mpe_batch only");
    check_evidence_list(ve->s.net, n, vars, states);
    Compiled& c = ve->program(std::vector<int>(vars, vars + n));
    if (out_len < c.prog->out_width) throw Error(HB_E_ARG, "output buffer too small");
    Session& s = ve->s;
    s.call_start = s.tick;
    std::vector<int8_t> row(s.net.n(), -1);
    for (int i = 0; i < n; ++i) row[vars[i]] = (int8_t)states[i];
    s.reserve(row.size(), (size_t)c.prog->out_width * sizeof(double));
    cuda_ok(cudaMemcpyAsync(s.d_ev, row.data(), row.size(), cudaMemcpyHostToDevice, s.stream), "cudaMemcpy(evidence)");
    RunStats rs;
    executor_run(s.executor(c), 1, s.d_ev, s.d_out, s.stream, &rs);
    s.note(c, rs, false);
    executor_check(s.executor(c), s.stream);
    cuda_ok(cudaMemcpyAsync(out, s.d_out, (size_t)c.prog->out_width * sizeof(double), cudaMemcpyDeviceToHost, s.stream), "cudaMemcpy(out)");
    cuda_ok(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
    const Group& g = c.prog->groups.back();
    if (out_scope)
      for (size_t i = 0; i < g.out_scope.size(); ++i) out_scope[i] = g.out_scope[i];
    if (n_written) *n_written = c.prog->out_width;
  });
}
int hb_ve_posterior_batch(hb_ve* ve, int64_t n_rows, const int8_t* evidence, double* out, int32_t flags) {
  return guarded([&] {
    if (!ve) throw Error(HB_E_ARG, "null handle");
    if (ve->mpe) throw Error(HB_E_ARG, "an mpe handle answers hb_ve_mpe / hb_ve_mpe_batch only");
    run_matrix(ve->s, [&](const std::vector<int>& obs) -> Compiled& { return ve->program(obs); }, n_rows, evidence, out, flags);
  });
}
int hb_ve_change_factor(hb_ve* ve, int32_t n_scope, const int32_t* scope, const double* values, int32_t* replaced) {
  return guarded([&] {
    if (!ve) throw Error(HB_E_ARG, "null handle");
    const bool hit = change_factor(ve->s, n_scope, scope, values);
    if (replaced) *replaced = hit ? 1 : 0;
  });
}
int hb_ve_specialise(hb_ve* ve, int32_t n_observed, const int32_t* observed, int32_t* n_kernels) {
  return guarded([&] {
    if (!ve || n_observed < 0 || (n_observed > 0 && !observed)) throw Error(HB_E_ARG, "null argument");
    Compiled& c = ve->program(std::vector<int>(observed, observed + n_observed));
    const int n = executor_specialise(ve->s.executor(c));
    if (n_kernels) *n_kernels = n;
  });
}
int hb_ve_set_stream(hb_ve* ve, void* cuda_stream) {
  if (!ve) return HB_E_ARG;
  ve->s.stream = (cudaStream_t)cuda_stream;
  return HB_OK;
}
int hb_ve_set_workspace(hb_ve* ve, int64_t bytes) {
  if (!ve || bytes < 0) return HB_E_ARG;
  ve->s.workspace = bytes;
  ve->s.cache.clear();
  ve->s.last_ex = ve->s.last_post_ex = nullptr;
  return HB_OK;
}
int hb_ve_set_timing(hb_ve* ve, int32_t on) {
  if (!ve) return HB_E_ARG;
  ve->s.timing = on != 0;
  return HB_OK;
}
int hb_ve_last_timing(hb_ve* ve, double* out4) {
  return guarded([&] {
    if (!ve || !out4 || !ve->s.last_ex) throw Error(HB_E_ARG, "no timed run");
    executor_last_timing(ve->s.last_ex, ve->s.stream, out4);
  });
}
int hb_ve_last_stats(const hb_ve* ve, int64_t* out8) {
  if (!ve || !out8) return HB_E_ARG;
  memcpy(out8, ve->s.stats, sizeof ve->s.stats);
  return HB_OK;
}
void hb_ve_free(hb_ve* ve) { delete ve; }

}  // extern "C"
