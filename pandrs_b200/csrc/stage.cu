// Host -> device staging engine of the chunked (out-of-core) operators.
//
// pandrs columns live in ordinary Rust heap memory (Arc<[i64]> / Arc<[f64]>, src/column/int64_column.rs:9-13): PAGEABLE for
// CUDA.  cudaMemcpy from pageable memory goes through one driver-internal bounce buffer and reaches ~11 GB/s on this box, a
// fifth of the PCIe 5 x16 link.  Here a pool of worker threads copies 8 MB pieces of the source into their own pinned slots
// (two per worker, so a worker fills one slot while the DMA engine drains the other) and issues the H2D copies on their own
// streams; eight workers keep the link busy.  Sources that are already pinned (cudaHostAlloc / cudaHostRegister) skip the
// bounce copy.  The consumer stream is ordered behind the copies with one event per worker (pdrs_stage_join).
// Large results travel the other way through the same workers (pdrs_copy_to_host: DMA into a pinned slot, copy-out by the thread).
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace {

constexpr size_t PIECE = 8u << 20;

struct Piece { char* dst; const char* src; size_t bytes; bool direct; bool d2h; };

struct Worker {
  cudaStream_t s = nullptr;
  char* slot[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool used[2] = {false, false};
  int cur = 0;
  cudaEvent_t done = nullptr;
};

}  // namespace

struct PdrsStager {
  pdrs_ctx* c = nullptr;
  std::vector<std::thread> threads;
  std::vector<Worker> w;
  std::mutex m;
  std::condition_variable cv_work, cv_done;
  std::deque<Piece> q;
  size_t inflight = 0;
  bool stop = false;
  cudaError_t err = cudaSuccess;

  void run(int id) {
    cudaSetDevice(c->device);
    Worker& me = w[id];
    for (;;) {
      Piece p;
      {
        std::unique_lock<std::mutex> lk(m);
        cv_work.wait(lk, [&] { return stop || !q.empty(); });
        if (q.empty()) return;          // stop requested and nothing left
        p = q.front();
        q.pop_front();
      }
      cudaError_t e = cudaSuccess;
      if (p.direct) {
        e = cudaMemcpyAsync(p.dst, p.src, p.bytes, p.d2h ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice, me.s);
      } else if (p.d2h) {
        // device -> pageable host: DMA into the pinned slot, then this thread copies it out (the other workers' DMAs run meanwhile)
        const int k = me.cur;
        me.cur ^= 1;
        if (me.used[k]) e = cudaEventSynchronize(me.ev[k]);
        if (e == cudaSuccess) e = cudaMemcpyAsync(me.slot[k], p.src, p.bytes, cudaMemcpyDeviceToHost, me.s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(me.s);
        if (e == cudaSuccess) memcpy(p.dst, me.slot[k], p.bytes);
        me.used[k] = false;
      } else {
        const int k = me.cur;
        me.cur ^= 1;
        if (me.used[k]) e = cudaEventSynchronize(me.ev[k]);     // the DMA that last read this slot
        if (e == cudaSuccess) {
          memcpy(me.slot[k], p.src, p.bytes);
          e = cudaMemcpyAsync(p.dst, me.slot[k], p.bytes, cudaMemcpyHostToDevice, me.s);
          if (e == cudaSuccess) e = cudaEventRecord(me.ev[k], me.s);
          me.used[k] = true;
        }
      }
      {
        std::lock_guard<std::mutex> lk(m);
        if (e != cudaSuccess && err == cudaSuccess) err = e;
        if (--inflight == 0) cv_done.notify_all();
      }
    }
  }
};

static int32_t stager_get(pdrs_ctx* c, PdrsStager** out) {
  if (c->stager) { *out = c->stager; return PDRS_OK; }
  int T = (int)c->opt_stage_threads;
  if (T <= 0) {
    const unsigned hc = std::thread::hardware_concurrency();
    T = (int)std::min<unsigned>(8u, std::max<unsigned>(2u, hc / 2u));
  }
  T = std::min(T, 32);
  auto* st = new PdrsStager();
  st->c = c;
  st->w.resize(T);
  for (int i = 0; i < T; i++) {
    Worker& k = st->w[i];
    cudaError_t e = cudaStreamCreateWithFlags(&k.s, cudaStreamNonBlocking);
    for (int b = 0; b < 2 && e == cudaSuccess; b++) {
      e = cudaHostAlloc((void**)&k.slot[b], PIECE, cudaHostAllocDefault);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&k.ev[b], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&k.done, cudaEventDisableTiming);
    if (e != cudaSuccess) {
      c->stager = st;               // so that destroy releases what exists
      pdrs_stage_destroy(c);
      return pdrs_fail(c, e == cudaErrorMemoryAllocation ? PDRS_ERR_OOM : PDRS_ERR_CUDA, "staging engine: %s", cudaGetErrorString(e));
    }
  }
  for (int i = 0; i < T; i++) st->threads.emplace_back([st, i] { st->run(i); });
  c->stager = st;
  *out = st;
  return PDRS_OK;
}

int32_t pdrs_stage_copy_async(pdrs_ctx* c, void* dst_dev, const void* src_host, size_t bytes, int d2h) {
  if (bytes == 0) return PDRS_OK;
  PdrsStager* st = nullptr;
  PDRS_TRY(stager_get(c, &st));
  cudaPointerAttributes at{};
  bool direct = false;
  if (cudaPointerGetAttributes(&at, d2h ? (const void*)dst_dev : src_host) == cudaSuccess) direct = at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged || at.type == cudaMemoryTypeDevice;
  else cudaGetLastError();
  if (c->opt_stage_threads < 0) direct = true;   // -1: plain cudaMemcpyAsync from whatever the source is (the driver's pageable path)
  {
    std::lock_guard<std::mutex> lk(st->m);
    const size_t step = direct ? (64u << 20) : PIECE;
    for (size_t off = 0; off < bytes; off += step) {
      st->q.push_back({(char*)dst_dev + off, (const char*)src_host + off, std::min(step, bytes - off), direct, d2h != 0});
      st->inflight++;
    }
  }
  st->cv_work.notify_all();
  return PDRS_OK;
}

// Waits until every queued piece has been ISSUED, then orders `consumer` behind the workers' streams.
int32_t pdrs_stage_join(pdrs_ctx* c, cudaStream_t consumer) {
  PdrsStager* st = c->stager;
  if (!st) return PDRS_OK;
  cudaError_t e;
  {
    std::unique_lock<std::mutex> lk(st->m);
    st->cv_done.wait(lk, [&] { return st->inflight == 0; });
    e = st->err;
    st->err = cudaSuccess;
  }
  if (e != cudaSuccess) return pdrs_fail(c, PDRS_ERR_CUDA, "staging copy failed: %s", cudaGetErrorString(e));
  for (auto& k : st->w) {
    PDRS_CUDA(c, cudaEventRecord(k.done, k.s));
    PDRS_CUDA(c, cudaStreamWaitEvent(consumer, k.done, 0));
  }
  return PDRS_OK;
}

void pdrs_stage_destroy(pdrs_ctx* c) {
  PdrsStager* st = c->stager;
  if (!st) return;
  {
    std::lock_guard<std::mutex> lk(st->m);
    st->stop = true;
  }
  st->cv_work.notify_all();
  for (auto& t : st->threads) t.join();
  for (auto& k : st->w) {
    if (k.s) { cudaStreamSynchronize(k.s); cudaStreamDestroy(k.s); }
    for (int b = 0; b < 2; b++) { if (k.slot[b]) cudaFreeHost(k.slot[b]); if (k.ev[b]) cudaEventDestroy(k.ev[b]); }
    if (k.done) cudaEventDestroy(k.done);
  }
  delete st;
  c->stager = nullptr;
}
