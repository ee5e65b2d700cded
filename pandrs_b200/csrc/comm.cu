// Multi-GPU plumbing behind the C ABI: one process per GPU, NCCL over NVLink / NVSwitch for the collectives.
//
// pandrs has no communication backend; the only reference-side notion is PartitionStrategy::Hash
// (src/distributed/core/partition.rs:11-18).  The operators built on this file (pdrs_groupby_agg_dist in groupby.cu,
// pdrs_join_pairs_dist below) apply the single-frame semantics of the reference to the union of the ranks' rows.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 - the copy already loaded by the host process, e.g. torch's, else the
// system one), so libpandrs_b200.so has no link-time dependency on it and single-GPU users never load it.  Only the
// stable C entry points are used; their prototypes are restated here (nccl.h 2.x).
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "comm.cuh"

namespace {

typedef struct { char internal[128]; } nccl_unique_id;
typedef void* nccl_comm_t;
enum { NCCL_UINT8 = 1 };
struct NcclApi {
  int (*GetUniqueId)(nccl_unique_id*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
  std::string err;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy the host process already uses (torch bundles one)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { g_nccl.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "?"); return; }
  auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) g_nccl.err += std::string(" missing symbol ") + n; return p; };
  g_nccl.GetUniqueId = (int (*)(nccl_unique_id*))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(nccl_comm_t*, int, nccl_unique_id, int))sym("ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(nccl_comm_t))sym("ncclCommDestroy");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t))sym("ncclAllGather");
  g_nccl.Send = (int (*)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclSend");
  g_nccl.Recv = (int (*)(void*, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclRecv");
  g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
  g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
  g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  g_nccl.ok = g_nccl.err.empty();
}

#define PDRS_NCCL(c, call)                                                                                       \
  do {                                                                                                           \
    int _r = (call);                                                                                             \
    if (_r != 0) return pdrs_fail((c), PDRS_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
  } while (0)

}  // namespace

int32_t pdrs_comm_allgather(pdrs_comm* cm, const void* send, void* recv, size_t bytes) {
  pdrs_ctx* c = cm->ctx;
  if (cm->world == 1) { if (bytes && send != recv) PDRS_CUDA(c, cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, c->stream)); return PDRS_OK; }
  PDRS_NCCL(c, g_nccl.AllGather(send, recv, bytes, NCCL_UINT8, (nccl_comm_t)cm->nccl, c->stream));
  return PDRS_OK;
}

int32_t pdrs_comm_alltoallv(pdrs_comm* cm, const void* send, const size_t* soff, const size_t* sbytes, void* recv, const size_t* roff, const size_t* rbytes) {
  pdrs_ctx* c = cm->ctx;
  if (cm->world == 1) { if (sbytes[0]) PDRS_CUDA(c, cudaMemcpyAsync((char*)recv + roff[0], (const char*)send + soff[0], sbytes[0], cudaMemcpyDeviceToDevice, c->stream)); return PDRS_OK; }
  PDRS_NCCL(c, g_nccl.GroupStart());
  for (int r = 0; r < cm->world; r++) {
    if (sbytes[r]) PDRS_NCCL(c, g_nccl.Send((const char*)send + soff[r], sbytes[r], NCCL_UINT8, r, (nccl_comm_t)cm->nccl, c->stream));
    if (rbytes[r]) PDRS_NCCL(c, g_nccl.Recv((char*)recv + roff[r], rbytes[r], NCCL_UINT8, r, (nccl_comm_t)cm->nccl, c->stream));
  }
  PDRS_NCCL(c, g_nccl.GroupEnd());
  return PDRS_OK;
}

int32_t pdrs_comm_allgather_host(pdrs_comm* cm, const void* mine, void* all, size_t bytes) {
  pdrs_ctx* c = cm->ctx;
  if (cm->world == 1) { memcpy(all, mine, bytes); return PDRS_OK; }
  DevBuf s, r;
  PDRS_TRY(s.alloc(c, bytes));
  PDRS_TRY(r.alloc(c, bytes * cm->world));
  PDRS_CUDA(c, cudaMemcpyAsync(s.p, mine, bytes, cudaMemcpyHostToDevice, c->stream));
  PDRS_TRY(pdrs_comm_allgather(cm, s.p, r.p, bytes));
  PDRS_CUDA(c, cudaMemcpyAsync(all, r.p, bytes * cm->world, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PDRS_OK;
}

extern "C" {

int32_t pdrs_comm_unique_id(uint8_t* id128) {
  if (!id128) return PDRS_ERR_BAD_ARG;
  std::call_once(g_nccl_once, load_nccl);
  if (!g_nccl.ok) return pdrs_fail(nullptr, PDRS_ERR_NCCL, "NCCL is not available: %s", g_nccl.err.c_str());
  nccl_unique_id id;
  const int r = g_nccl.GetUniqueId(&id);
  if (r != 0) return pdrs_fail(nullptr, PDRS_ERR_NCCL, "ncclGetUniqueId failed: %s", g_nccl.GetErrorString(r));
  memcpy(id128, &id, 128);
  return PDRS_OK;
}

int32_t pdrs_comm_init(pdrs_ctx* c, int32_t nranks, int32_t rank, const uint8_t* id128, pdrs_comm** out) {
  if (!c) return PDRS_ERR_BAD_ARG;
  if (!out || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id128)) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_comm_init: bad argument (nranks %d, rank %d)", nranks, rank);
  PDRS_CUDA(c, cudaSetDevice(c->device));
  auto* cm = new pdrs_comm();
  cm->ctx = c; cm->rank = rank; cm->world = nranks;
  if (nranks > 1) {
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.ok) { delete cm; return pdrs_fail(c, PDRS_ERR_NCCL, "NCCL is not available: %s", g_nccl.err.c_str()); }
    nccl_unique_id id;
    memcpy(&id, id128, 128);
    nccl_comm_t nc = nullptr;
    const int r = g_nccl.CommInitRank(&nc, nranks, id, rank);
    if (r != 0) { delete cm; return pdrs_fail(c, PDRS_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); }
    cm->nccl = nc;
  }
  *out = cm;
  return PDRS_OK;
}

int32_t pdrs_comm_rank(const pdrs_comm* cm) { return cm ? cm->rank : -1; }
int32_t pdrs_comm_size(const pdrs_comm* cm) { return cm ? cm->world : -1; }

int32_t pdrs_comm_set_option(pdrs_comm* cm, const char* name, int64_t value) {
  if (!cm || !name) return PDRS_ERR_BAD_ARG;
  if (!strcmp(name, "groups_cap")) { if (value < 1) return pdrs_fail(cm->ctx, PDRS_ERR_BAD_ARG, "groups_cap must be positive"); cm->groups_cap = value; }
  else return pdrs_fail(cm->ctx, PDRS_ERR_BAD_ARG, "unknown communicator option '%s'", name);
  return PDRS_OK;
}

// time (CUDA events on the context stream) and bytes sent to other ranks by the exchange step of the last *_dist call
int32_t pdrs_comm_last_exchange(const pdrs_comm* cm, float* ms, int64_t* bytes_to_peers) {
  if (!cm) return PDRS_ERR_BAD_ARG;
  if (ms) *ms = cm->last_exchange_ms;
  if (bytes_to_peers) *bytes_to_peers = cm->last_exchange_bytes;
  return PDRS_OK;
}

int32_t pdrs_comm_barrier(pdrs_comm* cm) {
  if (!cm) return PDRS_ERR_BAD_ARG;
  uint64_t mine = 1;
  std::vector<uint64_t> all((size_t)cm->world);
  return pdrs_comm_allgather_host(cm, &mine, all.data(), 8);
}

// ---------------------------------------------------------------- sharded join behind one call
// The exchange join of join.cu (pdrs_xjoin_*: the partition pass stores its runs straight into the destination GPU's receive
// area through NVLink) with its protocol - IPC handle exchange, shuffle, barrier, local join - carried by the communicator.
// Every rank passes its shard of both key columns and the global number of its first row on each side; rank r returns the
// pairs of the keys whose rank hash maps to r, in GLOBAL row numbers.  max_* describe the largest shard of any rank (same
// values on every rank): the receive areas are allocated for them once and reused by later calls.
// one shuffle + local join; right_key == NULL: a later round (only left rows travel, the table of round 0 is probed again);
// the pairs are appended to `res`
static int32_t join_pairs_dist_round(pdrs_comm* cm, const pdrs_col* left_key, const pdrs_col* right_key, int32_t how, int64_t left_row0, int64_t right_row0,
                                     int64_t max_left_rows, int64_t max_right_rows, int64_t total_right_rows, pdrs_join_result* res, int64_t cap_hint);

int32_t pdrs_join_pairs_dist(pdrs_comm* cm, const pdrs_col* left_key, const pdrs_col* right_key, int32_t how, int64_t left_row0, int64_t right_row0,
                             int64_t max_left_rows, int64_t max_right_rows, int64_t total_right_rows, pdrs_join_result** out) {
  if (!cm) return PDRS_ERR_BAD_ARG;
  pdrs_ctx* c = cm->ctx;
  if (!left_key || !right_key || !out) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_join_pairs_dist: NULL argument");
  if (how != PDRS_INNER && how != PDRS_LEFT) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_join_pairs_dist: only Inner and Left joins are sharded");
  if (left_key->len > max_left_rows || max_left_rows < 0) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_join_pairs_dist: %lld left rows exceed max_left_rows = %lld", (long long)left_key->len, (long long)max_left_rows);
  // The staged exchange (line-aligned peer stores at ~80% of the NVLink peak) carries a left row as (source rank << s | local row) in 32
  // bits, s = 32 - log2(ranks): larger shards are joined in ROUNDS of < 2^s left rows each against the same build side (the
  // probe rows of a join are independent of each other): round 0 shuffles both sides and builds the hash table, the later rounds only
  // shuffle and probe their slice of the left rows.
  int log_world = 0;
  while ((1 << log_world) < cm->world) log_world++;
  int64_t lim = (1ll << (32 - log_world)) - 64;
  if (c->opt_xjoin_round_rows > 0) lim = std::min<int64_t>(lim, std::max<int64_t>(64, c->opt_xjoin_round_rows));
  int rounds = 1;
  if ((cm->world > 1 && c->opt_xjoin_mode != 1) || c->opt_xjoin_round_rows > 0) while ((max_left_rows + rounds - 1) / rounds > lim) rounds++;
  const int64_t chunk = rounds == 1 ? max_left_rows : ((max_left_rows + rounds - 1) / rounds + 63) / 64 * 64;       // bitmap bytes and 128-bit loads stay aligned
  pdrs_join_result* res = pdrs_join_result_new(c);
  float ex_ms = 0.f;
  int64_t ex_bytes = 0;
  int32_t rc = PDRS_OK;
  for (int r = 0; r < rounds && rc == PDRS_OK; r++) {
    const int64_t lo = std::min<int64_t>(left_key->len, r * chunk), hi = std::min<int64_t>(left_key->len, lo + chunk);
    pdrs_col sl = *left_key;
    sl.len = hi - lo;
    sl.data = (const char*)left_key->data + (left_key->dtype == PDRS_BOOL_BITS ? lo / 8 : lo * (int64_t)pdrs_dtype_bytes(left_key->dtype));
    if (left_key->null_bits) {
      const int64_t have = std::max<int64_t>(0, left_key->null_len - lo / 8);
      sl.null_bits = have > 0 ? left_key->null_bits + lo / 8 : nullptr;
      sl.null_len = have;
    }
    // collective: every rank runs every round.  Round 0 shuffles the build side too and builds the table; the later rounds only
    // shuffle and probe their slice of the left rows; the pairs of all rounds are appended to one result (room for all of them
    // is reserved in round 0: a join on unique build keys emits at most one pair per left row)
    rc = join_pairs_dist_round(cm, &sl, r == 0 ? right_key : nullptr, how, left_row0 + lo, right_row0, chunk, max_right_rows, total_right_rows, res,
                               rounds > 1 ? left_key->len + left_key->len / 16 + 1024 : 0);
    if (rc == PDRS_OK) { ex_ms += cm->last_exchange_ms; ex_bytes += cm->last_exchange_bytes; }
  }
  if (rc != PDRS_OK) { pdrs_join_result_free(res); return rc; }
  cm->last_exchange_ms = ex_ms; cm->last_exchange_bytes = ex_bytes;
  *out = res;
  return PDRS_OK;
}

static int32_t join_pairs_dist_round(pdrs_comm* cm, const pdrs_col* left_key, const pdrs_col* right_key, int32_t how, int64_t left_row0, int64_t right_row0,
                                     int64_t max_left_rows, int64_t max_right_rows, int64_t total_right_rows, pdrs_join_result* res, int64_t cap_hint) {
  pdrs_ctx* c = cm->ctx;
  PDRS_CUDA(c, cudaSetDevice(c->device));
  if (!cm->xj || cm->xj_left != max_left_rows || cm->xj_right != max_right_rows || cm->xj_total_right != total_right_rows) {
    if (cm->xj) { PDRS_TRY(pdrs_comm_barrier(cm)); pdrs_xjoin_destroy(cm->xj); cm->xj = nullptr; }
    uint8_t mine[72] = {0};
    int32_t rc = pdrs_xjoin_create(c, cm->rank, cm->world, max_left_rows, max_right_rows, total_right_rows, &cm->xj);
    if (rc == PDRS_OK) rc = pdrs_xjoin_ipc_handle(cm->xj, mine);
    mine[64] = rc == PDRS_OK ? 1 : 0;
    std::vector<uint8_t> all((size_t)72 * cm->world);
    PDRS_TRY(pdrs_comm_allgather_host(cm, mine, all.data(), 72));
    bool ok = true;
    for (int r = 0; r < cm->world; r++) ok = ok && all[(size_t)72 * r + 64];
    if (ok && cm->world > 1) {
      std::vector<uint8_t> handles((size_t)64 * cm->world);
      for (int r = 0; r < cm->world; r++) memcpy(&handles[(size_t)64 * r], &all[(size_t)72 * r], 64);
      rc = pdrs_xjoin_attach_ipc(cm->xj, handles.data());
      uint8_t f = rc == PDRS_OK ? 1 : 0;
      std::vector<uint8_t> fa((size_t)cm->world);
      PDRS_TRY(pdrs_comm_allgather_host(cm, &f, fa.data(), 1));
      for (int r = 0; r < cm->world; r++) ok = ok && fa[r];
    }
    if (!ok) {
      const std::string why = c->err;
      if (cm->xj) { pdrs_xjoin_destroy(cm->xj); cm->xj = nullptr; }
      return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_join_pairs_dist: the exchange areas could not be set up on every rank (%s)", why.c_str());
    }
    cm->xj_left = max_left_rows; cm->xj_right = max_right_rows; cm->xj_total_right = total_right_rows;
  } else {
    PDRS_TRY(pdrs_comm_barrier(cm));       // every rank has finished reading its receive area (previous call)
  }
  int32_t rc = pdrs_xjoin_shuffle(cm->xj, left_key, right_key, right_row0);
  cm->last_exchange_ms = c->stats.total_ms;
  cm->last_exchange_bytes = (int64_t)((double)(left_key->len + (right_key ? right_key->len : 0)) * 12.0 * (cm->world - 1) / cm->world);
  struct Info { int64_t ok, left_row0; } mine{rc == PDRS_OK ? 1 : 0, left_row0};
  std::vector<Info> all((size_t)cm->world);
  PDRS_TRY(pdrs_comm_allgather_host(cm, &mine, all.data(), sizeof(Info)));      // doubles as the barrier: all stores into my area are complete
  std::vector<int64_t> row0((size_t)cm->world);
  bool ok = true;
  for (int r = 0; r < cm->world; r++) { ok = ok && all[r].ok; row0[r] = all[r].left_row0; }
  if (!ok) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_join_pairs_dist: a padded region overflowed on some rank (skewed keys)");
  rc = pdrs_xjoin_local_append(cm->xj, how, row0.data(), res, cap_hint);
  const std::string local_err = c->err;
  uint8_t f = rc == PDRS_OK ? 1 : 0;           // local() can fail on one rank alone: agree before anybody returns
  std::vector<uint8_t> fa((size_t)cm->world);
  PDRS_TRY(pdrs_comm_allgather_host(cm, &f, fa.data(), 1));
  for (int r = 0; r < cm->world; r++) ok = ok && fa[r];
  if (!ok) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_join_pairs_dist: the local join failed on some rank (%s)", rc == PDRS_OK ? "another rank" : local_err.c_str());
  return PDRS_OK;
}

void pdrs_comm_destroy(pdrs_comm* cm) {
  if (!cm) return;
  cudaSetDevice(cm->ctx->device);
  cudaStreamSynchronize(cm->ctx->stream);
  if (cm->xj) pdrs_xjoin_destroy(cm->xj);
  if (cm->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((nccl_comm_t)cm->nccl);
  delete cm;
}

}  // extern "C"
