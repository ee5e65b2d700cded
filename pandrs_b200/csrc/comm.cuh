// Internal: the multi-GPU communicator behind pdrs_comm_* (comm.cu) - one rank = one process = one GPU.
#pragma once
#include "common.cuh"

struct pdrs_xjoin;

struct pdrs_comm {
  pdrs_ctx* ctx = nullptr;
  int rank = 0, world = 1;
  void* nccl = nullptr;                 // ncclComm_t (NULL when world == 1: every collective is a copy)
  DevBuf send, recv;                    // exchange buffers of the groupby (grown on demand, reused between calls)
  int64_t groups_cap = 4096;            // replicated groupby: state rows per rank in the fixed-size all-gather
  // exchange join (pdrs_join_pairs_dist): the receive areas are set up once per (max rows) signature and reused
  pdrs_xjoin* xj = nullptr;
  int64_t xj_left = 0, xj_right = 0, xj_total_right = 0;
  float last_exchange_ms = 0.f;         // CUDA-event time of the last collective (all_gather / all_to_all / peer-store shuffle)
  int64_t last_exchange_bytes = 0;      // bytes this rank sent to OTHER ranks in it
};

// all on the context's stream; world == 1 degenerates to device-to-device copies
int32_t pdrs_comm_allgather(pdrs_comm* cm, const void* send, void* recv, size_t bytes_per_rank);
// rank r gets send[send_off[r] .. + send_bytes[r]) of every peer into recv[recv_off[src] .. + recv_bytes[src])
int32_t pdrs_comm_alltoallv(pdrs_comm* cm, const void* send, const size_t* send_off, const size_t* send_bytes, void* recv, const size_t* recv_off,
                            const size_t* recv_bytes);
// host-visible all-gather of a few bytes per rank (device staging + one stream synchronisation): the "everybody agrees" step
int32_t pdrs_comm_allgather_host(pdrs_comm* cm, const void* mine, void* all, size_t bytes_per_rank);

struct pdrs_join_result;
int32_t pdrs_join_result_concat(pdrs_ctx* c, pdrs_join_result** parts, int n, pdrs_join_result** out);   // join.cu
// join.cu: build (unless the table of the current build side exists) + probe, pairs appended to `res`; cap_hint = pairs to make room for
int32_t pdrs_xjoin_local_append(pdrs_xjoin* x, int32_t how, const int64_t* left_row0, pdrs_join_result* res, int64_t cap_hint);
pdrs_join_result* pdrs_join_result_new(pdrs_ctx* c);
