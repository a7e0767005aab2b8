// NCCL communicator group + the packed all-gather / merge exchange step (host API; see comm.cu).
#pragma once
#include "common.hpp"

#include <memory>
#include <vector>

namespace b200 {

struct comm_group;

/** ncclCommInitAll over the given devices of this process (single-process multi-GPU handles). */
comm_group* make_local_comm_group(const std::vector<int>& devices);
void destroy_comm_group(comm_group* g);

/**
 * Single-process exchange: device c holds the partial top-k (d[c], i[c]) [nq, k] (global ids) on streams[c]; one grouped
 * ncclAllGather of the packed partials, then the k-way merge on device `only_output_device` (or on every device when -1)
 * into out_d / out_i.  Everything is enqueued on the devices' streams; no host synchronisation.
 */
void allgather_merge_topk_all(comm_group& g, const std::vector<cudaStream_t>& streams, const std::vector<const float*>& d,
                              const std::vector<const int64_t*>& i, int64_t nq, int k, bool select_min, const std::vector<float*>& out_d,
                              const std::vector<int64_t*>& out_i, int only_output_device);

}  // namespace b200
