// nccl_dyn.hpp -- NCCL, bound at run time (dlopen), for the one exchange step of the path: the sum over the split
// variables of a clique table that is sliced over the GPUs of a box.  The library itself only links
// the CUDA runtime; a process that never splits a clique never loads NCCL.
#pragma once
#include <cstddef>
#include <string>

namespace hb {

constexpr size_t NCCL_ID_BYTES = 128;  // sizeof(ncclUniqueId)

// throws Error(HB_E_NCCL) when libnccl cannot be loaded or a call fails
void nccl_unique_id(void* out128);
void* nccl_comm_init(int world, int rank, const void* id128);  // ncclCommInitRank on the current device
void nccl_comm_destroy(void* comm);
// in-place sum of `count` doubles over the ranks of `comm`, asynchronous on `stream`
void nccl_allreduce_sum(void* comm, double* buf, size_t count, void* stream);
void nccl_group_start();
void nccl_group_end();

}  // namespace hb
