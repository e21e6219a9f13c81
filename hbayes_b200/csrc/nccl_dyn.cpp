// nccl_dyn.cpp -- see nccl_dyn.hpp
#include "nccl_dyn.hpp"

#include <dlfcn.h>

#include <cstring>

#include "model.hpp"

namespace hb {
namespace {

struct Api {
  int (*GetUniqueId)(void*) = nullptr;
  void* CommInitRank = nullptr;  // ncclCommInitRank takes the 128-byte id BY VALUE: called through a typed pointer below
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, void*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

struct IdByValue {  // same size and alignment as ncclUniqueId { char internal[128]; }
  char internal[NCCL_ID_BYTES];
};

const Api& api() {
  static Api a = [] {
    Api r;
    void* h = nullptr;
    for (const char* n : {"libnccl.so.2", "libnccl.so"})
      if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) return r;
    *(void**)(&r.GetUniqueId) = dlsym(h, "ncclGetUniqueId");
    r.CommInitRank = dlsym(h, "ncclCommInitRank");
    *(void**)(&r.CommDestroy) = dlsym(h, "ncclCommDestroy");
    *(void**)(&r.AllReduce) = dlsym(h, "ncclAllReduce");
    *(void**)(&r.GroupStart) = dlsym(h, "ncclGroupStart");
    *(void**)(&r.GroupEnd) = dlsym(h, "ncclGroupEnd");
    *(void**)(&r.GetErrorString) = dlsym(h, "ncclGetErrorString");
    r.ok = r.GetUniqueId && r.CommInitRank && r.CommDestroy && r.AllReduce && r.GroupStart && r.GroupEnd && r.GetErrorString;
    return r;
  }();
  if (!a.ok) throw Error(HB_E_NCCL, "libnccl.so.2 could not be loaded (a split clique over several GPUs needs NCCL)");
  return a;
}

void check(int rc, const char* what) {
  if (rc != 0) throw Error(HB_E_NCCL, std::string(what) + ": " + api().GetErrorString(rc));
}

}  // namespace

void nccl_unique_id(void* out128) { check(api().GetUniqueId(out128), "ncclGetUniqueId"); }

void* nccl_comm_init(int world, int rank, const void* id128) {
  // ncclCommInitRank(ncclComm_t*, int nranks, ncclUniqueId commId /* by value */, int rank)
  typedef int (*InitFn)(void**, int, IdByValue, int);
  InitFn init = (InitFn)api().CommInitRank;
  IdByValue id;
  memcpy(id.internal, id128, NCCL_ID_BYTES);
  void* comm = nullptr;
  check(init(&comm, world, id, rank), "ncclCommInitRank");
  return comm;
}

void nccl_comm_destroy(void* comm) {
  if (comm) api().CommDestroy(comm);
}

void nccl_allreduce_sum(void* comm, double* buf, size_t count, void* stream) {
  // ncclAllReduce(sendbuff, recvbuff, count, ncclDouble = 8, ncclSum = 0, comm, stream)
  check(api().AllReduce(buf, buf, count, 8, 0, comm, stream), "ncclAllReduce");
}

void nccl_group_start() { check(api().GroupStart(), "ncclGroupStart"); }
void nccl_group_end() { check(api().GroupEnd(), "ncclGroupEnd"); }

}  // namespace hb
