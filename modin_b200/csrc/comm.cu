// comm.cu — the NCCL entry points of the C ABI (comm_init, allreduce, reduce_scatter, allgather,
// broadcast, alltoallv).
//
// The reference never issues a collective: its cross-partition data movement is "gather every block of an axis into
// one task" (axis_partition.py:445-452), "hand one object to many tasks" (pm.py:443-494) and the range-partitioning
// shuffle (pm.py:1937-2052).  On one box of B200s those become NCCL collectives over NVLink 5 / NVSwitch, issued from
// this library on the caller's stream, on the buffers the kernels of this library produce:
//   TreeReduce combine            -> mb200_comm_allreduce on the W-vector
//   GroupByReduce reduce (dense)  -> mb200_comm_reduce_scatter on the [keys][values] accumulator arrays
//   broadcast merge               -> mb200_comm_allgather of the dim shards
//   range-partitioning shuffle    -> mb200_comm_alltoallv (grouped ncclSend / ncclRecv) of raw rows / partial tables
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 -- the copy PyTorch ships is already mapped into the process,
// else a path handed to mb200_comm_load), so the library has no link-time dependency on it and builds on a box
// without NCCL.  The communicator is this library's own: rank 0 obtains the unique id (mb200_comm_unique_id), the host
// side carries its 128 bytes to the other ranks by whatever control channel it has, every rank calls
// mb200_comm_init_rank.
#include <dlfcn.h>

#include "common.cuh"

namespace {

typedef struct {
  char internal[128];
} NcclUniqueId;
typedef void* NcclComm;

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
};
NcclApi g_nccl;

// ncclDataType_t / ncclRedOp_t values (nccl.h; stable across NCCL 2.x)
constexpr int kNcclUint8 = 1, kNcclInt64 = 4, kNcclFloat64 = 8;
constexpr int kNcclSum = 0, kNcclMax = 2, kNcclMin = 3;

int nccl_fail(const char* what, int rc) {
  snprintf(mb200::g_err, sizeof(mb200::g_err), "%s: NCCL error %d (%s)", what, rc,
           g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  return 3;
}

int nccl_dtype(int dtype, int* out, int* bytes) {
  switch (dtype) {
    case MB200_F64: *out = kNcclFloat64; *bytes = 8; return 0;
    case MB200_I64: *out = kNcclInt64; *bytes = 8; return 0;
    case MB200_U8: *out = kNcclUint8; *bytes = 1; return 0;
  }
  return mb200::fail("mb200_comm", "dtype must be MB200_F64 / MB200_I64 / MB200_U8");
}

int nccl_op(int op, int* out) {
  switch (op) {
    case MB200_COMM_SUM: *out = kNcclSum; return 0;
    case MB200_COMM_MIN: *out = kNcclMin; return 0;
    case MB200_COMM_MAX: *out = kNcclMax; return 0;
  }
  return mb200::fail("mb200_comm", "op must be MB200_COMM_SUM / MIN / MAX");
}

}  // namespace

struct mb200_comm {
  NcclComm comm;
  int rank, nranks;
};

using namespace mb200;

#define MB_NCCL(call)                              \
  do {                                             \
    int _r = (call);                               \
    if (_r != 0) return nccl_fail(#call, _r);      \
  } while (0)

extern "C" int mb200_comm_load(const char* libnccl_path) {
  if (g_nccl.handle) return 0;
  void* h = nullptr;
  if (libnccl_path && libnccl_path[0]) h = dlopen(libnccl_path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail("mb200_comm_load", "libnccl.so.2 not found (pass its path)");
#define MB_SYM(field, name)                                                   \
  do {                                                                        \
    *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);                \
    if (!g_nccl.field) return fail("mb200_comm_load", "symbol missing: " name); \
  } while (0)
  MB_SYM(GetUniqueId, "ncclGetUniqueId");
  MB_SYM(CommInitRank, "ncclCommInitRank");
  MB_SYM(CommDestroy, "ncclCommDestroy");
  MB_SYM(GetErrorString, "ncclGetErrorString");
  MB_SYM(AllReduce, "ncclAllReduce");
  MB_SYM(ReduceScatter, "ncclReduceScatter");
  MB_SYM(AllGather, "ncclAllGather");
  MB_SYM(Broadcast, "ncclBroadcast");
  MB_SYM(Send, "ncclSend");
  MB_SYM(Recv, "ncclRecv");
  MB_SYM(GroupStart, "ncclGroupStart");
  MB_SYM(GroupEnd, "ncclGroupEnd");
#undef MB_SYM
  g_nccl.handle = h;
  return 0;
}

extern "C" int mb200_comm_unique_id(void* out128) {
  if (!g_nccl.handle) return fail("mb200_comm_unique_id", "call mb200_comm_load first");
  if (!out128) return fail("mb200_comm_unique_id", "null output");
  NcclUniqueId id;
  MB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out128, id.internal, 128);
  return 0;
}

extern "C" int mb200_comm_init_rank(mb200_comm** comm, int nranks, const void* id128, int rank) {
  if (!g_nccl.handle) return fail("mb200_comm_init_rank", "call mb200_comm_load first");
  if (!comm || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail("mb200_comm_init_rank", "bad arguments");
  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;  // sm_100 device selected on this thread
  NcclUniqueId id;
  memcpy(id.internal, id128, 128);
  NcclComm c = nullptr;
  MB_NCCL(g_nccl.CommInitRank(&c, nranks, id, rank));
  mb200_comm* m = new mb200_comm();
  m->comm = c;
  m->rank = rank;
  m->nranks = nranks;
  *comm = m;
  return 0;
}

extern "C" int mb200_comm_destroy(mb200_comm* comm) {
  if (!comm) return 0;
  if (g_nccl.handle && comm->comm) g_nccl.CommDestroy(comm->comm);
  delete comm;
  return 0;
}

extern "C" int mb200_comm_allreduce(mb200_comm* comm, const void* send, void* recv, int64_t count, int dtype, int op,
                                    mb200_stream_t stream) {
  if (!comm) return fail("mb200_comm_allreduce", "null communicator");
  if (count < 0) return fail("mb200_comm_allreduce", "negative count");
  if (count == 0) return 0;
  int dt, bytes, ro;
  if (int rc = nccl_dtype(dtype, &dt, &bytes)) return rc;
  if (int rc = nccl_op(op, &ro)) return rc;
  MB_NCCL(g_nccl.AllReduce(send, recv, (size_t)count, dt, ro, comm->comm, (cudaStream_t)stream));
  return 0;
}

extern "C" int mb200_comm_reduce_scatter(mb200_comm* comm, const void* send, void* recv, int64_t recvcount, int dtype,
                                         int op, mb200_stream_t stream) {
  if (!comm) return fail("mb200_comm_reduce_scatter", "null communicator");
  if (recvcount < 0) return fail("mb200_comm_reduce_scatter", "negative count");
  if (recvcount == 0) return 0;
  int dt, bytes, ro;
  if (int rc = nccl_dtype(dtype, &dt, &bytes)) return rc;
  if (int rc = nccl_op(op, &ro)) return rc;
  MB_NCCL(g_nccl.ReduceScatter(send, recv, (size_t)recvcount, dt, ro, comm->comm, (cudaStream_t)stream));
  return 0;
}

extern "C" int mb200_comm_allgather(mb200_comm* comm, const void* send, void* recv, int64_t sendcount, int dtype,
                                    mb200_stream_t stream) {
  if (!comm) return fail("mb200_comm_allgather", "null communicator");
  if (sendcount < 0) return fail("mb200_comm_allgather", "negative count");
  if (sendcount == 0) return 0;
  int dt, bytes;
  if (int rc = nccl_dtype(dtype, &dt, &bytes)) return rc;
  MB_NCCL(g_nccl.AllGather(send, recv, (size_t)sendcount, dt, comm->comm, (cudaStream_t)stream));
  return 0;
}

extern "C" int mb200_comm_broadcast(mb200_comm* comm, void* buf, int64_t count, int dtype, int root,
                                    mb200_stream_t stream) {
  if (!comm) return fail("mb200_comm_broadcast", "null communicator");
  if (count < 0 || root < 0 || root >= comm->nranks) return fail("mb200_comm_broadcast", "bad count / root");
  if (count == 0) return 0;
  int dt, bytes;
  if (int rc = nccl_dtype(dtype, &dt, &bytes)) return rc;
  MB_NCCL(g_nccl.Broadcast(buf, buf, (size_t)count, dt, root, comm->comm, (cudaStream_t)stream));
  return 0;
}

extern "C" int mb200_comm_alltoallv(mb200_comm* comm, const void* send, const int64_t* sendcounts,
                                    const int64_t* sdispls, void* recv, const int64_t* recvcounts, const int64_t* rdispls,
                                    int elem_bytes, mb200_stream_t stream) {
  if (!comm) return fail("mb200_comm_alltoallv", "null communicator");
  if (!sendcounts || !sdispls || !recvcounts || !rdispls || elem_bytes < 1)
    return fail("mb200_comm_alltoallv", "bad arguments");
  const char* s = static_cast<const char*>(send);
  char* r = static_cast<char*>(recv);
  MB_NCCL(g_nccl.GroupStart());
  for (int p = 0; p < comm->nranks; ++p) {
    if (sendcounts[p] > 0) {
      int rc = g_nccl.Send(s + (size_t)sdispls[p] * elem_bytes, (size_t)sendcounts[p] * elem_bytes, kNcclUint8, p,
                           comm->comm, (cudaStream_t)stream);
      if (rc != 0) {
        g_nccl.GroupEnd();
        return nccl_fail("ncclSend", rc);
      }
    }
    if (recvcounts[p] > 0) {
      int rc = g_nccl.Recv(r + (size_t)rdispls[p] * elem_bytes, (size_t)recvcounts[p] * elem_bytes, kNcclUint8, p,
                           comm->comm, (cudaStream_t)stream);
      if (rc != 0) {
        g_nccl.GroupEnd();
        return nccl_fail("ncclRecv", rc);
      }
    }
  }
  MB_NCCL(g_nccl.GroupEnd());
  return 0;
}
