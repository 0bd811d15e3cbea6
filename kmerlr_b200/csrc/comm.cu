// comm.cu -- sample sharding over the GPUs of one box: one process per GPU, NCCL over
// NVLink.  NCCL is bound at run time with dlopen so that the library shares the libnccl the host
// process already loaded (torch's bundled one under bench.py) and has no link-time dependency.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace kl {

namespace {

struct Nccl {
  void *lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
Nccl g_nccl;

void load_nccl() {
  if (g_nccl.lib) return;
  const char *names[] = {getenv("KMERLR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    if (!nm || !*nm) continue;
    g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) fail(KMERLR_ERR_CUDA, std::string("cannot load libnccl.so.2: ") + dlerror());
#define KL_SYM(name)                                                                      \
  g_nccl.name = (decltype(g_nccl.name))dlsym(g_nccl.lib, "nccl" #name);                   \
  if (!g_nccl.name) fail(KMERLR_ERR_CUDA, "libnccl lacks nccl" #name)
  KL_SYM(GetUniqueId); KL_SYM(CommInitRank); KL_SYM(CommDestroy); KL_SYM(AllReduce); KL_SYM(AllGather);
  KL_SYM(GetErrorString);
#undef KL_SYM
}

#define KL_NCCL(expr)                                                                             \
  do {                                                                                            \
    ncclResult_t _r = (expr);                                                                     \
    if (_r != ncclSuccess) fail(KMERLR_ERR_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
  } while (0)

ncclComm_t comm() { return (ncclComm_t)ctx().comm; }

}  // namespace

void comm_unique_id(void *id128) {
  load_nccl();
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  KL_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
}

void comm_init(int rank, int world, const void *id128) {
  require_ready();
  KL_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm_init: bad rank / world");
  if (ctx().comm) comm_destroy();
  ctx().rank = rank; ctx().world = world;
  if (world == 1) return;
  load_nccl();
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  KL_NCCL(g_nccl.CommInitRank(&c, world, id, rank));
  ctx().comm = c;
}

void comm_destroy() {
  if (ctx().comm) {
    g_nccl.CommDestroy(comm());
    ctx().comm = nullptr;
  }
  ctx().rank = 0; ctx().world = 1;
}

static void need_comm() {
  if (!ctx().comm) fail(KMERLR_ERR_ARG, "sharded call without a communicator (kmerlr_comm_init)");
}

void comm_allreduce_sum_f64(double *dev, int64_t count) {
  if (ctx().world == 1) return;
  need_comm();
  KL_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, comm(), ctx().stream));
}
void comm_allreduce_sum_i64(int64_t *dev, int64_t count) {
  if (ctx().world == 1) return;
  need_comm();
  KL_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclInt64, ncclSum, comm(), ctx().stream));
}
void comm_allreduce_max_f64(double *dev, int64_t count) {
  if (ctx().world == 1) return;
  need_comm();
  KL_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclMax, comm(), ctx().stream));
}
void comm_allreduce_max_u8(uint8_t *dev, int64_t count) {
  if (ctx().world == 1) return;
  need_comm();
  KL_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclUint8, ncclMax, comm(), ctx().stream));
}
void comm_allgather_f64(const double *dev_in, double *dev_out, int64_t count_per_rank) {
  if (ctx().world == 1) {
    KL_CUDA(cudaMemcpyAsync(dev_out, dev_in, (size_t)count_per_rank * sizeof(double), cudaMemcpyDeviceToDevice,
                            ctx().stream));
    return;
  }
  need_comm();
  KL_NCCL(g_nccl.AllGather(dev_in, dev_out, (size_t)count_per_rank, ncclFloat64, comm(), ctx().stream));
}

}  // namespace kl
