// comm.cu -- sample sharding over the GPUs of one box: one process per GPU, NCCL over
// NVLink.  NCCL is bound at run time with dlopen so that the library shares the libnccl the host
// process already loaded (torch's bundled one under bench.py) and has no link-time dependency.
#include <dlfcn.h>
#include <nccl.h>

#include <vector>

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

// ---- peer-memory mailboxes (CUDA IPC over NVLink) ---------------------------------------------------
PeerMail *g_my_mail = nullptr;                    // this rank's mailbox (cudaMalloc)
PeerMail *g_peer_mail[PEER_MAX_WORLD] = {nullptr};
PeerBox *g_peer_dev = nullptr;

void peer_teardown() {
  for (int r = 0; r < PEER_MAX_WORLD; r++) {
    if (g_peer_mail[r] && g_peer_mail[r] != g_my_mail) cudaIpcCloseMemHandle(g_peer_mail[r]);
    g_peer_mail[r] = nullptr;
  }
  if (g_my_mail) { cudaFree(g_my_mail); g_my_mail = nullptr; }
  if (g_peer_dev) { cudaFree(g_peer_dev); g_peer_dev = nullptr; }
  ctx().peer = nullptr;
}

// every rank allocates its mailbox, the IPC handles go round with an all-gather, every rank maps the
// others'.  Any failure on any rank leaves ctx().peer == nullptr everywhere and the NCCL collectives in
// charge: a rank that failed locally still takes part in both agreement collectives (ok = 0), so nobody is
// left waiting inside comm_init.  The exchange is also refused when two ranks of the communicator sit on
// the same device: their kernels would spin on each other's flags without being guaranteed to run at the
// same time (B200_PROFILING.md: Xid 109).
void peer_setup() {
  const int world = ctx().world, rank = ctx().rank;
  if (world < 2 || world > PEER_MAX_WORLD) return;
  constexpr size_t IDB = 32;                                   // PCI bus id of the rank's device
  constexpr size_t HB = sizeof(cudaIpcMemHandle_t) + 8 + IDB;
  unsigned char *din = nullptr, *dout = nullptr;
  int *dflag = nullptr;
  int ok = 1;
  if (cudaMalloc((void **)&din, HB) != cudaSuccess || cudaMalloc((void **)&dout, HB * world) != cudaSuccess ||
      cudaMalloc((void **)&dflag, sizeof(int)) != cudaSuccess) {
    // without these three buffers the collectives below cannot run on this rank either: nothing was entered yet,
    // and the other ranks would wait -- this is the one failure that has to be fatal
    cudaGetLastError();
    fail(KMERLR_ERR_CUDA, "comm_init: out of device memory for the peer-memory handshake");
  }
  if (cudaMalloc((void **)&g_my_mail, sizeof(PeerMail)) != cudaSuccess) { cudaGetLastError(); g_my_mail = nullptr; ok = 0; }
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) {
    cudaMemset(g_my_mail, 0, sizeof(PeerMail));
    if (cudaIpcGetMemHandle(&mine, g_my_mail) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  }
  std::vector<unsigned char> h(HB, 0), all(HB * world, 0);
  memcpy(h.data(), &mine, sizeof(mine));
  h[sizeof(mine)] = (unsigned char)ok;
  char busid[IDB] = {0};
  if (cudaDeviceGetPCIBusId(busid, (int)IDB, ctx().device) != cudaSuccess) { cudaGetLastError(); snprintf(busid, IDB, "dev%d", ctx().device); }
  memcpy(h.data() + sizeof(mine) + 8, busid, IDB);
  bool good = cudaMemcpy(din, h.data(), HB, cudaMemcpyHostToDevice) == cudaSuccess;
  // agreement step 1 (always entered): handles, ok bytes and device ids of all ranks
  KL_NCCL(g_nccl.AllGather(din, dout, HB, ncclUint8, comm(), ctx().stream));
  KL_CUDA(cudaStreamSynchronize(ctx().stream));
  good = good && cudaMemcpy(all.data(), dout, HB * world, cudaMemcpyDeviceToHost) == cudaSuccess;
  for (int r = 0; r < world; r++) good = good && all[r * HB + sizeof(mine)] == 1;
  for (int r = 0; r < world && good; r++)
    for (int q = r + 1; q < world; q++)
      if (!memcmp(all.data() + r * HB + sizeof(mine) + 8, all.data() + q * HB + sizeof(mine) + 8, IDB)) good = false;
  if (good) {
    for (int r = 0; r < world && good; r++) {
      if (r == rank) { g_peer_mail[r] = g_my_mail; continue; }
      cudaIpcMemHandle_t hr;
      memcpy(&hr, all.data() + r * HB, sizeof(hr));
      void *p = nullptr;
      if (cudaIpcOpenMemHandle(&p, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); good = false; }
      g_peer_mail[r] = (PeerMail *)p;
    }
  }
  // agreement step 2 (always entered): a rank that failed to map makes everybody fall back
  int hflag = good ? 1 : 0;
  if (cudaMemcpy(dflag, &hflag, sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); cudaMemset(dflag, 0, sizeof(int)); }
  KL_NCCL(g_nccl.AllReduce(dflag, dflag, 1, ncclInt32, ncclMin, comm(), ctx().stream));
  KL_CUDA(cudaStreamSynchronize(ctx().stream));
  if (cudaMemcpy(&hflag, dflag, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); hflag = 0; }
  cudaFree(din); cudaFree(dout); cudaFree(dflag);
  if (!hflag) { peer_teardown(); return; }
  PeerBox hb{};
  for (int r = 0; r < world; r++) hb.box[r] = g_peer_mail[r];
  hb.rank = rank; hb.world = world;
  if (cudaMalloc((void **)&g_peer_dev, sizeof(PeerBox)) != cudaSuccess ||
      cudaMemcpy(g_peer_dev, &hb, sizeof(PeerBox), cudaMemcpyHostToDevice) != cudaSuccess) {
    // (after the agreement: this rank alone would fall back, so this late failure has to be fatal as well)
    cudaGetLastError();
    fail(KMERLR_ERR_CUDA, "comm_init: out of device memory for the mailbox table");
  }
  ctx().peer = g_peer_dev;
}

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
  const char *e = getenv("KMERLR_P2P");
  if (!(e && *e == '0')) peer_setup();
}

void comm_destroy() {
  peer_teardown();
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
  if (ctx().profiling) profile_begin("nccl_allreduce_sum_f64");
  KL_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, comm(), ctx().stream));
  if (ctx().profiling) profile_end();
}
void comm_allreduce_sum_i64(int64_t *dev, int64_t count) {
  if (ctx().world == 1) return;
  need_comm();
  if (ctx().profiling) profile_begin("nccl_allreduce_sum_i64");
  KL_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclInt64, ncclSum, comm(), ctx().stream));
  if (ctx().profiling) profile_end();
}
void comm_allreduce_max_f64(double *dev, int64_t count) {
  if (ctx().world == 1) return;
  need_comm();
  if (ctx().profiling) profile_begin("nccl_allreduce_max_f64");
  KL_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclMax, comm(), ctx().stream));
  if (ctx().profiling) profile_end();
}
void comm_allreduce_max_u8(uint8_t *dev, int64_t count) {
  if (ctx().world == 1) return;
  need_comm();
  if (ctx().profiling) profile_begin("nccl_allreduce_max_u8");
  KL_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclUint8, ncclMax, comm(), ctx().stream));
  if (ctx().profiling) profile_end();
}
void comm_allgather_bytes(const void *dev_in, void *dev_out, int64_t bytes_per_rank) {
  if (ctx().world == 1) {
    KL_CUDA(cudaMemcpyAsync(dev_out, dev_in, (size_t)bytes_per_rank, cudaMemcpyDeviceToDevice, ctx().stream));
    return;
  }
  need_comm();
  if (ctx().profiling) profile_begin("nccl_allgather_bytes");
  KL_NCCL(g_nccl.AllGather(dev_in, dev_out, (size_t)bytes_per_rank, ncclUint8, comm(), ctx().stream));
  if (ctx().profiling) profile_end();
}
void comm_allgather_f64(const double *dev_in, double *dev_out, int64_t count_per_rank) {
  if (ctx().world == 1) {
    KL_CUDA(cudaMemcpyAsync(dev_out, dev_in, (size_t)count_per_rank * sizeof(double), cudaMemcpyDeviceToDevice,
                            ctx().stream));
    return;
  }
  need_comm();
  if (ctx().profiling) profile_begin("nccl_allgather_f64");
  KL_NCCL(g_nccl.AllGather(dev_in, dev_out, (size_t)count_per_rank, ncclFloat64, comm(), ctx().stream));
  if (ctx().profiling) profile_end();
}

}  // namespace kl
