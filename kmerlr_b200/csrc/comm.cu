// comm.cu -- sample sharding over the GPUs of one box: one process per GPU, NCCL over
// NVLink.  NCCL is bound at run time with dlopen so that the library shares the libnccl the host
// process already loaded (torch's bundled one under bench.py) and has no link-time dependency.
#include <cooperative_groups.h>
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
unsigned long long *g_my_sym = nullptr;           // this rank's all-reduce buffer (IN | OUT)
unsigned long long *g_peer_sym[PEER_MAX_WORLD] = {nullptr};
PeerBox *g_peer_dev = nullptr;

// ---- all-reduce (int64 sum) over peer memory, two-shot, in one cooperative launch --------------------------
// Every rank copies its vector into its IN buffer; barrier; rank r adds slice r of all IN buffers (remote reads
// over NVLink) and stores the sums into slice r of every rank's OUT buffer (remote writes); barrier; every rank
// copies its OUT buffer back.  Integer sums: the same bits on every rank, whatever the order.  A barrier = every
// rank writes its sequence number into its flag in every mailbox and waits for the flags in its own.
__global__ void __launch_bounds__(256) p2p_allreduce_kernel(const PeerBox *pb, unsigned long long *data, int64_t count) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  const int me = pb->rank, world = pb->world;
  PeerMail *mine = pb->box[me];
  const unsigned long long seq = mine->ar_seq + 1ull;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  unsigned long long *in = pb->sym[me], *out = in + PEER_AR_CAP;
  // 16-byte copies where the vector allows it (count even, data 16-byte aligned like the buffers)
  const int64_t pairs = ((count & 1) || ((unsigned long long)data & 15ull)) ? 0 : count >> 1;
  if (pairs) {
    const ulonglong2 *s2 = reinterpret_cast<const ulonglong2 *>(data);
    ulonglong2 *d2 = reinterpret_cast<ulonglong2 *>(in);
    for (int64_t j = tid; j < pairs; j += nth) d2[j] = s2[j];
  } else {
    for (int64_t j = tid; j < count; j += nth) in[j] = data[j];
  }
  // one system-scope fence per block, after the block's stores (cumulative over the block barrier)
  auto publish = [&]() {
    __syncthreads();
    if (threadIdx.x == 0) __threadfence_system();
  };
  auto barrier = [&](int b) {
    publish();
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x < (unsigned)world) {
      const int t = threadIdx.x;
      *(volatile unsigned long long *)&pb->box[t]->ar_flag[b][me] = seq;
      volatile unsigned long long *f = &mine->ar_flag[b][t];
      const long long t0 = clock64();
      while (*f < seq)
        if (clock64() - t0 > 6000000000LL) { mine->ar_error = 1u; break; }       // ~3 s: a peer is gone
      __threadfence_system();
    }
    grid.sync();
  };
  barrier(0);
  // my slice [lo, hi): four elements per thread and step, all remote loads of a step in flight together
  const int64_t per = (count + world - 1) / world, lo = (int64_t)me * per, hi = lo + per < count ? lo + per : count;
  unsigned long long *symp[PEER_MAX_WORLD];
#pragma unroll
  for (int r = 0; r < PEER_MAX_WORLD; r++) symp[r] = r < world ? pb->sym[r] : nullptr;
  for (int64_t j0 = lo + 4 * tid; j0 < hi; j0 += 4 * nth) {
    unsigned long long s[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
    for (int r = 0; r < PEER_MAX_WORLD; r++) {
      if (r < world) {
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (j0 + q < hi) s[q] += __ldcv(symp[r] + j0 + q);
      }
    }
#pragma unroll
    for (int r = 0; r < PEER_MAX_WORLD; r++) {
      if (r < world) {
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (j0 + q < hi) symp[r][PEER_AR_CAP + j0 + q] = s[q];
      }
    }
  }
  barrier(1);
  if (pairs) {
    const ulonglong2 *s2 = reinterpret_cast<const ulonglong2 *>(out);
    ulonglong2 *d2 = reinterpret_cast<ulonglong2 *>(data);
    for (int64_t j = tid; j < pairs; j += nth) d2[j] = __ldcv(s2 + j);
  } else {
    for (int64_t j = tid; j < count; j += nth) data[j] = __ldcv(out + j);
  }
  if (tid == 0) mine->ar_seq = seq;
}

void peer_teardown() {
  for (int r = 0; r < PEER_MAX_WORLD; r++) {
    if (g_peer_mail[r] && g_peer_mail[r] != g_my_mail) cudaIpcCloseMemHandle(g_peer_mail[r]);
    g_peer_mail[r] = nullptr;
  }
  for (int r = 0; r < PEER_MAX_WORLD; r++) {
    if (g_peer_sym[r] && g_peer_sym[r] != g_my_sym) cudaIpcCloseMemHandle(g_peer_sym[r]);
    g_peer_sym[r] = nullptr;
  }
  if (g_my_mail) { cudaFree(g_my_mail); g_my_mail = nullptr; }
  if (g_my_sym) { cudaFree(g_my_sym); g_my_sym = nullptr; }
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
  constexpr size_t HB = 2 * sizeof(cudaIpcMemHandle_t) + 8 + IDB;   // mailbox handle, all-reduce buffer handle, ok, device id
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
  if (ok && cudaMalloc((void **)&g_my_sym, 2 * PEER_AR_CAP * sizeof(unsigned long long)) != cudaSuccess) {
    cudaGetLastError(); g_my_sym = nullptr; ok = 0;
  }
  cudaIpcMemHandle_t mine, mine2;
  memset(&mine, 0, sizeof(mine)); memset(&mine2, 0, sizeof(mine2));
  if (ok) {
    cudaMemset(g_my_mail, 0, sizeof(PeerMail));
    if (cudaIpcGetMemHandle(&mine, g_my_mail) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    if (ok && cudaIpcGetMemHandle(&mine2, g_my_sym) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  }
  std::vector<unsigned char> h(HB, 0), all(HB * world, 0);
  memcpy(h.data(), &mine, sizeof(mine));
  memcpy(h.data() + HB - sizeof(mine2), &mine2, sizeof(mine2));
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
      if (r == rank) { g_peer_mail[r] = g_my_mail; g_peer_sym[r] = g_my_sym; continue; }
      cudaIpcMemHandle_t hr, hr2;
      memcpy(&hr, all.data() + r * HB, sizeof(hr));
      memcpy(&hr2, all.data() + (r + 1) * HB - sizeof(hr2), sizeof(hr2));
      void *p = nullptr, *p2 = nullptr;
      if (cudaIpcOpenMemHandle(&p, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); good = false; }
      g_peer_mail[r] = (PeerMail *)p;
      if (good && cudaIpcOpenMemHandle(&p2, hr2, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); good = false; }
      g_peer_sym[r] = (unsigned long long *)p2;
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
  for (int r = 0; r < world; r++) { hb.box[r] = g_peer_mail[r]; hb.sym[r] = g_peer_sym[r]; }
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
void comm_allreduce_sum_i64_fast(int64_t *dev, int64_t count) {
  if (ctx().world == 1) return;
  need_comm();
  // Off by default: measured at 2 GPUs it loses to NCCL (351 KB: 49 vs 19 us, 5.6 MB: 130 vs 92 us) -- four grid
  // barriers and two cross-GPU flag rounds cost more than NCCL's LL protocol at these sizes.  kmerlr_option
  // ("p2p_allreduce", 1) enables it (same bits as NCCL: integer sums).
  if (!(ctx().p2p_allreduce && ctx().peer && ctx().p2p_ok && ctx().coop_supported && count > 0 && count <= PEER_AR_CAP)) {
    comm_allreduce_sum_i64(dev, count);
    return;
  }
  int blocks = ctx().sm_count;                                   // one block per SM: resident as a whole
  if ((int64_t)blocks * 256 > count) blocks = (int)((count + 255) / 256);
  const PeerBox *pb = ctx().peer;
  unsigned long long *data = (unsigned long long *)dev;
  void *args[] = {(void *)&pb, (void *)&data, (void *)&count};
  if (ctx().profiling) profile_begin("p2p_allreduce_kernel");
  KL_CUDA(cudaLaunchCooperativeKernel((const void *)p2p_allreduce_kernel, dim3((unsigned)blocks), dim3(256), args, 0, ctx().stream));
  if (ctx().profiling) profile_end();
  ctx().launches++;
}
// a peer-memory all-reduce that timed out leaves garbage behind: callers check at their next host sync
void comm_check_peer_errors() {
  if (!g_my_mail || ctx().world == 1) return;
  unsigned int e = 0;
  KL_CUDA(cudaMemcpyAsync(&e, &g_my_mail->ar_error, sizeof(e), cudaMemcpyDeviceToHost, ctx().stream));
  KL_CUDA(cudaStreamSynchronize(ctx().stream));
  if (e) fail(KMERLR_ERR_CUDA, "a rank did not show up at the gradient all-reduce over peer memory");
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
