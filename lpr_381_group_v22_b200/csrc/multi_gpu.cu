// multi_gpu.cu -- branch & bound over the GPUs of one box, driven from INSIDE the library: the C#
// callers are single-process (Program.cs:385-389 branch & bound simplex, :444-468 knapsack), so the drop-in reaches
// more than one GPU only if the C ABI does it itself.  One host thread per device, one NCCL communicator per device
// (ncclCommInitAll, cached for the life of the process); the open-node pool is partitioned across the devices and
// every round the ranks
//   1. all-reduce(MAX) one short fp64 vector [incumbent value, "my incumbent changed", error, count_0 .. count_{N-1}]
//      over NVLink (the north star's "allreduce-min" on the negated objective; each rank fills its own count slot,
//      MAX against the others' zeros returns it);
//   2. only when some incumbent changed: the holders of the best value compare DFS keys (shared host memory, the
//      ranks are threads of one process) and everybody adopts the DFS-first one, which makes the answer independent
//      of the GPU count, of the batch size and of where the rounds are cut;
//   3. steal: ranks at or below the low-water mark receive half of the fullest rank's pool, node records travelling
//      device to device with ncclSend / ncclRecv inside one ncclGroupStart / ncclGroupEnd;
//   4. expand their pools for one time slice (lpr_bb_run_timed / lpr_knap_run_timed).
// NCCL is resolved with dlopen (libnccl.so.2: the copy torch already mapped when the caller is a Python process, the
// system one otherwise) so that single-GPU users need no NCCL at all; with n_gpus > 1 a missing NCCL is an error
// (LPR_E_NCCL), never a fallback.  The torch.distributed driver (lpr_381_group_v22_b200/distributed.py, one process
// per GPU) stays as the second way in; both run the same pools through the same C ABI.
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace lpr {
void bb_set_prealloc_share(int pools_per_device);  // bb.cu: pools created from now on carve 1/n of LPR_BB_PREALLOC_MB
}

namespace {

using namespace lpr;

// ---- NCCL through dlopen (prototypes of nccl.h 2.x; the ABI of these entry points is stable) ---------------------
typedef struct ncclComm* ncclComm_t;
enum { kNcclUint8 = 1, kNcclFloat64 = 8, kNcclMax = 2 };
struct NcclApi {
  void* lib = nullptr;
  int version = 0;
  std::string error;
  int (*GetVersion)(int*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
};

NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("LPR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) {
      api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "unknown");
      return;
    }
#define SYM(field, name)                                                   \
  *(void**)(&api.field) = dlsym(api.lib, name);                            \
  if (!api.field) {                                                        \
    api.error = std::string("libnccl lacks ") + name;                      \
    return;                                                                \
  }
    SYM(GetVersion, "ncclGetVersion")
    SYM(GetErrorString, "ncclGetErrorString")
    SYM(CommInitAll, "ncclCommInitAll")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
#undef SYM
    api.GetVersion(&api.version);
  });
  return api;
}

struct CommSet {
  std::vector<int> devices;
  std::vector<ncclComm_t> comms;
};
std::mutex g_comm_mutex;
std::map<std::vector<int>, CommSet*> g_comm_cache;

int get_comms(const std::vector<int>& devices, CommSet** out) {
  NcclApi& api = nccl_api();
  if (!api.error.empty()) return fail(LPR_E_NCCL, "%s", api.error.c_str());
  std::lock_guard<std::mutex> lock(g_comm_mutex);
  auto it = g_comm_cache.find(devices);
  if (it != g_comm_cache.end()) {
    *out = it->second;
    return LPR_OK;
  }
  CommSet* cs = new CommSet();
  cs->devices = devices;
  cs->comms.resize(devices.size());
  const int rc = api.CommInitAll(cs->comms.data(), (int)devices.size(), devices.data());
  if (rc != 0) {
    delete cs;
    return fail(LPR_E_NCCL, "ncclCommInitAll over %d devices failed: %s", (int)devices.size(), api.GetErrorString(rc));
  }
  g_comm_cache[devices] = cs;
  *out = cs;
  return LPR_OK;
}

// ---- a pool seen through the C ABI ----------------------------------------------------------------------------------
struct Incumbent {
  bool has = false;
  double z = -INFINITY;
  std::vector<uint8_t> key;     // DFS path, one byte per branch (0 first child, 1 second)
  std::vector<double> payload;  // x (B&B simplex) or the chosen-item vector (knapsack)
};
int key_order(const std::vector<uint8_t>& a, const std::vector<uint8_t>& b) {
  const size_t n = std::min(a.size(), b.size());
  for (size_t i = 0; i < n; i++)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  if (a.size() == b.size()) return 0;
  return a.size() < b.size() ? -1 : 1;
}
bool better(const Incumbent& a, const Incumbent& b) {  // is a ahead of b?
  if (!b.has) return a.has;
  if (!a.has) return false;
  if (a.z != b.z) return a.z > b.z;
  return key_order(a.key, b.key) < 0;
}

struct Pool {
  virtual ~Pool() {}
  virtual int create(int device, bool with_root) = 0;
  virtual void destroy() = 0;
  virtual int64_t open_count() = 0;
  virtual int run(int64_t max_nodes, double max_seconds, int64_t* done) = 0;
  virtual int get_incumbent(Incumbent* inc) = 0;
  virtual int set_incumbent(const Incumbent& inc) = 0;
  virtual int export_nodes(int max_nodes, void* dbuf, int64_t cap, int64_t* bytes, int* n) = 0;
  virtual int import_nodes(const void* dbuf, int64_t bytes) = 0;
  virtual int keep_stride(int, int) { return LPR_OK; }
  virtual bool replicated_root() const { return false; }
  virtual int64_t pivots() { return 0; }
  virtual int64_t depth_overflow() { return 0; }
};

struct BBProblem {
  int rows, cols, n_vars, prune;
  const double* root;
};
struct BBPoolC : Pool {
  const BBProblem& p;
  lpr_bb* h = nullptr;
  int64_t piv = 0;
  explicit BBPoolC(const BBProblem& pp) : p(pp) {}
  // every rank is given the same root: all expand it identically and keep every world-th node (no transfer)
  int create(int device, bool) override { return lpr_bb_create(device, p.rows, p.cols, p.root, p.n_vars, p.prune, &h); }
  void destroy() override {
    lpr_bb_destroy(h);
    h = nullptr;
  }
  int64_t open_count() override {
    int64_t n = 0;
    lpr_bb_open_count(h, &n);
    return n;
  }
  int run(int64_t max_nodes, double max_seconds, int64_t* done) override {
    int64_t pv = 0;
    const int rc = lpr_bb_run_timed(h, max_nodes, max_seconds, done, &pv);
    piv += pv;
    return rc;
  }
  int get_incumbent(Incumbent* inc) override {
    int has = 0, klen = 0;
    inc->payload.assign(p.n_vars, 0.0);
    int rc = lpr_bb_get_incumbent(h, &has, &inc->z, inc->payload.data(), nullptr, &klen);
    if (rc) return rc;
    inc->has = has != 0;
    inc->key.clear();
    if (has && klen > 0) {
      std::vector<int> k(klen);
      int cap = klen;
      if ((rc = lpr_bb_get_incumbent(h, &has, &inc->z, inc->payload.data(), k.data(), &cap))) return rc;
      inc->key.assign(k.begin(), k.end());
    }
    return LPR_OK;
  }
  int set_incumbent(const Incumbent& inc) override {
    std::vector<int> k(inc.key.begin(), inc.key.end());
    return lpr_bb_set_incumbent(h, inc.z, inc.payload.data(), k.empty() ? nullptr : k.data(), (int)k.size());
  }
  int export_nodes(int max_nodes, void* dbuf, int64_t cap, int64_t* bytes, int* n) override {
    return lpr_bb_export_nodes(h, max_nodes, dbuf, cap, bytes, n);
  }
  int import_nodes(const void* dbuf, int64_t bytes) override { return lpr_bb_import_nodes(h, dbuf, bytes); }
  int keep_stride(int off, int stride) override { return lpr_bb_keep_stride(h, off, stride); }
  bool replicated_root() const override { return true; }
  int64_t pivots() override { return piv; }
  int64_t depth_overflow() override {
    int64_t ovf = 0;
    lpr_bb_stats(h, nullptr, nullptr, &ovf, nullptr);
    return ovf;
  }
};

struct KnapProblem {
  double capacity;
  int n;
  const double *w, *v;
};
struct KnapPoolC : Pool {
  const KnapProblem& p;
  lpr_knap* h = nullptr;
  explicit KnapPoolC(const KnapProblem& pp) : p(pp) {}
  // every rank is given the root: all expand it identically (the level loop is deterministic) and keep every
  // world-th node, so nobody starts empty and nothing is transferred
  int create(int device, bool) override { return lpr_knap_create(device, p.capacity, p.n, p.w, p.v, &h); }
  int keep_stride(int off, int stride) override { return lpr_knap_keep_stride(h, off, stride); }
  bool replicated_root() const override { return true; }
  void destroy() override {
    lpr_knap_destroy(h);
    h = nullptr;
  }
  int64_t open_count() override {
    int64_t n = 0;
    lpr_knap_open_count(h, &n);
    return n;
  }
  int run(int64_t max_nodes, double max_seconds, int64_t* done) override {
    int st = 0;
    return lpr_knap_run_timed(h, max_nodes, max_seconds, done, &st);
  }
  int get_incumbent(Incumbent* inc) override {
    const int W = (p.n + 63) / 64;
    std::vector<uint8_t> ch(p.n);
    std::vector<uint64_t> key(W, 0);
    int bits = -1;
    int rc = lpr_knap_get_incumbent(h, &inc->z, ch.data(), key.data(), &bits);
    if (rc) return rc;
    inc->has = bits >= 0;
    inc->key.clear();
    inc->payload.assign(ch.begin(), ch.end());
    for (int i = 0; i < bits; i++) inc->key.push_back((uint8_t)((key[i >> 6] >> (i & 63)) & 1));
    return LPR_OK;
  }
  int set_incumbent(const Incumbent& inc) override {
    const int W = (p.n + 63) / 64;
    std::vector<uint64_t> key(W, 0);
    for (size_t i = 0; i < inc.key.size(); i++)
      if (inc.key[i]) key[i >> 6] |= 1ull << (i & 63);
    std::vector<uint8_t> ch(p.n);
    for (int i = 0; i < p.n; i++) ch[i] = inc.payload[i] != 0.0;
    return lpr_knap_set_incumbent(h, inc.z, ch.data(), key.data(), (int)inc.key.size());
  }
  int export_nodes(int max_nodes, void* dbuf, int64_t cap, int64_t* bytes, int* n) override {
    return lpr_knap_export_nodes(h, max_nodes, dbuf, cap, bytes, n);
  }
  int import_nodes(const void* dbuf, int64_t bytes) override { return lpr_knap_import_nodes(h, dbuf, bytes); }
};

// ---- the round loop -------------------------------------------------------------------------------------------------
struct Barrier {
  std::mutex m;
  std::condition_variable cv;
  int n, waiting = 0, generation = 0;
  explicit Barrier(int nn) : n(nn) {}
  void wait() {
    std::unique_lock<std::mutex> lock(m);
    const int gen = generation;
    if (++waiting == n) {
      waiting = 0;
      generation++;
      cv.notify_all();
    } else {
      cv.wait(lock, [&] { return gen != generation; });
    }
  }
};

struct Steal {
  int donor, recv;
  int64_t give;
};
// deterministic plan computed identically by every rank: ranks at or below the low-water mark receive half of the
// pool of the currently fullest rank (same rule as distributed.steal_plan)
std::vector<Steal> steal_plan(std::vector<int64_t> counts, int64_t low_water) {
  std::vector<Steal> plan;
  const int n = (int)counts.size();
  std::vector<int> receivers;
  for (int r = 0; r < n; r++)
    if (counts[r] <= low_water) receivers.push_back(r);
  for (int r : receivers) {
    int donor = 0;
    for (int q = 1; q < n; q++)
      if (counts[q] > counts[donor]) donor = q;
    const int64_t give = counts[donor] / 2;
    if (donor == r || counts[donor] < 2 || give < 1 || counts[donor] <= 2 * std::max<int64_t>(1, low_water)) continue;
    plan.push_back({donor, r, give});
    counts[donor] -= give;
    counts[r] += give;
  }
  return plan;
}

struct Config {
  int64_t max_nodes = -1;      // total node budget over all ranks, < 0 = none
  int64_t max_rounds = -1;     // < 0 = until the pools are empty
  double slice_seconds = 0.0;  // > 0: a round is a time slice
  bool untimed_when_alone = false;  // one rank and no round limit: nobody to exchange with, run to the end in one go
  int64_t chunk_nodes = 4096;  // at most this many nodes per rank and round
  int64_t seed_nodes_per_rank = 8, low_water = 0;
  size_t stage_bytes = 256u << 20;
  int ranks_per_gpu = 1;  // pools (host thread + stream each) per device: their kernels overlap on the GPU
};

struct Shared {
  explicit Shared(int n)
      : barrier(n), slots(n), rc(n, LPR_OK), err(n), nodes(n, 0), run_s(n, 0.0), plan_bytes(64, 0),
        vec_slots((size_t)n * (4 + n), 0.0), vec_result(4 + n, 0.0), stage_out_ptr(n, nullptr), stage_in_ptr(n, nullptr) {}
  Barrier barrier;
  std::vector<double> vec_slots, vec_result;          // per-round status vectors of every rank / their maximum
  std::vector<uint8_t*> stage_out_ptr, stage_in_ptr;  // device staging buffers of every rank (steals)
  std::vector<Incumbent> slots;
  std::vector<int> rc;
  std::vector<std::string> err;
  std::vector<int64_t> nodes;
  std::vector<double> run_s;
  std::vector<int64_t> plan_bytes;
  Incumbent final_inc;
  int64_t rounds = 0, steals = 0, moved = 0, open_left = 0, pivots = 0, depth_overflow = 0;
  double t_seed = 0, t_exchange = 0, t_steal = 0, t_loop = 0;
  std::mutex m;
};

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Ranks are host threads.  With ranks_per_gpu > 1 several pools share a device (each with its own stream, so the
// bandwidth-bound child construction of one overlaps the latency-bound pivot chains of another); the first rank of a
// device is its leader and owns the device's NCCL communicator: status vectors are combined in host memory inside a
// device and all-reduced over NCCL between the leaders, node records move with cudaMemcpyAsync inside a device and
// with ncclSend / ncclRecv (issued by the leaders) between devices.
void rank_main(int rank, int world, int rpg, int device, Pool* pool, ncclComm_t comm, const Config& cfg, Shared* sh) {
  NcclApi& api = nccl_api();
  const int gpu = rank / rpg, n_gpus = world / rpg;
  const bool leader = rank % rpg == 0;
  auto gpu_of = [&](int r) { return r / rpg; };
  int rc = LPR_OK;
  auto note = [&](int code) {  // remember the first failure of this rank (the message is thread local)
    if (code != LPR_OK && rc == LPR_OK) {
      rc = code;
      sh->err[rank] = last_error();
    }
  };
  cudaStream_t stream = nullptr;
  double *d_vec = nullptr, *h_vec = nullptr;
  uint8_t *stage_out = nullptr, *stage_in = nullptr;
  const int nvec = 4 + world;  // z, changed, error, active, counts
  auto cuda_ok = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess) note(fail(LPR_E_CUDA, "%s failed on device %d: %s", what, device, cudaGetErrorString(e)));
  };
  note(select_device(device));
  if (rc == LPR_OK) note(pool->create(device, rank == 0));
  if (rc == LPR_OK) {
    cuda_ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate");
    cuda_ok(cudaMalloc(&d_vec, sizeof(double) * nvec), "cudaMalloc");
    cuda_ok(cudaMallocHost(&h_vec, sizeof(double) * nvec), "cudaMallocHost");
    if (world > 1) {
      cuda_ok(cudaMalloc(&stage_out, cfg.stage_bytes), "cudaMalloc(stage)");
      cuda_ok(cudaMalloc(&stage_in, cfg.stage_bytes), "cudaMalloc(stage)");
      sh->stage_out_ptr[rank] = stage_out;
      sh->stage_in_ptr[rank] = stage_in;
    }
  }
  // MAX over every rank's vector: host memory inside a device, ncclAllReduce between the devices' leaders
  auto allreduce = [&]() {
    if (world == 1 || !h_vec) return;
    double* slot = &sh->vec_slots[(size_t)rank * nvec];
    for (int i = 0; i < nvec; i++) slot[i] = h_vec[i];
    sh->barrier.wait();
    if (leader) {
      for (int r = gpu * rpg + 1; r < (gpu + 1) * rpg; r++)
        for (int i = 0; i < nvec; i++) h_vec[i] = std::max(h_vec[i], sh->vec_slots[(size_t)r * nvec + i]);
      if (n_gpus > 1) {
        cudaError_t e = cudaMemcpyAsync(d_vec, h_vec, sizeof(double) * nvec, cudaMemcpyHostToDevice, stream);
        int nr = 0;
        if (e == cudaSuccess) nr = api.AllReduce(d_vec, d_vec, nvec, kNcclFloat64, kNcclMax, comm, stream);
        if (e == cudaSuccess && nr == 0) e = cudaMemcpyAsync(h_vec, d_vec, sizeof(double) * nvec, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess && nr == 0) e = cudaStreamSynchronize(stream);
        if (nr != 0) note(fail(LPR_E_NCCL, "ncclAllReduce failed: %s", api.GetErrorString(nr)));
        cuda_ok(e, "incumbent all-reduce");
      }
      if (rank == 0)
        for (int i = 0; i < nvec; i++) sh->vec_result[i] = h_vec[i];
    }
    sh->barrier.wait();
    for (int i = 0; i < nvec; i++) h_vec[i] = sh->vec_result[i];
  };
  sh->rc[rank] = rc;
  sh->barrier.wait();  // every pool exists -- or some rank failed to set up, which all ranks learn here, before NCCL
  bool setup_failed = false;
  for (int r = 0; r < world; r++) setup_failed |= sh->rc[r] != LPR_OK;
  if (!setup_failed) {
    // outside the timed region: NCCL builds the channels of a collective and of every point-to-point pair lazily
    // (the first ncclSend/ncclRecv between two devices costs ~100 ms), so run each once
    for (int i = 0; i < nvec; i++) h_vec[i] = 0.0;
    allreduce();
    if (leader && n_gpus > 1) {
      int nr = api.GroupStart();
      for (int g = 0; g < n_gpus && nr == 0; g++) {
        if (g == gpu) continue;
        nr = api.Send(stage_out, 8, kNcclUint8, g, comm, stream);
        if (nr == 0) nr = api.Recv(stage_in + 64 * (size_t)g, 8, kNcclUint8, g, comm, stream);
      }
      const int ne = api.GroupEnd();
      if (nr == 0) nr = ne;
      if (nr != 0) note(fail(LPR_E_NCCL, "point-to-point warm-up failed: %s", api.GetErrorString(nr)));
      cuda_ok(cudaStreamSynchronize(stream), "point-to-point warm-up");
    }
  }
  sh->barrier.wait();
  const double t_begin = now_s();
  int64_t processed = 0, steals = 0, moved = 0, rounds = 0;
  double t_exchange = 0, t_steal = 0, t_run = 0, t_seed = 0;
  // seeding: expand the root a little so that the first round has work to share
  if (rc == LPR_OK && world > 1 && (rank == 0 || pool->replicated_root())) {
    int guard = 0;
    while (pool->open_count() > 0 && pool->open_count() < cfg.seed_nodes_per_rank * world && guard++ < 64) {
      int64_t done = 0;
      note(pool->run(std::max<int64_t>(1, cfg.seed_nodes_per_rank), 0.0, &done));
      if (rc != LPR_OK) break;
      if (rank == 0) processed += done;  // the replicas' copies of the seed nodes are not counted
    }
    if (rc == LPR_OK && pool->replicated_root()) note(pool->keep_stride(rank, world));
  }
  t_seed = now_s() - t_begin;
  Incumbent agreed;  // the incumbent every rank agreed on last
  Incumbent mine;
  while (!setup_failed) {
    rounds++;
    const double ta = now_s();
    if (rc == LPR_OK) note(pool->get_incumbent(&mine));
    const bool same = mine.has ? (agreed.has && mine.z == agreed.z && mine.key == agreed.key) : !agreed.has;
    std::vector<int64_t> counts(world, 0);
    double zmax = mine.has ? mine.z : -INFINITY;
    bool changed = !same, any_error = rc != LPR_OK;
    const int64_t my_open = rc == LPR_OK ? pool->open_count() : 0;
    const int64_t share = cfg.max_nodes >= 0 ? (cfg.max_nodes + world - 1) / world : -1;  // node budget of a rank
    bool active = my_open > 0 && (share < 0 || processed < share);
    if (world > 1) {
      for (int i = 0; i < nvec; i++) h_vec[i] = 0.0;
      h_vec[0] = zmax;
      h_vec[1] = changed ? 1.0 : 0.0;
      h_vec[2] = rc != LPR_OK ? 1.0 : 0.0;
      h_vec[3] = (share < 0 || processed < share) ? 1.0 : 0.0;  // some rank can still take work
      h_vec[4 + rank] = (double)my_open;
      allreduce();
      zmax = h_vec[0];
      changed = h_vec[1] != 0.0;
      any_error = h_vec[2] != 0.0 || rc != LPR_OK;
      active = h_vec[3] != 0.0;
      for (int r = 0; r < world; r++) counts[r] = (int64_t)h_vec[4 + r];
    } else {
      counts[0] = my_open;
    }
    if (world > 1 && changed) {  // DFS-first key among the holders of the best value; everybody adopts it
      sh->slots[rank] = (mine.has && mine.z == zmax) ? mine : Incumbent();
      sh->barrier.wait();
      Incumbent best;
      for (int r = 0; r < world; r++)
        if (better(sh->slots[r], best)) best = sh->slots[r];
      if (rc == LPR_OK && better(best, mine)) note(pool->set_incumbent(best));
      agreed = best;
      sh->barrier.wait();
    } else if (changed) {
      agreed = mine;
    }
    const double tb = now_s();
    t_exchange += tb - ta;
    int64_t total_open = 0;
    for (int64_t c : counts) total_open += c;
    if (any_error || total_open == 0 || !active) break;
    if (cfg.max_rounds >= 0 && rounds > cfg.max_rounds) break;
    // ---- work stealing: node records device to device over NVLink
    if (world > 1) {
      std::vector<Steal> plan = steal_plan(counts, cfg.low_water);
      if (!plan.empty()) {
        if (plan.size() > sh->plan_bytes.size()) plan.resize(sh->plan_bytes.size());
        int64_t used = 0;
        for (size_t k = 0; k < plan.size(); k++) {
          if (plan[k].donor != rank) continue;
          int64_t bytes = 0;
          int n = 0;
          if (rc == LPR_OK)
            note(pool->export_nodes((int)std::min<int64_t>(plan[k].give, 1 << 30), stage_out + used,
                                    (int64_t)cfg.stage_bytes - used, &bytes, &n));
          if (rc != LPR_OK) bytes = 0;
          sh->plan_bytes[k] = bytes;
          used += bytes;
          if (bytes > 0) {
            steals++;
            moved += n;
          }
        }
        sh->barrier.wait();  // byte counts of every transfer are known to everybody
        auto out_off = [&](size_t k) {  // where transfer k starts in its donor's staging buffer
          int64_t o = 0;
          for (size_t q = 0; q < k; q++)
            if (plan[q].donor == plan[k].donor) o += sh->plan_bytes[q];
          return o;
        };
        // inside a device the receiver copies; between devices the leaders of both ends drive NCCL
        for (size_t k = 0; k < plan.size(); k++) {
          const int64_t bytes = sh->plan_bytes[k];
          if (bytes <= 0 || plan[k].recv != rank || gpu_of(plan[k].donor) != gpu) continue;
          cuda_ok(cudaMemcpyAsync(stage_in, sh->stage_out_ptr[plan[k].donor] + out_off(k), (size_t)bytes,
                                  cudaMemcpyDeviceToDevice, stream), "node transfer inside a device");
        }
        if (leader && n_gpus > 1) {
          int nr = api.GroupStart();
          for (size_t k = 0; k < plan.size() && nr == 0; k++) {
            const int64_t bytes = sh->plan_bytes[k];
            const int gd = gpu_of(plan[k].donor), gr = gpu_of(plan[k].recv);
            if (bytes <= 0 || gd == gr) continue;
            if (gd == gpu)
              nr = api.Send(sh->stage_out_ptr[plan[k].donor] + out_off(k), (size_t)bytes, kNcclUint8, gr, comm, stream);
            if (gr == gpu && nr == 0)  // a receiver appears once in a plan: its staging buffer is filled from the start
              nr = api.Recv(sh->stage_in_ptr[plan[k].recv], (size_t)bytes, kNcclUint8, gd, comm, stream);
          }
          const int ne = api.GroupEnd();
          if (nr == 0) nr = ne;
          if (nr != 0) note(fail(LPR_E_NCCL, "node transfer failed: %s", api.GetErrorString(nr)));
        }
        cuda_ok(cudaStreamSynchronize(stream), "node transfer");
        sh->barrier.wait();  // every record has landed
        for (size_t k = 0; k < plan.size(); k++)
          if (plan[k].recv == rank && sh->plan_bytes[k] > 0 && rc == LPR_OK)
            note(pool->import_nodes(stage_in, sh->plan_bytes[k]));
      }
    }
    const double tc = now_s();
    t_steal += tc - tb;
    // ---- one slice of node work
    if (rc == LPR_OK && pool->open_count() > 0) {
      int64_t budget = cfg.chunk_nodes;
      if (share >= 0) budget = std::min<int64_t>(budget, std::max<int64_t>(0, share - processed));
      int64_t done = 0;
      const bool alone = world == 1 && cfg.untimed_when_alone && cfg.max_rounds < 0;
      if (budget > 0) note(pool->run(budget, alone ? 0.0 : cfg.slice_seconds, &done));
      processed += done;
    }
    t_run += now_s() - tc;
  }
  const double t_loop = now_s() - t_begin;
  // ---- results
  if (rc == LPR_OK) note(pool->get_incumbent(&mine));
  {
    std::lock_guard<std::mutex> lock(sh->m);
    if (better(mine, sh->final_inc)) sh->final_inc = mine;
    sh->steals += steals;
    sh->moved += moved;
    sh->nodes[rank] = processed;
    sh->run_s[rank] = t_run;
    sh->rounds = std::max(sh->rounds, rounds);
    sh->t_loop = std::max(sh->t_loop, t_loop);
    if (pool) {
      sh->open_left += rc == LPR_OK ? pool->open_count() : 0;
      sh->pivots += pool->pivots();
      sh->depth_overflow += rc == LPR_OK ? pool->depth_overflow() : 0;
    }
    if (rank == 0) {
      sh->t_seed = t_seed;
      sh->t_exchange = t_exchange;
      sh->t_steal = t_steal;
    }
  }
  sh->rc[rank] = rc;
  if (rc != LPR_OK && sh->err[rank].empty()) sh->err[rank] = last_error();
  pool->destroy();
  if (stage_out) cudaFree(stage_out);
  if (stage_in) cudaFree(stage_in);
  if (d_vec) cudaFree(d_vec);
  if (h_vec) cudaFreeHost(h_vec);
  if (stream) cudaStreamDestroy(stream);
}

template <class MakePool>
int solve_mgpu(int n_gpus, const int* devices, const Config& cfg, MakePool make_pool, Incumbent* inc_out, Shared** sh_out,
               lpr_mgpu_stats* stats) {
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) return fail(LPR_E_CUDA, "no CUDA device");
  if (n_gpus < 1 || n_gpus > 64) return fail(LPR_E_BADARG, "n_gpus=%d out of range", n_gpus);
  if (n_gpus > have && !devices) return fail(LPR_E_BADARG, "n_gpus=%d but the box has %d devices", n_gpus, have);
  std::vector<int> devs(n_gpus);
  for (int r = 0; r < n_gpus; r++) {
    devs[r] = devices ? devices[r] : r;
    if (devs[r] < 0 || devs[r] >= have) return fail(LPR_E_BADARG, "device %d does not exist", devs[r]);
  }
  CommSet* cs = nullptr;
  if (n_gpus > 1) {
    int rc = get_comms(devs, &cs);
    if (rc) return rc;
  }
  const int rpg = std::max(1, std::min(cfg.ranks_per_gpu, 8));
  const int world = n_gpus * rpg;
  Shared* sh = new Shared(world);
  std::vector<Pool*> pools(world);
  for (int r = 0; r < world; r++) pools[r] = make_pool();
  const double t0 = now_s();
  std::vector<std::thread> threads;
  auto comm_of = [&](int r) { return (cs && r % rpg == 0) ? cs->comms[r / rpg] : (ncclComm_t) nullptr; };
  for (int r = 1; r < world; r++)
    threads.emplace_back(rank_main, r, world, rpg, devs[r / rpg], pools[r], comm_of(r), std::cref(cfg), sh);
  rank_main(0, world, rpg, devs[0], pools[0], comm_of(0), cfg, sh);
  for (auto& t : threads) t.join();
  const double dt = now_s() - t0;
  for (Pool* p : pools) delete p;
  cudaSetDevice(devs[0]);
  for (int r = 0; r < world; r++)
    if (sh->rc[r] != LPR_OK) {
      const int rc = sh->rc[r];
      const std::string msg = sh->err[r];
      delete sh;
      return fail(rc, "rank %d (device %d): %s", r, devs[r / rpg], msg.c_str());
    }
  if (stats) {
    memset(stats, 0, sizeof *stats);
    stats->n_gpus = n_gpus;
    stats->ranks_per_gpu = rpg;
    stats->nccl_version = n_gpus > 1 ? nccl_api().version : 0;
    stats->rounds = sh->rounds;
    stats->steals = sh->steals;
    stats->nodes_moved = sh->moved;
    stats->open_left = sh->open_left;
    stats->depth_overflow = sh->depth_overflow;
    stats->seconds = sh->t_loop;          // seeding + rounds, max over the ranks (pool creation and NCCL set-up excluded)
    stats->setup_seconds = dt - sh->t_loop;
    stats->seed_seconds = sh->t_seed;
    stats->exchange_seconds = sh->t_exchange;
    stats->steal_seconds = sh->t_steal;
    for (int r = 0; r < world; r++) {
      const int g = r / rpg;
      if (g >= 16) break;
      stats->nodes_per_gpu[g] += sh->nodes[r];
      stats->run_seconds_per_gpu[g] = std::max(stats->run_seconds_per_gpu[g], sh->run_s[r]);
    }
  }
  *inc_out = sh->final_inc;
  *sh_out = sh;
  return LPR_OK;
}

}  // namespace

extern "C" {

int lpr_bb_solve_mgpu(int n_gpus, const int* devices, int rows, int cols, const double* final_tableau, int n_vars,
                      int enable_pruning, int64_t max_nodes, int64_t max_rounds, double slice_seconds, double* x,
                      double* z, int* has_solution, int64_t* nodes, int64_t* pivots, int* status,
                      lpr_mgpu_stats* stats) {
  if (!final_tableau) return fail(LPR_E_BADARG, "null tableau");
  BBProblem prob{rows, cols, n_vars, enable_pruning, final_tableau};
  Config cfg;
  cfg.max_nodes = max_nodes;
  cfg.max_rounds = max_rounds;
  cfg.slice_seconds = slice_seconds;
  cfg.chunk_nodes = slice_seconds > 0.0 ? (1 << 20) : 1024;
  cfg.seed_nodes_per_rank = 8;
  cfg.low_water = 0;
  const char* sm = getenv("LPR_MG_STAGE_MB");
  cfg.stage_bytes = (size_t)(sm ? std::max(16, atoi(sm)) : 512) << 20;
  // two pools per device by default: the child construction of one (HBM bound) runs beside the pivot chains of the
  // other (latency bound); each gets its share of the slab budget
  const char* rp = getenv("LPR_MG_RANKS_PER_GPU");
  // ... when the host has the cores for it: every pool is a host thread that spins on its stream (measured on a
  // 16-core box: 8 GPUs x 2 pools ran at 4.4x one GPU, the threads were starving each other)
  const unsigned cores = std::max(1u, std::thread::hardware_concurrency());
  cfg.ranks_per_gpu = rp ? std::max(1, atoi(rp)) : ((unsigned)n_gpus * 4u <= cores ? 2 : 1);
  bb_set_prealloc_share(cfg.ranks_per_gpu);
  Incumbent inc;
  Shared* sh = nullptr;
  int rc = solve_mgpu(n_gpus, devices, cfg, [&]() -> Pool* { return new BBPoolC(prob); }, &inc, &sh, stats);
  bb_set_prealloc_share(1);
  if (rc) return rc;
  int64_t total = 0;
  for (int64_t c : sh->nodes) total += c;
  if (has_solution) *has_solution = inc.has ? 1 : 0;
  if (z) *z = inc.has ? inc.z : -INFINITY;
  if (x)
    for (int i = 0; i < n_vars; i++) x[i] = inc.has ? inc.payload[i] : 0.0;
  if (nodes) *nodes = total;
  if (pivots) *pivots = sh->pivots;
  if (status) *status = sh->depth_overflow > 0 ? LPR_DEPTH_LIMIT : (sh->open_left > 0 ? LPR_NODE_LIMIT : LPR_OPTIMAL);
  delete sh;
  return LPR_OK;
}

int lpr_knap_solve_mgpu(int n_gpus, const int* devices, double capacity, int n, const double* weights,
                        const double* values, int64_t max_nodes, int64_t max_rounds, double slice_seconds, double* best,
                        uint8_t* chosen, int64_t* nodes, int* status, lpr_mgpu_stats* stats) {
  if (n < 1 || !weights || !values) return fail(LPR_E_BADARG, "bad knapsack instance (n=%d)", n);
  KnapProblem prob{capacity, n, weights, values};
  Config cfg;
  cfg.max_nodes = max_nodes;
  cfg.max_rounds = max_rounds;
  cfg.slice_seconds = slice_seconds > 0.0 ? slice_seconds : 2e-3;
  cfg.untimed_when_alone = true;
  cfg.chunk_nodes = 1LL << 40;
  cfg.seed_nodes_per_rank = 256;
  cfg.low_water = 4096;  // a rank that cannot fill a fraction of a batch is about to run dry: refill it early
  const char* sm = getenv("LPR_MG_STAGE_MB");
  cfg.stage_bytes = (size_t)(sm ? std::max(16, atoi(sm)) : 256) << 20;
  const char* rp = getenv("LPR_MG_KNAP_RANKS_PER_GPU");
  cfg.ranks_per_gpu = rp ? std::max(1, atoi(rp)) : 1;
  Incumbent inc;
  Shared* sh = nullptr;
  int rc = solve_mgpu(n_gpus, devices, cfg, [&]() -> Pool* { return new KnapPoolC(prob); }, &inc, &sh, stats);
  if (rc) return rc;
  int64_t total = 0;
  for (int64_t c : sh->nodes) total += c;
  if (best) *best = inc.has ? inc.z : 0.0;
  if (chosen)
    for (int i = 0; i < n; i++) chosen[i] = inc.has ? (uint8_t)(inc.payload[i] != 0.0) : 0;
  if (nodes) *nodes = total;
  if (status) *status = sh->open_left > 0 ? LPR_NODE_LIMIT : LPR_OPTIMAL;
  delete sh;
  return LPR_OK;
}

int lpr_nccl_version(int* version) {
  if (!version) return fail(LPR_E_BADARG, "null argument");
  NcclApi& api = nccl_api();
  if (!api.error.empty()) return fail(LPR_E_NCCL, "%s", api.error.c_str());
  *version = api.version;
  return LPR_OK;
}

}  // extern "C"
