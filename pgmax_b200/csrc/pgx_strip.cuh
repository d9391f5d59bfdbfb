// pgx_strip.cuh — row strips of ONE 2-D Ising-type lattice across the GPUs of a box
// (BASELINE.json configs[4]).  Included at the end of pgx.cu.
//
// The reference has no multi-GPU path; the contract is the north-star's: "a single giant grid
// graph is split into row strips, and each iteration exchanges the boundary variable-to-factor
// messages with NCCL over NVLink".  One process per GPU; rank g owns `rows` consecutive rows of
// the n_rows x n_cols torus (variable (i, j) owns the vertical factor to (i + 1, j) and the
// horizontal factor to (i, j + 1), examples/ising_model.ipynb cell 12) and their factors.  Only
// the vertical factors of a strip's last row straddle two ranks, and both halo quantities depend
// only on the previous iteration's messages, so ONE simultaneous exchange per iteration suffices:
//
//   g -> g+1 : the boundary factors' messages into the first row of g+1 (2 n floats), which g+1
//              adds to that row's sums where the single graph's message order puts them
//              (LatticeBinArgs::up_add / up_last);
//   g+1 -> g : evidence and the three messages from g+1's own factors of its first row, UNSUMMED
//              (5 of 8 n floats): g adds them to its ghost row after its own vertical message, in
//              the single graph's order (LatticeBinArgs::ghost_terms).
//
// Per iteration, on the strip's own streams (all of it captured ONCE per (buffers, iteration
// count, damping, temperature) in a CUDA graph and replayed with one cudaGraphLaunch per run):
//
//   main : k_strip_pack                       -> event
//   comm : ncclGroup{send down, send up, recv up_add, recv ghost}   (NCCL p2p over NVLink)
//   main : k_lattice_bin on the interior rows [1, rows - 1)         (overlaps the exchange)
//   main : wait for the exchange; k_lattice_bin on rows 0 and rows - 1
//
// Messages stay in binary-difference storage between iterations (lattice_bin.cuh); the ABI arrays
// are the reference's flat layout.  Every variable sum - the boundary rows' included - is formed in
// the single graph's order, so N ranks are BIT-IDENTICAL to one rank (and, for max-product, to the
// oracle) at any horizon; asserted by the tests.  NCCL is resolved at run time from the process
// (torch's bundled libnccl.so.2, or the system one): the library does not link against it.

#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the process already uses (torch's)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
#define PGX_NCCL_SYM(field, name)                                          \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));       \
  if (!api.field) return nullptr
  PGX_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
  PGX_NCCL_SYM(CommInitRank, "ncclCommInitRank");
  PGX_NCCL_SYM(CommDestroy, "ncclCommDestroy");
  PGX_NCCL_SYM(Send, "ncclSend");
  PGX_NCCL_SYM(Recv, "ncclRecv");
  PGX_NCCL_SYM(GroupStart, "ncclGroupStart");
  PGX_NCCL_SYM(GroupEnd, "ncclGroupEnd");
  PGX_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef PGX_NCCL_SYM
  api.handle = h;
  return &api;
}

#define PGX_NCCL(api, expr)                                                                        \
  do {                                                                                             \
    ncclResult_t r__ = (expr);                                                                     \
    if (r__ != ncclSuccess)                                                                        \
      return fail(PGX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, (api)->GetErrorString(r__), __FILE__, __LINE__); \
  } while (0)

struct StripGraph {
  const void *lp, *ev, *in, *out;
  int32_t iters;
  float damping, temperature;
  uint32_t flags;
  cudaGraphExec_t exec;
};

}  // namespace

struct pgx_strip {
  int device = 0, num_sms = 0, rank = 0, world = 1;
  int64_t N = 0, R = 0;
  float *cA = nullptr, *cB = nullptr;                                    // compressed messages, float4 per cell
  float *send_down = nullptr, *send_up = nullptr, *up_add = nullptr, *ghost = nullptr;  // [2 N] each
  ncclComm_t comm = nullptr;
  cudaStream_t s_main = nullptr, s_comm = nullptr;
  cudaEvent_t e_pack = nullptr, e_comm = nullptr;
  uint32_t attr_done = 0;
  int64_t launches = 0, graph_launches = 0;
  std::vector<StripGraph> graphs;
};

namespace {

struct StripGuard {
  int prev = -1;
  bool switched = false;
  int enter(const pgx_strip* s) {
    if (cudaGetDevice(&prev) != cudaSuccess) return fail(PGX_ERR_NO_DEVICE, "no CUDA device");
    if (prev != s->device) {
      PGX_CUDA(cudaSetDevice(s->device));
      switched = true;
    }
    return PGX_OK;
  }
  ~StripGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

int strip_launch_ok(pgx_strip* s, const char* what) {
  ++s->launches;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(PGX_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(err));
  return PGX_OK;
}

pgx::LatticeBinArgs strip_args(const pgx_strip* s) {
  pgx::LatticeBinArgs g{};
  g.R = int32_t(s->R);
  g.N = int32_t(s->N);
  g.torus = s->world == 1 ? 1 : 0;
  g.up_add = s->world == 1 ? nullptr : s->up_add;
  g.ghost_ev = nullptr;
  g.ghost_terms = s->world == 1 ? nullptr : s->ghost;
  g.up_last = s->rank == 0 ? 1 : 0;  // the strip that holds the torus' row 0
  g.ghost_up_last = s->rank == s->world - 1 ? 1 : 0;  // ... and the strip whose ghost row it is
  return g;
}

// pack (from compressed messages `c`) + the halo exchange; `st` carries the pack, the exchange
// runs on s_comm between the events e_pack and e_comm (the caller waits on e_comm).
int strip_exchange(pgx_strip* s, cudaStream_t st, const float* ev_own, const float* c) {
  NcclApi* api = nccl_api();
  int rc;
  pgx::k_strip_pack<<<unsigned((s->N + pgx::kThreads - 1) / pgx::kThreads), pgx::kThreads, 0, st>>>(
      int32_t(s->R), int32_t(s->N), ev_own, reinterpret_cast<const float4*>(c), reinterpret_cast<float2*>(s->send_down),
      reinterpret_cast<float4*>(s->send_up));
  if ((rc = strip_launch_ok(s, "k_strip_pack"))) return rc;
  PGX_CUDA(cudaEventRecord(s->e_pack, st));
  PGX_CUDA(cudaStreamWaitEvent(s->s_comm, s->e_pack, 0));
  const int down = (s->rank + 1) % s->world, up = (s->rank + s->world - 1) % s->world;
  const size_t n = size_t(2 * s->N), n_up = size_t(8 * s->N);
  PGX_NCCL(api, api->GroupStart());
  PGX_NCCL(api, api->Send(s->send_down, n, ncclFloat, down, s->comm, s->s_comm));
  PGX_NCCL(api, api->Send(s->send_up, n_up, ncclFloat, up, s->comm, s->s_comm));
  PGX_NCCL(api, api->Recv(s->up_add, n, ncclFloat, up, s->comm, s->s_comm));
  PGX_NCCL(api, api->Recv(s->ghost, n_up, ncclFloat, down, s->comm, s->s_comm));
  PGX_NCCL(api, api->GroupEnd());
  ++s->launches;
  PGX_CUDA(cudaEventRecord(s->e_comm, s->s_comm));
  return PGX_OK;
}

// The whole run on stream `st`: compress, num_iters iterations, expand.
int strip_enqueue(pgx_strip* s, cudaStream_t st, const float* lp, const float* ev_own, const float* msgs_in,
                  float* msgs_out, int32_t num_iters, float damping, float temperature, uint32_t flags) {
  int rc;
  const int64_t cells = s->R * s->N;
  const unsigned ew_grid = unsigned(s->num_sms * 8);
  if (msgs_in == nullptr) {
    PGX_CUDA(cudaMemsetAsync(s->cA, 0, size_t(cells) * sizeof(float4), st));
  } else {
    pgx::k_lattice_compress<<<ew_grid, pgx::kThreads, 0, st>>>(reinterpret_cast<const float4*>(msgs_in),
                                                               reinterpret_cast<float4*>(s->cA), cells);
    if ((rc = strip_launch_ok(s, "k_lattice_compress"))) return rc;
  }
  pgx::RunArgs a = make_run_args(damping, temperature, nullptr, num_iters, 8 * cells, 2 * cells);
  pgx::LatticeBinArgs g = strip_args(s);
  const bool sum = temperature > 0.f;
  const bool overlap = !(flags & PGX_STRIP_NO_OVERLAP);
  const float* cur = s->cA;
  float* nxt = s->cB;
  for (int it = 0; it < num_iters; ++it) {
    if (s->world == 1) {
      g.seg_begin[0] = 0; g.seg_end[0] = int32_t(s->R); g.seg_begin[1] = g.seg_end[1] = 0;
      if ((rc = launch_lattice_bin(nullptr, &s->attr_done, s->num_sms, st, g, ev_own, lp, cur, nxt, a, sum, false))) return rc;
      ++s->launches;
    } else {
      if ((rc = strip_exchange(s, st, ev_own, cur))) return rc;
      if (overlap && s->R > 2) {
        // interior rows need no halo: they run while the exchange is in flight
        g.seg_begin[0] = 1; g.seg_end[0] = int32_t(s->R - 1); g.seg_begin[1] = g.seg_end[1] = 0;
        if ((rc = launch_lattice_bin(nullptr, &s->attr_done, s->num_sms, st, g, ev_own, lp, cur, nxt, a, sum, false))) return rc;
        ++s->launches;
        PGX_CUDA(cudaStreamWaitEvent(st, s->e_comm, 0));
        g.seg_begin[0] = 0; g.seg_end[0] = 1; g.seg_begin[1] = int32_t(s->R - 1); g.seg_end[1] = int32_t(s->R);
      } else {
        PGX_CUDA(cudaStreamWaitEvent(st, s->e_comm, 0));
        g.seg_begin[0] = 0; g.seg_end[0] = int32_t(s->R); g.seg_begin[1] = g.seg_end[1] = 0;
      }
      if ((rc = launch_lattice_bin(nullptr, &s->attr_done, s->num_sms, st, g, ev_own, lp, cur, nxt, a, sum, false))) return rc;
      ++s->launches;
    }
    const float* t = cur;
    cur = nxt;
    nxt = const_cast<float*>(t);
  }
  pgx::k_lattice_expand<<<ew_grid, pgx::kThreads, 0, st>>>(reinterpret_cast<const float4*>(cur),
                                                           reinterpret_cast<float4*>(msgs_out), cells);
  return strip_launch_ok(s, "k_lattice_expand");
}

}  // namespace

extern "C" {

int pgx_nccl_unique_id(void* id_out) {
  if (!id_out) return fail(PGX_ERR_INVALID, "null argument");
  NcclApi* api = nccl_api();
  if (!api) return fail(PGX_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded: %s", dlerror());
  static_assert(sizeof(ncclUniqueId) == PGX_NCCL_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  PGX_NCCL(api, api->GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return PGX_OK;
}

void pgx_strip_destroy(pgx_strip* s) {
  if (!s) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(s->device);
  for (StripGraph& g : s->graphs) cudaGraphExecDestroy(g.exec);
  if (s->comm && nccl_api()) nccl_api()->CommDestroy(s->comm);
  free_dev(s->cA); free_dev(s->cB); free_dev(s->send_down); free_dev(s->send_up); free_dev(s->up_add); free_dev(s->ghost);
  if (s->e_pack) cudaEventDestroy(s->e_pack);
  if (s->e_comm) cudaEventDestroy(s->e_comm);
  if (s->s_main) cudaStreamDestroy(s->s_main);
  if (s->s_comm) cudaStreamDestroy(s->s_comm);
  if (prev >= 0) cudaSetDevice(prev);
  delete s;
}

int pgx_strip_create(int64_t n_cols, int64_t rows, int rank, int world, const void* nccl_id, pgx_strip** out) {
  if (!out) return fail(PGX_ERR_INVALID, "null argument");
  *out = nullptr;
  PGX_CHECK(world >= 1 && rank >= 0 && rank < world, "rank %d outside a world of %d", rank, world);
  PGX_CHECK(n_cols >= 2 && n_cols < (int64_t(1) << 30) && rows >= 2 && rows < (int64_t(1) << 30),
            "a strip needs at least 2 rows and 2 columns (got %lld x %lld)", (long long)rows, (long long)n_cols);
  PGX_CHECK(world == 1 || nccl_id != nullptr, "nccl_id is null");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(PGX_ERR_NO_DEVICE, "no CUDA device visible; pgx has no CPU fallback");
  pgx_strip* s = new pgx_strip();
  auto bail = [&](int rc) {
    pgx_strip_destroy(s);
    return rc;
  };
#define PGX_STRIP_TRY(expr)                                                                          \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess) return bail(fail(PGX_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__))); \
  } while (0)
  PGX_STRIP_TRY(cudaGetDevice(&s->device));
  cudaDeviceProp prop;
  PGX_STRIP_TRY(cudaGetDeviceProperties(&prop, s->device));
  s->num_sms = prop.multiProcessorCount;
  s->rank = rank;
  s->world = world;
  s->N = n_cols;
  s->R = rows;
  const size_t cbytes = size_t(rows) * size_t(n_cols) * sizeof(float4);
  PGX_STRIP_TRY(cudaMalloc(reinterpret_cast<void**>(&s->cA), cbytes));
  PGX_STRIP_TRY(cudaMalloc(reinterpret_cast<void**>(&s->cB), cbytes));
  const size_t hbytes = size_t(8 * n_cols) * sizeof(float);  // (the two "up" buffers use all 8 floats per column)
  for (float** p : {&s->send_down, &s->send_up, &s->up_add, &s->ghost}) {
    PGX_STRIP_TRY(cudaMalloc(reinterpret_cast<void**>(p), hbytes));
    PGX_STRIP_TRY(cudaMemset(*p, 0, hbytes));
  }
  PGX_STRIP_TRY(cudaStreamCreateWithFlags(&s->s_main, cudaStreamNonBlocking));
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  PGX_STRIP_TRY(cudaStreamCreateWithPriority(&s->s_comm, cudaStreamNonBlocking, prio_hi));
  PGX_STRIP_TRY(cudaEventCreateWithFlags(&s->e_pack, cudaEventDisableTiming));
  PGX_STRIP_TRY(cudaEventCreateWithFlags(&s->e_comm, cudaEventDisableTiming));
  if (world > 1) {
    NcclApi* api = nccl_api();
    if (!api) return bail(fail(PGX_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded"));
    ncclUniqueId id;
    std::memcpy(&id, nccl_id, sizeof(id));
    ncclResult_t r = api->CommInitRank(&s->comm, world, id, rank);
    if (r != ncclSuccess) return bail(fail(PGX_ERR_CUDA, "ncclCommInitRank failed: %s", api->GetErrorString(r)));
    // one eager exchange: NCCL sets up its peer connections on first use, which must not
    // happen inside a stream capture
    PGX_STRIP_TRY(cudaMemset(s->cA, 0, cbytes));
    PGX_STRIP_TRY(cudaDeviceSynchronize());
    int rc = strip_exchange(s, s->s_main, s->up_add /* any [2 N] zeros as evidence */, s->cA);
    if (rc) return bail(rc);
    PGX_STRIP_TRY(cudaStreamWaitEvent(s->s_main, s->e_comm, 0));
    PGX_STRIP_TRY(cudaStreamSynchronize(s->s_main));
    PGX_STRIP_TRY(cudaMemset(s->up_add, 0, hbytes));
    PGX_STRIP_TRY(cudaMemset(s->ghost, 0, hbytes));
    PGX_STRIP_TRY(cudaDeviceSynchronize());
  }
#undef PGX_STRIP_TRY
  *out = s;
  return PGX_OK;
}

int64_t pgx_strip_launch_count(const pgx_strip* s) { return s ? s->launches : 0; }
int64_t pgx_strip_graph_launch_count(const pgx_strip* s) { return s ? s->graph_launches : 0; }

int pgx_strip_run(pgx_strip* s, void* stream, const float* log_potentials, const float* evidence_own,
                  const float* msgs_in, float* msgs_out, int32_t num_iters, float damping, float temperature,
                  uint32_t flags) {
  if (!s) return fail(PGX_ERR_INVALID, "null strip");
  PGX_CHECK(num_iters >= 1, "num_iters must be >= 1, got %d", num_iters);
  PGX_CHECK(temperature >= 0.f, "temperature must be >= 0");
  PGX_CHECK(log_potentials && evidence_own && msgs_out, "null buffer");
  const auto misaligned = [](const void* p, uintptr_t mask) { return (reinterpret_cast<uintptr_t>(p) & mask) != 0; };
  PGX_CHECK(!misaligned(log_potentials, 15) && !misaligned(msgs_in, 15) && !misaligned(msgs_out, 15) &&
                !misaligned(evidence_own, 7),
            "strip buffers must be 16-byte aligned (evidence: 8)");
  int rc;
  StripGuard guard;
  if ((rc = guard.enter(s))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (flags & PGX_STRIP_NO_GRAPH)
    return strip_enqueue(s, st, log_potentials, evidence_own, msgs_in, msgs_out, num_iters, damping, temperature, flags);
  // one CUDA graph per (buffers, iteration count, scalars): captured on the strip's own stream,
  // replayed on the caller's
  StripGraph* hit = nullptr;
  for (StripGraph& g : s->graphs)
    if (g.lp == log_potentials && g.ev == evidence_own && g.in == msgs_in && g.out == msgs_out && g.iters == num_iters &&
        g.damping == damping && g.temperature == temperature && g.flags == flags)
      hit = &g;
  if (!hit) {
    const int64_t launches0 = s->launches;
    PGX_CUDA(cudaStreamBeginCapture(s->s_main, cudaStreamCaptureModeRelaxed));
    rc = strip_enqueue(s, s->s_main, log_potentials, evidence_own, msgs_in, msgs_out, num_iters, damping, temperature, flags);
    cudaGraph_t graph = nullptr;
    cudaError_t err = cudaStreamEndCapture(s->s_main, &graph);
    s->launches = launches0;  // nothing ran yet
    if (rc != PGX_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (err != cudaSuccess) return fail(PGX_ERR_CUDA, "capturing the strip run failed: %s", cudaGetErrorString(err));
    cudaGraphExec_t exec = nullptr;
    err = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (err != cudaSuccess) return fail(PGX_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(err));
    if (s->graphs.size() >= 8) {  // small cache, oldest out
      cudaGraphExecDestroy(s->graphs.front().exec);
      s->graphs.erase(s->graphs.begin());
    }
    s->graphs.push_back(StripGraph{log_potentials, evidence_own, msgs_in, msgs_out, num_iters, damping, temperature, flags, exec});
    hit = &s->graphs.back();
  }
  PGX_CUDA(cudaGraphLaunch(hit->exec, st));
  ++s->graph_launches;
  // kernels + exchanges one replay of the graph runs: compress (if any), expand, and per iteration
  // one lattice launch (one rank) or pack + exchange + interior launch + boundary launch
  const int per_iter = s->world == 1 ? 1 : (((flags & PGX_STRIP_NO_OVERLAP) || s->R <= 2) ? 3 : 4);
  s->launches += (msgs_in ? 1 : 0) + 1 + int64_t(num_iters) * per_iter;
  return PGX_OK;
}

int pgx_strip_beliefs(pgx_strip* s, void* stream, const float* evidence_own, const float* msgs, float* beliefs_out) {
  if (!s) return fail(PGX_ERR_INVALID, "null strip");
  PGX_CHECK(evidence_own && msgs && beliefs_out, "null buffer");
  int rc;
  StripGuard guard;
  if ((rc = guard.enter(s))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t cells = s->R * s->N;
  const unsigned ew_grid = unsigned(s->num_sms * 8);
  pgx::k_lattice_compress<<<ew_grid, pgx::kThreads, 0, st>>>(reinterpret_cast<const float4*>(msgs),
                                                             reinterpret_cast<float4*>(s->cA), cells);
  if ((rc = strip_launch_ok(s, "k_lattice_compress"))) return rc;
  if (s->world > 1) {
    if ((rc = strip_exchange(s, st, evidence_own, s->cA))) return rc;
    PGX_CUDA(cudaStreamWaitEvent(st, s->e_comm, 0));
  }
  pgx::k_lattice_bin_beliefs<<<ew_grid, pgx::kThreads, 0, st>>>(strip_args(s), evidence_own,
                                                                reinterpret_cast<const float4*>(s->cA),
                                                                reinterpret_cast<float2*>(beliefs_out));
  return strip_launch_ok(s, "k_lattice_bin_beliefs");
}

}  // extern "C"
