// pgx.cu — plan builder, workspace and launch sequencing behind include/pgx.h.
//
// Nothing in this file computes messages on the host: the plan builder only
// derives index structures (the variable -> incident-edges CSR, transposed
// configuration lists, parent offsets), uploads them once, and pgx_bp_run
// enqueues kernels.  There is no CPU fallback: without a CUDA device every
// entry point fails with PGX_ERR_NO_DEVICE.

#include "../../include/pgx.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

#include "pgx_kernels.cuh"

#define PGX_STR2(x) #x
#define PGX_STR(x) PGX_STR2(x)

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define PGX_CUDA(expr)                                                                    \
  do {                                                                                    \
    cudaError_t err__ = (expr);                                                           \
    if (err__ != cudaSuccess)                                                             \
      return fail(PGX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                  __FILE__, __LINE__);                                                    \
  } while (0)

#define PGX_CHECK(cond, ...) \
  do {                       \
    if (!(cond)) return fail(PGX_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// Index arrays are computed in int64 on the host and stored as int32 on the device
// (all counts are checked to be < 2^31 at plan creation).
std::vector<int32_t> narrow(const std::vector<int64_t>& v) { return std::vector<int32_t>(v.begin(), v.end()); }

template <typename T>
int upload(const std::vector<T>& host, T** dev, int64_t* bytes) {
  *dev = nullptr;
  if (host.empty()) return PGX_OK;
  PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(dev), host.size() * sizeof(T)));
  PGX_CUDA(cudaMemcpy(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  *bytes += int64_t(host.size() * sizeof(T));
  return PGX_OK;
}

enum EnumVariant { kPw2 = 0, kSmall = 1, kBig = 2, kUnary = 3 };
// groups of kernels whose shared-memory attribute has been raised (pgx_plan::attr_done)
enum AttrGroup { kAttrBigMax = 0, kAttrBipMax = 1, kAttrBipSum = 2, kAttrBigEnumMax = 3, kAttrBigEnumSum = 4,
                 kAttrMaxProd = 5, kAttrLattice = 6, kAttrSdlpMax = 7, kAttrSdlpSum = 8, kAttrLatticeBin = 9,
                 kAttrLatticeBinV1 = 10, kAttrLatticeBinV2 = 11, kAttrLatticeBinV3 = 12, kAttrLatticeBinV4 = 13,
                 kAttrLatticeBinV5 = 14, kAttrLatticeBinV6 = 15, kAttrLatticeBinV7 = 16, kAttrLatticeBinV8 = 17,
                 kAttrLatticeBinV9 = 18, kAttrOrAndMax = 19, kAttrOrAndSum = 20, kAttrBigSum = 21, kAttrEnumSmallMax = 22,
                 kAttrEnumSmallSum = 23, kAttrEnumCmMax = 24, kAttrEnumCmSum = 25,
                 kAttrEnumDenseMax = 26, kAttrEnumDenseSum = 27 };

struct EnumBlockPlan {
  pgx::EnumBlockDev dev{};
  EnumVariant variant = kSmall;
  int bip = -1;  // index into pgx_plan::bips when the block has the dense-grid structure
  int32_t *d_cfg_es = nullptr, *d_t_ptr = nullptr, *d_t_k = nullptr, *d_edge_off = nullptr;
  int32_t *d_fac_edge = nullptr, *d_fac_msg = nullptr, *d_fac_pot = nullptr;
  uint32_t* d_rounds = nullptr;    // round schedule of the merged max-product launch
  int32_t* d_round_ptr = nullptr;  // [num_groups + 1] first round of every lane-group
  uint32_t* d_rounds_b = nullptr;  // partner states in round order, two rounds per word (permuted-potential path)
  int32_t* d_lane_state = nullptr; // [num_groups][32] state of the first variable every lane owns
  int num_groups = 0, num_rounds = 0;
  int bigmax = -1;              // index into the plan's BigMaxGroup array, or -1
  int n0 = 0;                   // states of the first variable
  bool dense2 = false;          // arity 2 and the table is ALL n0 x n1 configurations in row-major order (k_enum_pair_dense)
};

// A pairwise-binary enum block whose factors form a dense I x J grid (see BipDev).
struct BipPlan {
  pgx::BipDev dev{};
  int32_t *d_row_vs = nullptr, *d_col_vs = nullptr, *d_row_part = nullptr, *d_col_part = nullptr;
};

struct LogicalPlan {
  pgx::LogicalDev dev{};
  int64_t max_parents = 0;
  // pull path (k_logical_pull_*): packed wiring with the "other edge" of low-degree variables
  pgx::LogicalPullDev pull{};
  pgx::EdgeW *d_parents_w = nullptr, *d_children_w = nullptr;
  int32_t* d_parent_factor = nullptr;  // [P] factor of every parent (two-launch wide update)
  int64_t num_parents = 0;
  std::vector<int32_t> h_ptr, h_pmsg, h_pvs, h_cmsg, h_cvs;  // host copies, dropped after plan creation
  std::vector<pgx::EdgeW> h_pw, h_cw;                         // packed pull wiring (same)
  bool needs_s = false;  // some edge reads its variable's sum from S
  int32_t *d_parent_ptr = nullptr, *d_parents_msg = nullptr, *d_parents_vs = nullptr, *d_children_msg = nullptr,
          *d_children_vs = nullptr;
};

struct Workspace {
  int64_t batch = 0;
  float *mA = nullptr, *mB = nullptr, *S = nullptr, *evT = nullptr, *lpT = nullptr, *part = nullptr;
  float *cA = nullptr, *cB = nullptr;  // compressed (one float per edge) messages of the fused blocks
  float* row = nullptr;                // one normalised [Es] message vector (initial messages shared by the batch)
  float* lpR = nullptr;                // round-ordered potentials of the merged max-product launch (not per batch)
  float* agg = nullptr;  // per-factor aggregates of the two-launch wide logical update
  // staging for pgx_infer_host
  float *h_lp = nullptr, *h_ev = nullptr, *h_msgs_in = nullptr, *h_msgs_out = nullptr,
        *h_marg = nullptr, *h_deltas = nullptr;
  int32_t *h_map = nullptr, *h_ties = nullptr;
  int64_t n_lp = 0, n_ev = 0, n_msgs_in = 0, n_msgs_out = 0, n_marg = 0, n_deltas = 0, n_map = 0,
          n_ties = 0;
};

// Deep copy of a graph description (kept by plans that may split off a batch tail, see
// pgx_plan::tail): the arrays the caller may free after pgx_plan_create.
struct DescCopy {
  pgx_graph_desc desc{};
  std::vector<int32_t> var_num_states, edge_var_start, edge_num_states;
  std::vector<pgx_enum_block> blocks;
  std::vector<std::vector<int32_t>> configs;
  std::vector<int32_t> lg[3][3];  // per logical desc: parents_factor, parents_msg, children_msg
  void assign(const pgx_graph_desc& d) {
    desc = d;
    var_num_states.assign(d.var_num_states, d.var_num_states + d.num_vars);
    edge_var_start.assign(d.edge_var_start, d.edge_var_start + d.num_edges);
    edge_num_states.assign(d.edge_num_states, d.edge_num_states + d.num_edges);
    desc.var_num_states = var_num_states.data();
    desc.edge_var_start = edge_var_start.data();
    desc.edge_num_states = edge_num_states.data();
    blocks.assign(d.enum_blocks, d.enum_blocks + d.num_enum_blocks);
    configs.resize(blocks.size());
    for (size_t i = 0; i < blocks.size(); ++i) {
      configs[i].assign(blocks[i].configs, blocks[i].configs + size_t(blocks[i].num_configs) * blocks[i].arity);
      blocks[i].configs = configs[i].data();
    }
    desc.enum_blocks = blocks.data();
    pgx_logical_desc* dst[3] = {&desc.or_factors, &desc.and_factors, &desc.pool_factors};
    for (int k = 0; k < 3; ++k) {
      pgx_logical_desc& l = *dst[k];
      if (l.num_factors == 0) continue;
      lg[k][0].assign(l.parents_factor, l.parents_factor + l.num_parents);
      lg[k][1].assign(l.parents_msg, l.parents_msg + l.num_parents);
      lg[k][2].assign(l.children_msg, l.children_msg + l.num_factors);
      l.parents_factor = lg[k][0].data();
      l.parents_msg = lg[k][1].data();
      l.children_msg = lg[k][2].data();
    }
  }
};

// pgx_sdlp_*: per-batch-size buffers of the smooth dual LP-MAP solver (tile-blocked)
struct SdlpWorkspace {
  int64_t batch = 0;
  float *eta = nullptr, *P = nullptr, *vval = nullptr, *eval = nullptr, *grad = nullptr;
  double* partial = nullptr;
};

}  // namespace

struct pgx_plan {
  int device = 0;
  int num_sms = 0;
  int64_t num_vars = 0, num_var_states = 0, num_edges = 0, num_edge_states = 0, num_potentials = 0;
  int64_t max_var_states = 0;
  int64_t device_bytes = 0;
  int64_t launches = 0;
  // device index structures
  int32_t* d_edge_vs = nullptr;          // [num_edges] var-state of each edge's state 0
  int32_t* d_edge_msg_start = nullptr;   // [num_edges + 1]
  int2* d_vs_csr = nullptr;              // [V_s] (CSR begin, degree << 12 | state offset) of each var-state
  int32_t* d_vs_var = nullptr;           // [V_s] variable of each var-state
  int32_t* d_var_first_state = nullptr;  // [num_vars + 1]
  int32_t* d_var_ptr = nullptr;          // [num_vars + 1] CSR offsets
  int32_t* d_var_edge_msg = nullptr;     // [num_edges] message start of incident edges, ascending
  std::vector<EnumBlockPlan> enum_blocks;
  LogicalPlan or_f, and_f, pool_f;
  Workspace ws;
  // Batch tail (PGX_PATH_TAIL_SPLIT): the OR / AND pull kernels work on full tiles of 32 samples and
  // are latency-bound, so a last tile with a few live samples costs as much as a full one
  // (deconvolution, B = 100: 0.354 ms per iteration against 0.296 at B = 96).  Those samples are
  // independent problems: they run through a second plan of the same graph (generic kernels, narrow
  // sample tiles) on its own stream beside the full tiles.
  DescCopy* desc_copy = nullptr;
  pgx_plan* tail = nullptr;
  cudaStream_t tail_stream = nullptr;
  cudaEvent_t ev_tail_fork = nullptr, ev_tail_join = nullptr;
  // smooth dual LP-MAP (pgx_sdlp_*): edges [d_factor_edge_start[f], [f + 1]) belong to factor f
  int64_t num_factors = 0;
  int32_t* d_factor_edge_start = nullptr;
  SdlpWorkspace sdlp;
  // pull mode (every factor is a pairwise-binary enum factor on low-degree variables)
  bool pull_ok = false;
  // every variable and edge is binary and every factor a pairwise-binary or unary EnumFactor: the generic
  // two-pass path can keep its messages in binary-difference storage (k_var_sums_bin + k_enum_pw2_bin)
  bool gbin_ok = false;
  int2* d_edge_csr = nullptr;          // [num_edges] CSR row (begin, end) of the edge's variable
  unsigned int* d_grid_bar = nullptr;  // barrier counter of the persistent kernel
  int coop_blocks_per_sm[2] = {0, 0};  // occupancy of k_enum_pw2_pull_resident<false / true, coop>
  // logical pull path: OR / AND kernels re-derive the sums of variables with <= 2 edges;
  // k_var_sums_list covers the rest
  bool logical_pull_ok = false;
  int32_t* d_hi_list = nullptr;
  int64_t hi_len = 0;
  int64_t hi_big = 0;  // leading variables of the list with >= kVsBigDegree edges (two-state variables only)
  // fused OR + AND launch (k_or_and_fused): every OR parent is the degree-2 child of one two-parent AND factor
  bool orand_fused_ok = false;
  pgx::OrAndFusedDev orand{};
  pgx::FusedW* d_fused_w = nullptr;
  // the smaller of the OR / AND groups runs on an auxiliary stream beside the larger one
  // (long serial parent chains of wide OR factors hide behind the bandwidth-bound AND kernel)
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // single-pass mode, >= 16 sample tiles: the two halves of the batch run as two independent
  // per-iteration chains on two streams, so that one half's dense-grid kernel fills the SMs
  // while the other half is in its short kernels (unary blocks, k_var_reduce) or in the tail of
  // its own dense-grid launch
  cudaStream_t half = nullptr;
  cudaEvent_t ev_half_fork = nullptr, ev_half_join = nullptr;
  int aux_group = 0;          // -1: OR group on aux, -2: AND group, 0: none
  bool aux_needs_s = false;   // that group reads the variable-sum array
  // merged max-product launch over all large sorted pairwise groups (k_enum_big_maxprod_all)
  pgx::BigMaxGroup* d_bigmax_groups = nullptr;
  int2* d_bigmax_units = nullptr;
  int64_t bigmax_units = 0, bigmax_es = 0;
  int64_t bigmax_perm_floats = 0;  // size of the round-ordered copy of the potentials
  bool bigmax_perm_active = false; // this run uses it (set by pgx_bp_run)
  const float* lpR_src = nullptr;  // potentials buffer the round-ordered copy was made from (PGX_RUN_POTENTIALS_UNCHANGED)
  float* d_energy_partial = nullptr;  // pgx_energy scratch
  int64_t energy_partial_floats = 0;
  size_t bigmax_smem = 0, bigsum_smem = 0;  // dynamic shared memory of the merged max- / sum-product launch
  unsigned int* d_bigmax_counter = nullptr;  // [2]: one work counter per chain of the half-batch pipeline
  // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remembered per plan (a plan
  // lives on one device), bit = kAttr* below
  uint32_t attr_done = 0;
  // lattice mode (the whole graph is one 2-D nearest-neighbour lattice block; LatticeDev)
  bool lattice_ok = false;
  pgx::LatticeDev lattice{};
  uint32_t disabled_paths = 0;         // PGX_PATH_* bits (pgx_plan_disable_paths)
  // fused single-pass structures (dense-grid pairwise blocks)
  std::vector<BipPlan> bips;
  bool exact_order = false;            // force the two-pass, serial-order path
  int64_t part_rows = 0;               // rows of the partial-sum buffer
  int64_t c_rows = 0;                  // rows of the compressed message arrays (edges of the fused blocks)
  int32_t* d_rest_ptr = nullptr;       // [num_vars + 1] CSR over edges NOT covered by a fused block
  int32_t* d_rest_edge_msg = nullptr;
  int32_t* d_part_first = nullptr;     // [num_vars] first partial row of the variable
  int32_t* d_part_count = nullptr;     // [num_vars] partial slots (each 2 rows: state 0, state 1)
  // roofline instrumentation (pgx_plan_profile_*): the dominant kernel is the f2v
  // launch covering the most edge-states; id = enum block index, or -1/-2/-3 for
  // the OR / AND / Pool launch.
  bool profiling = false;
  int dominant = 0;
  int64_t dominant_es = 0;  // edge-states the dominant launch updates
  int64_t dominant_grid = 0;  // CTAs of its last launch (bench.py matches ncu captures by kernel, batch and grid)
  const char* dominant_name = "";
  std::vector<cudaEvent_t> prof_events;  // pairs (begin, end)
  size_t prof_used = 0;
  // One CUDA graph per repeated call signature of pgx_bp_run (pgx_bp_run_flags): the launch
  // sequence of all iterations is captured the SECOND time the same (buffers, batch, iteration
  // count, scalars, path mask) is seen and replayed with one cudaGraphLaunch from then on.
  struct RunGraph {
    const void *lp, *ev, *in, *out, *deltas;
    int64_t batch;
    int32_t iters, lp_b, ev_b, in_b;
    float damping, temperature;
    uint32_t flags, paths;
    bool exact;
    cudaGraphExec_t exec;  // null: seen once, not captured yet
    int64_t launches;      // kernels + copies one replay runs
    int64_t epoch;         // workspace generation the graph's internal pointers belong to
    int64_t sums_batch;    // final_sums_batch the captured run leaves behind
  };
  int64_t ws_epoch = 0;    // bumped whenever a workspace buffer is (re)allocated: cached graphs go stale
  std::vector<RunGraph> run_graphs;
  cudaStream_t cap_stream = nullptr;
  int64_t final_sums_batch = 0;  // batch whose FINAL variable sums (= beliefs) the last run left in ws.S (PGX_RUN_FINAL_SUMS)
  int64_t graph_launches = 0;
  bool graphs_enabled = true;
};

namespace {

void free_dev(void* p) {
  if (p) cudaFree(p);
}

void free_workspace(Workspace& ws) {
  free_dev(ws.mA); free_dev(ws.mB); free_dev(ws.S); free_dev(ws.evT); free_dev(ws.lpT); free_dev(ws.part);
  free_dev(ws.agg); free_dev(ws.cA); free_dev(ws.cB); free_dev(ws.row); free_dev(ws.lpR);
  free_dev(ws.h_lp); free_dev(ws.h_ev); free_dev(ws.h_msgs_in); free_dev(ws.h_msgs_out);
  free_dev(ws.h_marg); free_dev(ws.h_deltas); free_dev(ws.h_map); free_dev(ws.h_ties);
  ws = Workspace{};
}

pgx::BatchMap make_map(int64_t batch) {
  pgx::BatchMap mp;
  mp.batch = int(batch);
  int lg = 0;
  while ((1 << lg) < batch && lg < 5) ++lg;
  mp.bx_log = lg;
  mp.nbt = int((batch + (1 << lg) - 1) >> lg);
  return mp;
}

// Floats of a tile-blocked array with n_rows elements per sample.
size_t tiled_floats(const pgx::BatchMap& mp, int64_t n_rows) {
  return (size_t(std::max<int64_t>(n_rows, 1)) * mp.nbt) << mp.bx_log;
}

// Grid for a "one thread per (element, sample)" kernel: y = sample tiles; x = enough
// CTAs to cover the elements of one tile, capped so that the whole grid is a few CTAs
// per SM (warps then loop over contiguous chunks).
dim3 grid_for(const pgx_plan* plan, const pgx::BatchMap& mp, int64_t units) {
  const int upw = 32 >> mp.bx_log;
  const int wpb = pgx::kThreads / 32;
  const int64_t warps = std::max<int64_t>(1, (units + upw - 1) / upw);
  int64_t blocks = (warps + wpb - 1) / wpb;
  const int64_t cap = std::max<int64_t>(1, (int64_t(plan->num_sms) * 8 + mp.nbt - 1) / mp.nbt);
  blocks = std::min<int64_t>(blocks, cap);
  return dim3(unsigned(blocks), unsigned(mp.nbt), 1);
}

int build_logical(pgx_plan* plan, const pgx_logical_desc& d, const std::vector<int64_t>& edge_msg_start,
                  const std::vector<int32_t>& edge_vs, const std::vector<int32_t>& edge_ns,
                  std::vector<uint8_t>& edge_covered, const char* what, LogicalPlan* out) {
  out->dev.num_factors = d.num_factors;
  out->dev.off = d.edge_states_offset;
  if (d.num_factors == 0) return PGX_OK;
  PGX_CHECK(d.edge_states_offset == 1 || d.edge_states_offset == -1,
            "%s: edge_states_offset must be +1 or -1, got %d", what, d.edge_states_offset);
  PGX_CHECK(d.parents_factor && d.parents_msg && d.children_msg, "%s: null wiring array", what);
  const int rel = d.edge_states_offset > 0 ? 0 : 1;  // state the wiring points at
  std::vector<int64_t> ptr(d.num_factors + 1, 0);
  for (int64_t i = 0; i < d.num_parents; ++i) {
    const int32_t f = d.parents_factor[i];
    PGX_CHECK(f >= 0 && f < d.num_factors, "%s: parent %lld has factor index %d out of range", what,
              (long long)i, f);
    PGX_CHECK(i == 0 || d.parents_factor[i - 1] <= f, "%s: parents must be grouped by ascending factor",
              what);
    ++ptr[f + 1];
  }
  for (int64_t f = 0; f < d.num_factors; ++f) {
    PGX_CHECK(ptr[f + 1] >= 1, "%s: factor %lld has no parent", what, (long long)f);
    ptr[f + 1] += ptr[f];
  }
  auto resolve = [&](int32_t msg, int32_t* vs) -> int {
    // edge containing message index `msg`
    const int64_t lo_msg = int64_t(msg) - rel;
    auto it = std::upper_bound(edge_msg_start.begin(), edge_msg_start.end(), lo_msg);
    const int64_t e = (it - edge_msg_start.begin()) - 1;
    PGX_CHECK(e >= 0 && e < plan->num_edges && edge_msg_start[e] == lo_msg && edge_ns[e] == 2,
              "%s: message index %d is not state %d of a binary edge", what, msg, rel);
    PGX_CHECK(!edge_covered[e], "%s: edge %lld is updated by more than one factor", what, (long long)e);
    edge_covered[e] = 1;
    *vs = edge_vs[e] + rel;
    return PGX_OK;
  };
  std::vector<int32_t> pvs(d.num_parents), cvs(d.num_factors);
  for (int64_t i = 0; i < d.num_parents; ++i)
    if (int rc = resolve(d.parents_msg[i], &pvs[i])) return rc;
  for (int64_t f = 0; f < d.num_factors; ++f)
    if (int rc = resolve(d.children_msg[f], &cvs[f])) return rc;
  std::vector<int32_t> pmsg(d.parents_msg, d.parents_msg + d.num_parents);
  std::vector<int32_t> cmsg(d.children_msg, d.children_msg + d.num_factors);
  int rc;
  out->h_ptr = narrow(ptr);
  out->h_pmsg = pmsg;
  out->h_pvs = pvs;
  out->h_cmsg = cmsg;
  out->h_cvs = cvs;
  if ((rc = upload(narrow(ptr), &out->d_parent_ptr, &plan->device_bytes))) return rc;
  if ((rc = upload(pmsg, &out->d_parents_msg, &plan->device_bytes))) return rc;
  if ((rc = upload(pvs, &out->d_parents_vs, &plan->device_bytes))) return rc;
  if ((rc = upload(cmsg, &out->d_children_msg, &plan->device_bytes))) return rc;
  if ((rc = upload(cvs, &out->d_children_vs, &plan->device_bytes))) return rc;
  {
    const int64_t n0 = ptr[1] - ptr[0];
    bool uniform = true;
    for (int64_t f = 0; f < d.num_factors && uniform; ++f) uniform = ptr[f + 1] - ptr[f] == n0;
    out->dev.uniform = uniform ? int32_t(n0) : 0;
    for (int64_t f = 0; f < d.num_factors; ++f) out->max_parents = std::max<int64_t>(out->max_parents, ptr[f + 1] - ptr[f]);
  }
  out->dev.parent_ptr = out->d_parent_ptr;
  out->dev.parents_msg = out->d_parents_msg;
  out->dev.parents_vs = out->d_parents_vs;
  out->dev.children_msg = out->d_children_msg;
  out->dev.children_vs = out->d_children_vs;
  return PGX_OK;
}

// Checks one pgx_enum_block against the edge table and marks its edges as covered.
int check_enum_block(pgx_plan* plan, const pgx_enum_block& b, int idx, const std::vector<int64_t>& edge_msg_start,
                     const std::vector<int32_t>& edge_ns, std::vector<uint8_t>& edge_covered) {
  PGX_CHECK(b.num_factors >= 1 && b.arity >= 1 && b.num_configs >= 0 && b.configs != nullptr,
            "enum block %d: bad sizes", idx);
  const int A = b.arity, K = b.num_configs;
  PGX_CHECK(b.first_edge >= 0 && b.first_edge + b.num_factors * A <= plan->num_edges,
            "enum block %d: edge range out of bounds", idx);
  PGX_CHECK(edge_msg_start[b.first_edge] == b.first_msg,
            "enum block %d: first_msg %lld does not match the edge table (%lld)", idx,
            (long long)b.first_msg, (long long)edge_msg_start[b.first_edge]);
  PGX_CHECK(b.first_potential >= 0 && b.first_potential + b.num_factors * K <= plan->num_potentials,
            "enum block %d: potential range out of bounds", idx);
  for (int64_t f = 0; f < b.num_factors; ++f)
    for (int a = 0; a < A; ++a) {
      const int64_t e = b.first_edge + f * A + a;
      if (edge_ns[e] != edge_ns[b.first_edge + a])
        return fail(PGX_ERR_UNSUPPORTED,
                    "enum block %d: factors with different numbers of states share one block; "
                    "split the group per state signature", idx);
      PGX_CHECK(!edge_covered[e], "enum block %d: edge %lld is updated by more than one factor", idx,
                (long long)e);
      edge_covered[e] = 1;
    }
  return PGX_OK;
}

// Builds the device structures of one group of enum blocks that share the configuration
// table and the per-variable state counts (`members` index desc_blocks; one member = the
// usual case of an EnumFactorGroup; many members = e.g. the RCN graphs, where every factor
// is its own group but only a handful of distinct tables exist, examples/rcn.ipynb cell 26).
int build_enum_block(pgx_plan* plan, const pgx_enum_block* desc_blocks, const std::vector<int>& members,
                     const std::vector<int64_t>& edge_msg_start, const std::vector<int32_t>& edge_ns,
                     EnumBlockPlan* out) {
  const pgx_enum_block& b = desc_blocks[members[0]];
  const int idx = members[0];
  const int A = b.arity, K = b.num_configs;
  std::vector<int32_t> edge_off(A + 1, 0);
  for (int a = 0; a < A; ++a) edge_off[a + 1] = edge_off[a] + edge_ns[b.first_edge + a];
  const int ns = edge_off[A];
  std::vector<int32_t> cfg_es(size_t(K) * A), t_ptr(ns + 1, 0), t_k(size_t(K) * A);
  for (int k = 0; k < K; ++k)
    for (int a = 0; a < A; ++a) {
      const int32_t st = b.configs[size_t(k) * A + a];
      PGX_CHECK(st >= 0 && st < edge_off[a + 1] - edge_off[a],
                "enum block %d: configuration %d assigns state %d to variable %d", idx, k, st, a);
      cfg_es[size_t(k) * A + a] = edge_off[a] + st;
      ++t_ptr[edge_off[a] + st + 1];
    }
  for (int s = 0; s < ns; ++s) t_ptr[s + 1] += t_ptr[s];
  {
    std::vector<int32_t> cursor(t_ptr.begin(), t_ptr.end() - 1);
    for (int k = 0; k < K; ++k)  // ascending k within every list
      for (int a = 0; a < A; ++a) t_k[cursor[cfg_es[size_t(k) * A + a]]++] = k;
  }
  const bool full_binary_pair =
      A == 2 && K == 4 && ns == 4 && b.configs[0] == 0 && b.configs[1] == 0 && b.configs[2] == 0 &&
      b.configs[3] == 1 && b.configs[4] == 1 && b.configs[5] == 0 && b.configs[6] == 1 &&
      b.configs[7] == 1;
  bool full_unary = A == 1 && K == ns;  // one variable, configuration k = state k
  for (int k = 0; k < K && full_unary; ++k) full_unary = b.configs[k] == k;
  out->variant = full_binary_pair ? kPw2 : (full_unary ? kUnary : (ns <= pgx::kSmallMaxNS ? kSmall : kBig));
  if (out->variant == kBig && size_t(2 * ns + 32) * sizeof(float) > 227 * 1024)
    return fail(PGX_ERR_UNSUPPORTED, "enum block %d: %d edge-states per factor exceed shared memory",
                idx, ns);
  int rc;
  if ((rc = upload(cfg_es, &out->d_cfg_es, &plan->device_bytes))) return rc;
  if ((rc = upload(t_ptr, &out->d_t_ptr, &plan->device_bytes))) return rc;
  if ((rc = upload(t_k, &out->d_t_k, &plan->device_bytes))) return rc;
  if ((rc = upload(edge_off, &out->d_edge_off, &plan->device_bytes))) return rc;
  int64_t total = 0;
  for (int m : members) total += desc_blocks[m].num_factors;
  if (members.size() > 1) {  // factors are not an arithmetic progression: per-factor offsets
    std::vector<int32_t> fe, fm, fp;
    fe.reserve(total); fm.reserve(total); fp.reserve(total);
    for (int m : members) {
      const pgx_enum_block& mb = desc_blocks[m];
      for (int64_t f = 0; f < mb.num_factors; ++f) {
        fe.push_back(int32_t(mb.first_edge + f * A));
        fm.push_back(int32_t(mb.first_msg + f * ns));
        fp.push_back(int32_t(mb.first_potential + f * K));
      }
    }
    if ((rc = upload(fe, &out->d_fac_edge, &plan->device_bytes))) return rc;
    if ((rc = upload(fm, &out->d_fac_msg, &plan->device_bytes))) return rc;
    if ((rc = upload(fp, &out->d_fac_pot, &plan->device_bytes))) return rc;
  }
  if (A == 2) {  // the complete table in row-major order?
    const int n0 = edge_off[1], n1 = ns - edge_off[1];
    bool dense = int64_t(K) == int64_t(n0) * n1;
    for (int k = 0; k < K && dense; ++k)
      dense = cfg_es[size_t(k) * 2] == k / n1 && cfg_es[size_t(k) * 2 + 1] == n0 + k % n1;
    out->dense2 = dense;
    out->n0 = n0;
  }
  {  // configs sorted by the first variable's state (arity 2)?
    bool sorted0 = A == 2;
    for (int s0 = 0; s0 < edge_off[1] && sorted0; ++s0)
      for (int j = t_ptr[s0]; j < t_ptr[s0 + 1] && sorted0; ++j) sorted0 = t_k[j] == j;
    out->dev.sorted0 = sorted0 ? 1 : 0;
    // merged max-product launch: additionally the partner states of every state's list must be
    // strictly ascending (no duplicate configuration: lanes of one step update distinct slots)
    bool strict = sorted0 && out->variant == kBig && ns - edge_off[1] <= 65535 && edge_off[1] <= 65535;
    for (int s0 = 0; s0 < edge_off[1] && strict; ++s0)
      for (int k = t_ptr[s0] + 1; k < t_ptr[s0 + 1] && strict; ++k)
        strict = cfg_es[size_t(k) * 2 + 1] > cfg_es[size_t(k - 1) * 2 + 1];
    // ... and a configuration packs into 32 bits (k: 20 bits, partner state: 12 bits)
    strict = strict && K < (1 << 20) && ns - edge_off[1] < 4096;
    if (strict) {
      // Round schedule (k_enum_big_maxprod_all): lane l of lane-group g owns state a = 32 g + l of
      // the first variable and walks its configuration list in rounds; within a round the
      // partner states of the 32 lanes are pairwise distinct (greedy: a lane whose next partner
      // state is already taken in this round idles), so the partner-side maxima can be updated
      // with a plain read-max-write.  Entry = k | partner << 20, 0xffffffff = idle.
      const int n0 = edge_off[1];
      const int num_groups = (n0 + 31) / 32;
      // Which 32 states share a lane-group is free (a lane only needs ITS state's list): groups of
      // states with lists of equal length finish together, and states with distinct indices mod 32
      // have translated partner sets in distinct banks.  Candidates: contiguous states (the first
      // version), states sorted by list length, and sorted with distinct residues mod 32 inside a
      // look-ahead window; the schedule with the fewest rounds wins (RCN tables: 80 - 88 % of the
      // lane slots used -> 90 - 94 %).
      auto list_len = [&](int a) { return t_ptr[a + 1] - t_ptr[a]; };
      auto make_grouping = [&](int mode) {
        std::vector<int32_t> lane_state(size_t(num_groups) * 32, n0);
        if (mode == 0) {
          for (int a = 0; a < n0; ++a) lane_state[a] = a;
          return lane_state;
        }
        std::vector<int32_t> order(n0);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return list_len(x) > list_len(y); });
        if (mode == 1) {
          for (int i = 0; i < n0; ++i) lane_state[i] = order[i];
          return lane_state;
        }
        const int window = mode == 2 ? 64 : 160;
        std::vector<uint8_t> used(n0, 0);
        int pos = 0, slot = 0;
        while (pos < n0) {
          while (pos < n0 && used[order[pos]]) ++pos;
          if (pos >= n0) break;
          uint32_t residues = 0;
          int filled = 0;
          for (int i = pos; i < std::min(n0, pos + window) && filled < 32; ++i) {
            const int a = order[i];
            if (used[a] || ((residues >> (a & 31)) & 1u)) continue;
            residues |= 1u << (a & 31);
            used[a] = 1;
            lane_state[slot + filled++] = a;
          }
          for (int i = pos; i < n0 && filled < 32; ++i) {
            const int a = order[i];
            if (used[a]) continue;
            used[a] = 1;
            lane_state[slot + filled++] = a;
          }
          slot += 32;
        }
        return lane_state;
      };
      std::vector<uint32_t> rounds;
      std::vector<uint8_t> idle_bank;  // per entry: the free bank an idle lane's dummy access goes to
      std::vector<int32_t> round_ptr(num_groups + 1, 0);
      auto build_schedule = [&](const std::vector<int32_t>& lane_state) {
      rounds.clear();
      idle_bank.clear();
      std::fill(round_ptr.begin(), round_ptr.end(), 0);
      for (int g = 0; g < num_groups; ++g) {
        // remaining configurations (k, partner state) of every lane; the order inside a list is
        // free (max is order-independent), so a lane takes ANY remaining configuration whose
        // shared-memory bank (partner state mod 32) is still free in this round: distinct banks
        // = distinct partner states and conflict-free accesses.  Lanes with the most work left
        // choose first.
        std::vector<std::pair<int32_t, int32_t>> rem[32];
        for (int l = 0; l < 32; ++l) {
          const int a = lane_state[size_t(g) * 32 + l];
          if (a >= n0) continue;
          for (int k = t_ptr[a + 1] - 1; k >= t_ptr[a]; --k) rem[l].push_back({k, cfg_es[size_t(k) * 2 + 1] - n0});
        }
        for (;;) {
          int order[32];
          size_t left = 0;
          for (int l = 0; l < 32; ++l) { order[l] = l; left += rem[l].size(); }
          if (left == 0) break;
          std::stable_sort(order, order + 32, [&](int x, int y) { return rem[x].size() > rem[y].size(); });
          uint32_t entry[32];
          uint32_t banks = 0;
          for (int l = 0; l < 32; ++l) entry[l] = 0xffffffffu;
          for (int oi = 0; oi < 32; ++oi) {
            const int l = order[oi];
            for (size_t i = rem[l].size(); i-- > 0;) {  // from the back: ascending k first
              const int32_t pb = rem[l][i].second;
              if ((banks >> (pb & 31)) & 1u) continue;
              banks |= 1u << (pb & 31);
              entry[l] = uint32_t(rem[l][i].first) | (uint32_t(pb) << 20);
              rem[l].erase(rem[l].begin() + i);
              break;
            }
          }
          // idle lanes read-max-write a dummy slot n1 + j: give each one a bank nobody uses
          int fb = 0;
          for (int l = 0; l < 32; ++l)
            if (entry[l] == 0xffffffffu) {
              while ((banks >> fb) & 1u) ++fb;
              idle_bank.push_back(uint8_t(fb));
              banks |= 1u << fb;
            } else {
              idle_bank.push_back(255);
            }
          rounds.insert(rounds.end(), entry, entry + 32);
        }
        // even number of rounds per lane-group (the partner states are fetched two rounds at a time)
        if ((rounds.size() / 32 - size_t(round_ptr[g])) & 1) {
          rounds.insert(rounds.end(), 32, 0xffffffffu);
          for (int l = 0; l < 32; ++l) idle_bank.push_back(uint8_t(l));
        }
        round_ptr[g + 1] = int32_t(rounds.size() / 32);
      }
      };
      std::vector<int32_t> best_grouping;
      {
        size_t best_rounds = 0;
        int best_mode = 0;
        const int modes = n0 > 32 ? 4 : 1;
        for (int mode = 0; mode < modes; ++mode) {
          const std::vector<int32_t> grouping = make_grouping(mode);
          build_schedule(grouping);
          if (mode == 0 || rounds.size() < best_rounds) { best_rounds = rounds.size(); best_mode = mode; best_grouping = grouping; }
        }
        if (best_mode != modes - 1) build_schedule(best_grouping);
      }
      if ((rc = upload(best_grouping, &out->d_lane_state, &plan->device_bytes))) return rc;
      if (rounds.empty()) rounds.push_back(0xffffffffu);
      {
        // [round pair][lane]: partner state of round 2p | partner state of round 2p + 1 << 16
        // dummy slot of an idle entry: n1 + j with (n1 + j) mod 32 = its free bank
        const int n1 = ns - n0;
        auto pb = [&](size_t i) {
          if (rounds[i] != 0xffffffffu) return rounds[i] >> 20;
          const int bank = i < idle_bank.size() ? idle_bank[i] : int(i & 31);
          return uint32_t(n1 + ((bank - n1) % 32 + 32) % 32);
        };
        std::vector<uint32_t> rb(std::max<size_t>(rounds.size() / 2, 32));
        for (size_t pr = 0; pr + 1 < rounds.size() / 32 + 1 && pr * 64 + 63 < rounds.size(); ++pr)
          for (size_t l = 0; l < 32; ++l) rb[pr * 32 + l] = (pb(pr * 64 + l) << 2) | (pb(pr * 64 + 32 + l) << 18);  // byte offsets
        if ((rc = upload(rb, &out->d_rounds_b, &plan->device_bytes))) return rc;
      }
      out->num_rounds = int32_t(rounds.size() / 32);
      if ((rc = upload(rounds, &out->d_rounds, &plan->device_bytes))) return rc;
      if ((rc = upload(round_ptr, &out->d_round_ptr, &plan->device_bytes))) return rc;
      out->num_groups = num_groups;
      out->bigmax = 0;  // index assigned by the caller
      out->n0 = edge_off[1];
    }
  }
  out->dev.fac_edge = out->d_fac_edge;
  out->dev.fac_msg = out->d_fac_msg;
  out->dev.fac_pot = out->d_fac_pot;
  out->dev.num_factors = total;
  out->dev.first_edge = b.first_edge;
  out->dev.first_msg = b.first_msg;
  out->dev.first_pot = b.first_potential;
  out->dev.arity = A;
  out->dev.num_configs = K;
  out->dev.ns = ns;
  out->dev.cfg_es = out->d_cfg_es;
  out->dev.t_ptr = out->d_t_ptr;
  out->dev.t_k = out->d_t_k;
  out->dev.edge_off = out->d_edge_off;
  return PGX_OK;
}

int ensure_workspace(pgx_plan* plan, int64_t batch, bool need_evT, bool need_lpT, bool need_part,
                     bool need_lbin = false) {
  Workspace& ws = plan->ws;
  const pgx::BatchMap mp = make_map(batch);
  if (ws.batch != batch) {
    free_dev(ws.mA); free_dev(ws.mB); free_dev(ws.S); free_dev(ws.evT); free_dev(ws.lpT); free_dev(ws.part);
    free_dev(ws.agg); free_dev(ws.cA); free_dev(ws.cB); free_dev(ws.row);
    ws.mA = ws.mB = ws.S = ws.evT = ws.lpT = ws.part = ws.agg = ws.cA = ws.cB = ws.row = nullptr;
    ws.batch = -1;  // valid only once every allocation below has succeeded (a failed call is retried in full)
    ++plan->ws_epoch;
    const size_t nm = tiled_floats(mp, plan->num_edge_states) * sizeof(float);
    const size_t nv = tiled_floats(mp, plan->num_var_states) * sizeof(float);
    PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(&ws.mA), nm));
    PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(&ws.mB), nm));
    PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(&ws.S), nv));
    PGX_CUDA(cudaMemset(ws.S, 0, nv));  // padded sample slots stay finite
    ws.batch = batch;
  }
  if (plan->logical_pull_ok && mp.bx_log == 5 && ws.agg == nullptr) {
    int64_t wide = 0;  // factors of the groups that take the two-launch wide update
    for (const LogicalPlan* lg : {&plan->or_f, &plan->and_f})
      if (lg->max_parents > pgx::kRegParents) wide = std::max<int64_t>(wide, lg->dev.num_factors);
    if (wide > 0) {
      PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(&ws.agg), size_t(wide) * pgx::kAggRows * 32 * mp.nbt * sizeof(float)));
      ++plan->ws_epoch;
    }
  }
  // lazily allocated buffers: each guarded by its own pointer (a failed allocation leaves the rest retryable)
  auto lazy = [plan](float** p, size_t bytes) -> int {
    if (*p != nullptr) return PGX_OK;
    PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(p), bytes));
    ++plan->ws_epoch;
    return PGX_OK;
  };
  int rc;
  if (need_part) {
    if ((rc = lazy(&ws.part, tiled_floats(mp, plan->part_rows) * sizeof(float)))) return rc;
    if ((rc = lazy(&ws.cA, tiled_floats(mp, plan->c_rows) * sizeof(float)))) return rc;
    if ((rc = lazy(&ws.cB, tiled_floats(mp, plan->c_rows) * sizeof(float)))) return rc;
    if ((rc = lazy(&ws.row, size_t(plan->num_edge_states) * sizeof(float)))) return rc;
  }
  if (need_lbin) {
    const size_t nc = tiled_floats(mp, plan->num_edge_states / 2) * sizeof(float);
    if ((rc = lazy(&ws.cA, nc))) return rc;
    if ((rc = lazy(&ws.cB, nc))) return rc;
  }
  if (need_evT && (rc = lazy(&ws.evT, tiled_floats(mp, plan->num_var_states) * sizeof(float)))) return rc;
  if (need_lpT && (rc = lazy(&ws.lpT, tiled_floats(mp, plan->num_potentials) * sizeof(float)))) return rc;
  return PGX_OK;
}

int check_launch(pgx_plan* plan, const char* what) {
  ++plan->launches;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(PGX_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(err));
  return PGX_OK;
}

int to_tiles(pgx_plan* plan, cudaStream_t st, const float* src, float* dst, int64_t n, const pgx::BatchMap& mp) {
  if (n == 0) return PGX_OK;
  const int padded = mp.nbt << mp.bx_log;
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((padded + 31) / 32)), block(32, 8);
  pgx::k_to_tiles<<<grid, block, 0, st>>>(src, dst, n, mp);
  return check_launch(plan, "k_to_tiles");
}

// rows [n_begin, n_end) of the n-row arrays
int from_tiles(pgx_plan* plan, cudaStream_t st, const float* src, float* dst, int64_t n, const pgx::BatchMap& mp,
               int64_t n_begin = 0, int64_t n_end = -1) {
  if (n_end < 0) n_end = n;
  if (n_end <= n_begin) return PGX_OK;
  dim3 grid((unsigned)((n_end - n_begin + 31) / 32), (unsigned)((mp.batch + 31) / 32)), block(32, 8);
  pgx::k_from_tiles<<<grid, block, 0, st>>>(src, dst, n, n_begin, n_end, mp);
  return check_launch(plan, "k_from_tiles");
}

// normalize_and_clip_msgs on the staged input messages (bp.py:92-96), in place.
int normalize_edges(pgx_plan* plan, cudaStream_t st, const pgx::BatchMap& mp, float* m) {
  if (plan->num_edges == 0) return PGX_OK;
  if (mp.batch == 1 && plan->num_edge_states >= 16 * plan->num_edges) {  // wide edges: a warp per edge
    const int64_t blocks = std::min<int64_t>((plan->num_edges + 7) / 8, int64_t(plan->num_sms) * 16);
    pgx::k_normalize_edges_warp<<<unsigned(blocks), pgx::kThreads, 0, st>>>(plan->num_edges, plan->d_edge_msg_start, m);
    return check_launch(plan, "k_normalize_edges_warp");
  }
  pgx::k_normalize_edges<<<grid_for(plan, mp, plan->num_edges), pgx::kThreads, 0, st>>>(
      mp, plan->num_edges, plan->num_edge_states, plan->d_edge_msg_start, m);
  return check_launch(plan, "k_normalize_edges");
}

// The fused blocks in ascending message order.
std::vector<const BipPlan*> bips_by_msg(const pgx_plan* plan) {
  std::vector<const BipPlan*> order;
  for (const BipPlan& bp : plan->bips) order.push_back(&bp);
  std::sort(order.begin(), order.end(), [](const BipPlan* x, const BipPlan* y) { return x->dev.first_msg < y->dev.first_msg; });
  return order;
}

// Records a profiling event if `id` is the plan's dominant launch.
int prof_mark(pgx_plan* plan, cudaStream_t st, int id) {
  if (!plan->profiling || id != plan->dominant) return PGX_OK;
  if (plan->prof_used == plan->prof_events.size()) {
    cudaEvent_t e;
    PGX_CUDA(cudaEventCreate(&e));
    plan->prof_events.push_back(e);
  }
  PGX_CUDA(cudaEventRecord(plan->prof_events[plan->prof_used++], st));
  return PGX_OK;
}

template <bool kSum>
int launch_f2v(pgx_plan* plan, cudaStream_t st, const pgx::BatchMap& mp, pgx::View lp, const float* S,
               const float* m_old, float* m_new, const pgx::RunArgs& a, bool fused, bool lpull, pgx::View ev,
               cudaStream_t aux, const float* c_old = nullptr, float* c_new = nullptr, bool lbin = false,
               float* part_override = nullptr, int chain = 0, bool gbin = false) {
  int rc;
  auto attr_needed = [plan](int group) {
    const bool need = !((plan->attr_done >> group) & 1u);
    plan->attr_done |= 1u << group;
    return need;
  };
  const bool merged_max = !kSum && plan->bigmax_units > 0 && !(plan->disabled_paths & PGX_PATH_MERGED_MAX);
  if (merged_max) {
    const bool dom = plan->dominant >= 0 && size_t(plan->dominant) < plan->enum_blocks.size() &&
                     plan->enum_blocks[plan->dominant].bigmax >= 0;
    if (attr_needed(kAttrBigMax)) {
      PGX_CUDA(cudaFuncSetAttribute(pgx::k_enum_big_maxprod_all<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      PGX_CUDA(cudaFuncSetAttribute(pgx::k_enum_big_maxprod_all<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      PGX_CUDA(cudaFuncSetAttribute(pgx::k_enum_big_maxprod_all<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    // the two chains of the half-batch pipeline run this launch concurrently: a counter each
    unsigned int* const counter = plan->d_bigmax_counter + chain;
    PGX_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    if (dom && (rc = prof_mark(plan, st, plan->dominant))) return rc;
    const int per_sm = int(std::max<size_t>(1, std::min<size_t>(6, (200 * 1024) / plan->bigmax_smem)));
    const int64_t grid = std::min<int64_t>(plan->bigmax_units * mp.batch, int64_t(plan->num_sms) * per_sm);
    const float* lpR = plan->ws.lpR;
#define PGX_BIGMAX(FLAT, PERM)                                                                                     \
  pgx::k_enum_big_maxprod_all<FLAT, PERM><<<unsigned(grid), pgx::kThreads, plan->bigmax_smem, st>>>(                \
      mp, plan->d_bigmax_groups, plan->d_bigmax_units, plan->bigmax_units, counter, plan->d_edge_vs, \
      lp, lpR, S, m_old, m_new, a)
    if (plan->bigmax_perm_active) PGX_BIGMAX(true, true);
    else if (lp.kind != 1) PGX_BIGMAX(true, false);
    else PGX_BIGMAX(false, false);
#undef PGX_BIGMAX
    if ((rc = check_launch(plan, "k_enum_big_maxprod_all"))) return rc;
    if (dom) {
      plan->dominant_name = "k_enum_big_maxprod_all";
      plan->dominant_grid = grid;
      if ((rc = prof_mark(plan, st, plan->dominant))) return rc;
    }
  }
  // ... and its sum-product sibling, on the round-ordered potentials only
  const bool merged_sum = kSum && plan->bigmax_units > 0 && plan->bigmax_perm_active && plan->bigsum_smem <= 200 * 1024 &&
                          !(plan->disabled_paths & PGX_PATH_MERGED_MAX);
  if (merged_sum) {
    const bool dom = plan->dominant >= 0 && size_t(plan->dominant) < plan->enum_blocks.size() &&
                     plan->enum_blocks[plan->dominant].bigmax >= 0;
    if (attr_needed(kAttrBigSum))
      PGX_CUDA(cudaFuncSetAttribute(pgx::k_enum_big_sumprod_all, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    unsigned int* const counter = plan->d_bigmax_counter + chain;
    PGX_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    if (dom && (rc = prof_mark(plan, st, plan->dominant))) return rc;
    const int per_sm = int(std::max<size_t>(1, std::min<size_t>(6, (227 * 1024) / (plan->bigsum_smem + 1024))));
    const int64_t grid = std::min<int64_t>(plan->bigmax_units * mp.batch, int64_t(plan->num_sms) * per_sm);
    pgx::k_enum_big_sumprod_all<<<unsigned(grid), pgx::kThreads, plan->bigsum_smem, st>>>(
        mp, plan->d_bigmax_groups, plan->d_bigmax_units, plan->bigmax_units, counter, plan->d_edge_vs, plan->ws.lpR, S,
        m_old, m_new, a);
    if ((rc = check_launch(plan, "k_enum_big_sumprod_all"))) return rc;
    if (dom) {
      plan->dominant_name = "k_enum_big_sumprod_all";
      plan->dominant_grid = grid;
      if ((rc = prof_mark(plan, st, plan->dominant))) return rc;
    }
  }
  for (size_t bi = 0; bi < plan->enum_blocks.size(); ++bi) {
    EnumBlockPlan& eb = plan->enum_blocks[bi];
    const int64_t F = eb.dev.num_factors;
    if ((merged_max || merged_sum) && eb.bigmax >= 0) continue;
    // fused mode: the blocks outside the dense-grid kernel run beside it on the auxiliary stream
    const cudaStream_t main_enum_st = st;
    const cudaStream_t st = (fused && aux != nullptr && !lpull && eb.bip < 0) ? aux : main_enum_st;  // NOLINT
    if ((rc = prof_mark(plan, st, int(bi)))) return rc;
    if (fused && eb.bip >= 0) {
      const pgx::BipDev& g = plan->bips[eb.bip].dev;
      constexpr int TJ = pgx::kBipTJ;
      // c_old == nullptr: first iteration of a run, the input rows are in the full layout
      const bool in_full = c_old == nullptr;
      // 4-warp CTAs when the last group of 8 sample tiles would be at most half full and the batch is
      // small enough for that to matter (strong-scaled shards: 128 samples = 4 tiles)
      const bool narrow = !in_full && mp.nbt <= 12 && ((mp.nbt + 7) / 8) * 8 - mp.nbt >= 4;
      const int warps = pgx::bip_warps(in_full, narrow);
      const int groups = (mp.nbt + warps - 1) / warps;
      const int64_t grid = int64_t(g.NS) * g.NR * groups;
      const size_t smem = pgx::bip_smem_bytes(g.RI, TJ, in_full, narrow);
      if (attr_needed(kSum ? kAttrBipSum : kAttrBipMax)) {
#define PGX_BIP_ATTR(DELTA, FULL, NARROW)                                                                   \
  PGX_CUDA(cudaFuncSetAttribute(pgx::k_enum_pw2_bip<kSum, TJ, DELTA, FULL, NARROW>,                          \
                                cudaFuncAttributeMaxDynamicSharedMemorySize, int(pgx::bip_smem_bytes(32, TJ, FULL, NARROW))))
        PGX_BIP_ATTR(true, true, false); PGX_BIP_ATTR(true, false, false); PGX_BIP_ATTR(false, true, false);
        PGX_BIP_ATTR(false, false, false); PGX_BIP_ATTR(true, false, true); PGX_BIP_ATTR(false, false, true);
#undef PGX_BIP_ATTR
      }
      const float* src = in_full ? m_old : c_old;
      const int64_t src_rows = in_full ? a.Es : plan->c_rows;
#define PGX_BIP_LAUNCH(DELTA, FULL, NARROW)                                                                 \
  pgx::k_enum_pw2_bip<kSum, TJ, DELTA, FULL, NARROW><<<unsigned(grid), warps * 32, smem, st>>>(              \
      mp.batch, groups, g, lp.p, S, src, src_rows, c_new, plan->c_rows,                                    \
      part_override ? part_override : plan->ws.part, plan->part_rows, a)
      if (a.deltas != nullptr) {
        if (in_full) PGX_BIP_LAUNCH(true, true, false);
        else if (narrow) PGX_BIP_LAUNCH(true, false, true);
        else PGX_BIP_LAUNCH(true, false, false);
      } else {
        if (in_full) PGX_BIP_LAUNCH(false, true, false);
        else if (narrow) PGX_BIP_LAUNCH(false, false, true);
        else PGX_BIP_LAUNCH(false, false, false);
      }
#undef PGX_BIP_LAUNCH
      if ((rc = check_launch(plan, "k_enum_pw2_bip"))) return rc;
      if (int(bi) == plan->dominant) { plan->dominant_name = "k_enum_pw2_bip"; plan->dominant_grid = grid; }
    } else if (eb.variant == kPw2) {
      if (gbin) {  // m_old / m_new are the one-float-per-edge arrays
        pgx::k_enum_pw2_bin<kSum><<<grid_for(plan, mp, F), pgx::kThreads, 0, st>>>(
            mp, F, eb.dev.first_edge, eb.dev.first_pot, a.Es / 2, plan->d_edge_vs, lp, S, m_old, m_new, a);
        if ((rc = check_launch(plan, "k_enum_pw2_bin"))) return rc;
        if (int(bi) == plan->dominant) plan->dominant_name = "k_enum_pw2_bin";
      } else {
        pgx::k_enum_pw2<kSum><<<grid_for(plan, mp, F), pgx::kThreads, 0, st>>>(
            mp, F, eb.dev.first_edge, eb.dev.first_msg, eb.dev.first_pot, plan->d_edge_vs, lp, S, m_old,
            m_new, a);
        if ((rc = check_launch(plan, "k_enum_pw2"))) return rc;
        if (int(bi) == plan->dominant) plan->dominant_name = "k_enum_pw2";
      }
    } else if (gbin) {  // (gbin_ok: the only other blocks are unary factors of two-state variables)
      pgx::k_enum_unary_bin<<<grid_for(plan, mp, F), pgx::kThreads, 0, st>>>(mp, eb.dev, a.Es / 2, plan->d_edge_vs, lp, S,
                                                                            m_old, m_new, a);
      if ((rc = check_launch(plan, "k_enum_unary_bin"))) return rc;
    } else if (eb.variant == kUnary && !(plan->disabled_paths & PGX_PATH_ENUM_UNARY)) {
      pgx::k_enum_unary<<<grid_for(plan, mp, F), pgx::kThreads, 0, st>>>(mp, eb.dev, plan->d_edge_vs, lp, S, m_old, m_new, a);
      if ((rc = check_launch(plan, "k_enum_unary"))) return rc;
    } else if (eb.variant == kSmall || (eb.variant == kUnary && eb.dev.ns <= pgx::kSmallMaxNS)) {
      // complete pairwise tables with a 2 ... 4-state side, potentials shared by the batch: that side in registers
      if (eb.dense2 && eb.dev.ns <= 32 && lp.kind == 0 && mp.bx_log == 5 &&
          (eb.dev.ns - eb.n0 <= 4 || eb.n0 <= 4) && std::min(eb.n0, eb.dev.ns - eb.n0) >= 2 &&
          !(plan->disabled_paths & (PGX_PATH_ENUM_CONFIG_MAJOR | PGX_PATH_ENUM_DENSE_PAIR | PGX_PATH_ENUM_PAIR_FEW))) {
        const bool second = eb.dev.ns - eb.n0 <= 4;  // the few-state variable is the factor's second one
        const int few = second ? eb.dev.ns - eb.n0 : eb.n0, many = eb.dev.ns - few;
        const size_t smem = (size_t(many) * pgx::kThreads + size_t(pgx::kThreads / 32) * pgx::kDenseMaxConfigs) * sizeof(float);
        const dim3 grid = grid_for(plan, mp, F);
        // (at most 30 columns + the staging area = 38 KB: below the 48 KB that need no attribute)
#define PGX_PAIR_FEW(FEW, SECOND)                                                                                         \
  pgx::k_enum_pair_few<kSum, FEW, SECOND><<<grid, pgx::kThreads, smem, st>>>(mp, eb.dev, plan->d_edge_vs, lp, S, m_old,  \
                                                                             m_new, a)
        if (few == 2) { if (second) PGX_PAIR_FEW(2, true); else PGX_PAIR_FEW(2, false); }
        else if (few == 3) { if (second) PGX_PAIR_FEW(3, true); else PGX_PAIR_FEW(3, false); }
        else { if (second) PGX_PAIR_FEW(4, true); else PGX_PAIR_FEW(4, false); }
#undef PGX_PAIR_FEW
        if ((rc = check_launch(plan, "k_enum_pair_few"))) return rc;
        if (int(bi) == plan->dominant) plan->dominant_name = "k_enum_pair_few";
        if ((rc = prof_mark(plan, st, int(bi)))) return rc;
        continue;
      }
      // complete pairwise tables: nested loops, the first variable's running values in registers
      if (eb.dense2 && eb.dev.ns <= 32 && !(plan->disabled_paths & (PGX_PATH_ENUM_CONFIG_MAJOR | PGX_PATH_ENUM_DENSE_PAIR))) {
        // potentials shared by the batch + full sample tiles: staged per warp in shared memory
        const bool stage = lp.kind == 0 && mp.bx_log == 5;
        const size_t smem = (pgx::pair_dense_cols(eb.dev.ns, eb.dev.ns - eb.n0, kSum) * pgx::kThreads +
                             (stage ? size_t(pgx::kThreads / 32) * pgx::kDenseMaxConfigs : 0)) * sizeof(float);
        if (attr_needed(kSum ? kAttrEnumDenseSum : kAttrEnumDenseMax)) {
          const int most = int((size_t(3) * 32 * pgx::kThreads + size_t(pgx::kThreads / 32) * pgx::kDenseMaxConfigs) * sizeof(float));
          PGX_CUDA(cudaFuncSetAttribute(pgx::k_enum_pair_dense<kSum, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
          PGX_CUDA(cudaFuncSetAttribute(pgx::k_enum_pair_dense<kSum, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
        }
        if (stage)
          pgx::k_enum_pair_dense<kSum, true><<<grid_for(plan, mp, F), pgx::kThreads, smem, st>>>(mp, eb.dev, plan->d_edge_vs,
                                                                                                lp, S, m_old, m_new, a);
        else
          pgx::k_enum_pair_dense<kSum, false><<<grid_for(plan, mp, F), pgx::kThreads, smem, st>>>(mp, eb.dev, plan->d_edge_vs,
                                                                                                 lp, S, m_old, m_new, a);
        if ((rc = check_launch(plan, "k_enum_pair_dense"))) return rc;
        if (int(bi) == plan->dominant) plan->dominant_name = "k_enum_pair_dense";
        if ((rc = prof_mark(plan, st, int(bi)))) return rc;
        continue;
      }
      // configuration-major walk (k_enum_small_cm) while its three columns leave room for two CTAs per SM
      if (eb.dev.ns <= 32 && !(plan->disabled_paths & PGX_PATH_ENUM_CONFIG_MAJOR)) {
        const size_t smem = size_t(kSum ? 3 : 2) * eb.dev.ns * pgx::kThreads * sizeof(float);
        if (attr_needed(kSum ? kAttrEnumCmSum : kAttrEnumCmMax)) {
          const int most = int(size_t(3) * 32 * pgx::kThreads * sizeof(float));
          PGX_CUDA(cudaFuncSetAttribute(pgx::k_enum_small_cm<kSum, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
          PGX_CUDA(cudaFuncSetAttribute(pgx::k_enum_small_cm<kSum, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
          PGX_CUDA(cudaFuncSetAttribute(pgx::k_enum_small_cm<kSum, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
        }
#define PGX_ENUM_CM(ARITY)                                                                                              \
  pgx::k_enum_small_cm<kSum, ARITY><<<grid_for(plan, mp, F), pgx::kThreads, smem, st>>>(mp, eb.dev, plan->d_edge_vs, lp, \
                                                                                         S, m_old, m_new, a)
        if (eb.dev.arity == 2) PGX_ENUM_CM(2);
        else if (eb.dev.arity == 3) PGX_ENUM_CM(3);
        else PGX_ENUM_CM(0);
#undef PGX_ENUM_CM
        if ((rc = check_launch(plan, "k_enum_small_cm"))) return rc;
        if (int(bi) == plan->dominant) plan->dominant_name = "k_enum_small_cm";
        if ((rc = prof_mark(plan, st, int(bi)))) return rc;
        continue;
      }
      // per-thread columns of q / damped values in shared memory (kernels/enum.cuh)
      const size_t smem = size_t(2) * eb.dev.ns * pgx::kThreads * sizeof(float);
      if (attr_needed(kSum ? kAttrEnumSmallSum : kAttrEnumSmallMax))
        PGX_CUDA(cudaFuncSetAttribute(pgx::k_enum_small<kSum, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(size_t(2) * pgx::kSmallMaxNS * pgx::kThreads * sizeof(float))));
      pgx::k_enum_small<kSum, false, true><<<grid_for(plan, mp, F), pgx::kThreads, smem, st>>>(
          mp, eb.dev, plan->d_edge_vs, lp, S, m_old, m_new, a);
      if ((rc = check_launch(plan, "k_enum_small"))) return rc;
    } else {
      const size_t smem = size_t(2 * eb.dev.ns + 32) * sizeof(float);
      const int64_t units = F * mp.batch;
      const int grid = int(std::min<int64_t>(units, int64_t(plan->num_sms) * 8));
      if (attr_needed(kSum ? kAttrBigEnumSum : kAttrBigEnumMax))
        PGX_CUDA(cudaFuncSetAttribute(pgx::k_enum_big<kSum>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      if (!kSum && eb.dev.sorted0) {
        if (attr_needed(kAttrMaxProd))
          PGX_CUDA(cudaFuncSetAttribute(pgx::k_enum_big_maxprod, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        pgx::k_enum_big_maxprod<<<grid, pgx::kThreads, smem, st>>>(mp, eb.dev, plan->d_edge_vs, lp, S, m_old,
                                                                   m_new, a);
        if ((rc = check_launch(plan, "k_enum_big_maxprod"))) return rc;
        if (int(bi) == plan->dominant) plan->dominant_name = "k_enum_big_maxprod";
      } else {
        pgx::k_enum_big<kSum><<<grid, pgx::kThreads, smem, st>>>(mp, eb.dev, plan->d_edge_vs, lp, S, m_old,
                                                                 m_new, a);
        if ((rc = check_launch(plan, "k_enum_big"))) return rc;
      }
    }
    if ((rc = prof_mark(plan, st, int(bi)))) return rc;
  }
  int lid = -1;
  const cudaStream_t main_st = st;
  const bool orand_fused = lpull && lbin && plan->orand_fused_ok && !(plan->disabled_paths & PGX_PATH_ORAND_FUSED);
  if (orand_fused) {
    // one launch for both groups: CTA per (OR factor, sample tile)
    const size_t smem = pgx::orand_fused_smem(plan->orand.max_parents);
    if (attr_needed(kSum ? kAttrOrAndSum : kAttrOrAndMax)) {
#define PGX_FUSED_ATTR(PACK)                                                                                                     \
  PGX_CUDA(cudaFuncSetAttribute(pgx::k_or_and_fused<kSum, true, PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); \
  PGX_CUDA(cudaFuncSetAttribute(pgx::k_or_and_fused<kSum, false, PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)))
      PGX_FUSED_ATTR(1); PGX_FUSED_ATTR(2); PGX_FUSED_ATTR(4); PGX_FUSED_ATTR(8);
#undef PGX_FUSED_ATTR
    }
    // A short batch tail (<= 16 samples in the last tile) takes the packed variant: 2 / 4 / 8 OR
    // factors per CTA instead of a tile of mostly idle lanes - beside the full tiles, on the
    // auxiliary stream when there is one.
    const int tail_n = mp.batch & 31;
    const int pack = (tail_n == 0 || tail_n > 16 || (plan->disabled_paths & PGX_PATH_TAIL_SPLIT)) ? 1
                     : tail_n <= 4 ? 8 : tail_n <= 8 ? 4 : 2;
    const int full_tiles = pack > 1 ? mp.batch / 32 : int(mp.nbt);
    const dim3 grid(unsigned(plan->orand.num_or), unsigned(full_tiles));
    const dim3 grid_tail(unsigned((plan->orand.num_or + pack - 1) / pack), 1);
    const bool tail_side = pack > 1 && full_tiles > 0 && plan->aux != nullptr && !(plan->disabled_paths & PGX_PATH_AUX_STREAM);
    const cudaStream_t st_tail = tail_side ? plan->aux : st;
    const bool delta = a.deltas != nullptr;
    if ((rc = prof_mark(plan, st, plan->dominant))) return rc;
    if (pack > 1) {
      if (tail_side) {
        PGX_CUDA(cudaEventRecord(plan->ev_fork, st));
        PGX_CUDA(cudaStreamWaitEvent(plan->aux, plan->ev_fork, 0));
      }
#define PGX_FUSED_TAIL(PACK)                                                                                           \
  if (delta)                                                                                                           \
    pgx::k_or_and_fused<kSum, true, PACK><<<grid_tail, pgx::kFusedWarps * 32, smem, st_tail>>>(mp.batch, full_tiles,   \
                                                                                                plan->orand, ev, S,    \
                                                                                                m_old, m_new, a);      \
  else                                                                                                                 \
    pgx::k_or_and_fused<kSum, false, PACK><<<grid_tail, pgx::kFusedWarps * 32, smem, st_tail>>>(mp.batch, full_tiles,  \
                                                                                                 plan->orand, ev, S,   \
                                                                                                 m_old, m_new, a)
      if (pack == 8) { PGX_FUSED_TAIL(8); }
      else if (pack == 4) { PGX_FUSED_TAIL(4); }
      else { PGX_FUSED_TAIL(2); }
#undef PGX_FUSED_TAIL
      if ((rc = check_launch(plan, "k_or_and_fused(tail)"))) return rc;
      if (tail_side) PGX_CUDA(cudaEventRecord(plan->ev_join, plan->aux));
    }
    if (full_tiles > 0) {
      if (delta)
        pgx::k_or_and_fused<kSum, true, 1><<<grid, pgx::kFusedWarps * 32, smem, st>>>(mp.batch, 0, plan->orand, ev, S, m_old, m_new, a);
      else
        pgx::k_or_and_fused<kSum, false, 1><<<grid, pgx::kFusedWarps * 32, smem, st>>>(mp.batch, 0, plan->orand, ev, S, m_old, m_new, a);
    }
    if (full_tiles > 0 && (rc = check_launch(plan, "k_or_and_fused"))) return rc;
    if (tail_side) PGX_CUDA(cudaStreamWaitEvent(st, plan->ev_join, 0));
    if (plan->dominant < 0) {
      plan->dominant_name = "k_or_and_fused";
      plan->dominant_grid = int64_t(grid.x) * grid.y;
    }
    if ((rc = prof_mark(plan, st, plan->dominant))) return rc;
  }
  for (LogicalPlan* lg : {&plan->or_f, &plan->and_f}) {
    const int id = lid--;
    if (lg->dev.num_factors == 0 || orand_fused) continue;
    const cudaStream_t st = (lpull && aux != nullptr && id == plan->aux_group) ? aux : main_st;  // NOLINT
    if ((rc = prof_mark(plan, st, id))) return rc;
    if (lpull) {
      const bool delta = a.deltas != nullptr;
      const int64_t F = lg->dev.num_factors;
      if (lg->max_parents <= pgx::kRegParents) {
        const int un = lg->dev.uniform;  // 0: ragged
        const int units = (un == 1 || un == 2) ? 2 : 1;
        const int64_t warps = (F + units - 1) / units;
        const dim3 grid(unsigned(std::min<int64_t>((warps + 7) / 8, std::max<int64_t>(1, int64_t(plan->num_sms) * 16 / mp.nbt))),
                        unsigned(mp.nbt));
#define PGX_LAUNCH_SMALL2(NP, U, UNI, BIN)                                                                     \
  do {                                                                                                          \
    if (delta)                                                                                                  \
      pgx::k_logical_pull_small<kSum, true, NP, U, UNI, BIN><<<grid, pgx::kThreads, 0, st>>>(mp.batch, lg->pull, ev, S, \
                                                                                             m_old, m_new, a); \
    else                                                                                                        \
      pgx::k_logical_pull_small<kSum, false, NP, U, UNI, BIN><<<grid, pgx::kThreads, 0, st>>>(mp.batch, lg->pull, ev, S, \
                                                                                              m_old, m_new, a); \
  } while (0)
#define PGX_LAUNCH_SMALL(NP, U, UNI)                  \
  do {                                                \
    if (lbin) PGX_LAUNCH_SMALL2(NP, U, UNI, true);    \
    else PGX_LAUNCH_SMALL2(NP, U, UNI, false);        \
  } while (0)
#define PGX_LAUNCH_STAGED2(NP, U, BIN)                                                                          \
  do {                                                                                                          \
    if (delta)                                                                                                  \
      pgx::k_logical_pull_small_staged<kSum, true, NP, U, BIN><<<sgrid, pgx::kThreads, 0, st>>>(mp.batch, lg->pull, ev, \
                                                                                             S, m_old, m_new, a); \
    else                                                                                                        \
      pgx::k_logical_pull_small_staged<kSum, false, NP, U, BIN><<<sgrid, pgx::kThreads, 0, st>>>(mp.batch, lg->pull, ev, \
                                                                                              S, m_old, m_new, a); \
  } while (0)
#define PGX_LAUNCH_STAGED(NP, U)                    \
  do {                                              \
    if (lbin) PGX_LAUNCH_STAGED2(NP, U, true);      \
    else PGX_LAUNCH_STAGED2(NP, U, false);          \
  } while (0)
        // uniform groups: wiring staged in shared memory, a contiguous factor range per CTA
        const bool staged = un >= 1 && un <= 4 && !(plan->disabled_paths & PGX_PATH_STAGED_WIRING);
        const dim3 sgrid(unsigned((F + pgx::kPullChunk - 1) / pgx::kPullChunk), unsigned(mp.nbt));
        if (staged && un == 1) PGX_LAUNCH_STAGED(1, 2);
        else if (staged && un == 2) PGX_LAUNCH_STAGED(2, 2);
        else if (staged && un == 3) PGX_LAUNCH_STAGED(3, 1);
        else if (staged && un == 4) PGX_LAUNCH_STAGED(4, 1);
        else if (un == 1) PGX_LAUNCH_SMALL(1, 2, true);
        else if (un == 2) PGX_LAUNCH_SMALL(2, 2, true);
        else if (un == 3) PGX_LAUNCH_SMALL(3, 1, true);
        else if (un == 4) PGX_LAUNCH_SMALL(4, 1, true);
        else PGX_LAUNCH_SMALL(4, 1, false);
#undef PGX_LAUNCH_STAGED
#undef PGX_LAUNCH_STAGED2
#undef PGX_LAUNCH_SMALL
#undef PGX_LAUNCH_SMALL2
        if ((rc = check_launch(plan, "k_logical_pull_small"))) return rc;
        if (id == plan->dominant) {
          plan->dominant_name = "k_logical_pull_small";
          plan->dominant_grid = staged ? int64_t(sgrid.x) * sgrid.y : int64_t(grid.x) * grid.y;
        }
      } else {
        const dim3 grid0(unsigned((F + 3) / 4), unsigned(mp.nbt));
        // two launches: serial accumulation per factor, then a parent-parallel emit
        const dim3 grid1(unsigned(std::min<int64_t>(F, 1 << 20)), unsigned(mp.nbt));
        const int64_t warps2 = (lg->num_parents + pgx::kEmitUnits - 1) / pgx::kEmitUnits;
        const dim3 grid2(unsigned(std::min<int64_t>((warps2 + 7) / 8, std::max<int64_t>(1, int64_t(plan->num_sms) * 16 / mp.nbt))),
                         unsigned(mp.nbt));
        const bool split = !(plan->disabled_paths & PGX_PATH_WIDE_SPLIT);
        const bool stage_emit = !(plan->disabled_paths & PGX_PATH_STAGED_WIRING);
        const dim3 grid2s(unsigned((lg->num_parents + pgx::kEmitChunk - 1) / pgx::kEmitChunk), unsigned(mp.nbt));
#define PGX_LAUNCH_WIDE(DELTA, BIN)                                                                              \
  do {                                                                                                          \
    if (!split) {                                                                                               \
      pgx::k_logical_pull_wide<kSum, DELTA, BIN><<<grid0, 128, 0, st>>>(mp.batch, lg->pull, ev, S, m_old, m_new, a); \
    } else {                                                                                                    \
      pgx::k_logical_wide_reduce<kSum, DELTA, BIN><<<grid1, 32, 0, st>>>(mp.batch, lg->pull, ev, S, m_old, m_new, \
                                                                         plan->ws.agg, a);                     \
      if (stage_emit)                                                                                           \
        pgx::k_logical_wide_emit_staged<kSum, DELTA, BIN><<<grid2s, pgx::kThreads, 0, st>>>(                    \
            mp.batch, lg->pull, lg->d_parent_factor, lg->num_parents, ev, S, m_old, m_new, plan->ws.agg, a);    \
      else                                                                                                      \
        pgx::k_logical_wide_emit<kSum, DELTA, BIN><<<grid2, pgx::kThreads, 0, st>>>(                            \
            mp.batch, lg->pull, lg->d_parent_factor, lg->num_parents, ev, S, m_old, m_new, plan->ws.agg, a);    \
    }                                                                                                           \
  } while (0)
        if (delta) {
          if (lbin) PGX_LAUNCH_WIDE(true, true); else PGX_LAUNCH_WIDE(true, false);
        } else {
          if (lbin) PGX_LAUNCH_WIDE(false, true); else PGX_LAUNCH_WIDE(false, false);
        }
#undef PGX_LAUNCH_WIDE
        if (split) ++plan->launches;
        if ((rc = check_launch(plan, "k_logical_pull_wide"))) return rc;
        if (id == plan->dominant) plan->dominant_name = "k_logical_pull_wide";
      }
      if ((rc = prof_mark(plan, st, id))) return rc;
      continue;
    }
    if (lg->dev.uniform > 0 && lg->dev.uniform <= pgx::kRegParents)
      pgx::k_logical_uniform<kSum><<<grid_for(plan, mp, lg->dev.num_factors), pgx::kThreads, 0, st>>>(
          mp, lg->dev, S, m_old, m_new, a);
    else if (lg->max_parents <= pgx::kRegParents)
      pgx::k_logical<kSum, true><<<grid_for(plan, mp, lg->dev.num_factors), pgx::kThreads, 0, st>>>(
          mp, lg->dev, S, m_old, m_new, a);
    else
      pgx::k_logical<kSum, false><<<grid_for(plan, mp, lg->dev.num_factors), pgx::kThreads, 0, st>>>(
          mp, lg->dev, S, m_old, m_new, a);
    if ((rc = check_launch(plan, "k_logical"))) return rc;
    if ((rc = prof_mark(plan, st, id))) return rc;
  }
  if (plan->pool_f.dev.num_factors > 0) {
    if ((rc = prof_mark(plan, st, -3))) return rc;
    pgx::k_pool<kSum><<<grid_for(plan, mp, plan->pool_f.dev.num_factors), pgx::kThreads, 0, st>>>(
        mp, plan->pool_f.dev, S, m_old, m_new, a);
    if ((rc = check_launch(plan, "k_pool"))) return rc;
    if ((rc = prof_mark(plan, st, -3))) return rc;
  }
  return PGX_OK;
}

// Makes the plan's device current for the duration of an entry point and restores the caller's
// device on exit (a caller that holds plans on several GPUs keeps its own current device).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  int enter(const pgx_plan* plan) {
    if (cudaGetDevice(&prev) != cudaSuccess) return fail(PGX_ERR_NO_DEVICE, "no CUDA device");
    if (prev != plan->device) {
      PGX_CUDA(cudaSetDevice(plan->device));
      switched = true;
    }
    return PGX_OK;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

// A lattice launch on binary-difference storage: variant (max / sum-product, with / without
// deltas), persistent grid of one CTA per SM.
// Tile shape / ring depth / consumer threads / CTAs per SM of k_lattice_bin.  Measured on Ising 8192^2
// (ms per iteration, profiles/r02_d_lattice_variants.txt): 4x128 tiles, 2 stages, 8 consumer warps, THREE
// independent CTAs per SM 0.830; 4x256x3 stages, 16 warps, one CTA 0.871; 6x256x2, 24 warps, one CTA 0.876;
// 4x128x2, 12 warps x 3 CTAs 0.872; 4x64x3, 8 warps x 4 CTAs 0.939; 2x256x3, 16 warps x 2 CTAs 0.980 - several
// small pipelines whose phases drift apart hide the barrier / mbarrier waits of one another.
using LbCfgDefault = pgx::LbCfg<4, 128, 2, 256, 3>;
using LbCfgV1 = pgx::LbCfg<4, 256, 3, 512, 1>;  // PGX_LB_VARIANT=1 (A/B)
using LbFn = void (*)(pgx::LatticeBinArgs, const float*, const float*, const float4*, float4*, pgx::RunArgs);

template <class Cfg>
int launch_lattice_bin_cfg(uint32_t* attr_done, int attr_bit, int num_sms, cudaStream_t st, const pgx::LatticeBinArgs& g,
                           const float* ev, const float* lp, const float* c_old, float* c_new, const pgx::RunArgs& a,
                           bool sum_product, bool want_delta) {
  static const LbFn fns[4] = {pgx::k_lattice_bin<false, false, Cfg>, pgx::k_lattice_bin<false, true, Cfg>,
                              pgx::k_lattice_bin<true, false, Cfg>, pgx::k_lattice_bin<true, true, Cfg>};
  if (!((*attr_done >> attr_bit) & 1u)) {
    *attr_done |= 1u << attr_bit;
    for (int v = 0; v < 4; ++v)
      PGX_CUDA(cudaFuncSetAttribute(fns[v], cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::smem_bytes())));
  }
  const int tiles_x = (g.N + Cfg::TC - 1) / Cfg::TC;
  int64_t tiles = 0;
  for (int s = 0; s < 2; ++s) tiles += int64_t(tiles_x) * std::max(0, (g.seg_end[s] - g.seg_begin[s] + Cfg::TR - 1) / Cfg::TR);
  if (tiles == 0) return PGX_OK;
  PGX_CHECK(tiles < (int64_t(1) << 31), "lattice too large for 32-bit tile indices");
  const unsigned grid = unsigned(std::min<int64_t>(tiles, int64_t(num_sms) * Cfg::kCtasPerSm));
  fns[(sum_product ? 2 : 0) + (want_delta ? 1 : 0)]<<<grid, Cfg::kThreads, Cfg::smem_bytes(), st>>>(
      g, ev, lp, reinterpret_cast<const float4*>(c_old), reinterpret_cast<float4*>(c_new), a);
  return PGX_OK;
}

int launch_lattice_bin(pgx_plan* plan_or_null, uint32_t* attr_done, int num_sms, cudaStream_t st,
                       const pgx::LatticeBinArgs& g, const float* ev, const float* lp, const float* c_old, float* c_new,
                       const pgx::RunArgs& a, bool sum_product, bool want_delta) {
  static const int variant = [] {
    const char* v = getenv("PGX_LB_VARIANT");
    return v ? atoi(v) : 0;
  }();
  int rc;
#define PGX_LB_CASE(V, CFG)                                                                                           \
  case V:                                                                                                             \
    rc = launch_lattice_bin_cfg<CFG>(attr_done, kAttrLatticeBin + V, num_sms, st, g, ev, lp, c_old, c_new, a, sum_product, \
                                     want_delta);                                                                     \
    break
  switch (variant) {
    PGX_LB_CASE(1, LbCfgV1);
    default:
      rc = launch_lattice_bin_cfg<LbCfgDefault>(attr_done, kAttrLatticeBin, num_sms, st, g, ev, lp, c_old, c_new, a,
                                                sum_product, want_delta);
  }
#undef PGX_LB_CASE
  if (rc) return rc;
  if (plan_or_null) return check_launch(plan_or_null, "k_lattice_bin");
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(PGX_ERR_CUDA, "launch of k_lattice_bin failed: %s", cudaGetErrorString(err));
  return PGX_OK;
}

pgx::RunArgs make_run_args(float damping, float temperature, float* deltas, int32_t num_iters, int64_t Es, int64_t Vs) {
  pgx::RunArgs a{};
  a.d = damping;
  a.one_minus_d = 1.0f - damping;
  a.T = temperature;
  if (temperature > 0.f) {
    a.c_exp = float(1.4426950408889634 / double(temperature));
    a.c_log = float(0.6931471805599453 * double(temperature));
  }
  a.deltas = deltas;
  a.delta_stride = num_iters;
  a.Es = Es;
  a.Vs = Vs;
  return a;
}

// pgx_bp_run for one sample of a large 2-D lattice: normalise + compress the input messages
// once, num_iters launches of k_lattice_bin ping-ponging between the two compressed buffers,
// expand once into the caller's buffer.
int run_lattice_bin(pgx_plan* plan, cudaStream_t st, const float* log_potentials, const float* evidence,
                    const float* ftov_in, float* ftov_out, float* deltas, int32_t num_iters, float damping,
                    float temperature) {
  int rc;
  Workspace& ws = plan->ws;
  const pgx::LatticeDev& lat = plan->lattice;
  const int64_t cells = int64_t(lat.R) * lat.N;
  const unsigned ew_grid = unsigned(plan->num_sms * 8);
  if (ftov_in == nullptr) {
    PGX_CUDA(cudaMemsetAsync(ws.cA, 0, size_t(cells) * sizeof(float4), st));  // NC(0) = 0
  } else {
    pgx::k_lattice_compress<<<ew_grid, pgx::kThreads, 0, st>>>(reinterpret_cast<const float4*>(ftov_in + lat.first_msg),
                                                               reinterpret_cast<float4*>(ws.cA), cells);
    if ((rc = check_launch(plan, "k_lattice_compress"))) return rc;
  }
  if (deltas) PGX_CUDA(cudaMemsetAsync(deltas, 0, size_t(num_iters) * sizeof(float), st));
  pgx::RunArgs a = make_run_args(damping, temperature, deltas, num_iters, plan->num_edge_states, plan->num_var_states);
  pgx::LatticeBinArgs g{};
  g.R = lat.R;
  g.N = lat.N;
  g.torus = lat.torus;
  g.seg_begin[0] = 0;
  g.seg_end[0] = lat.R;
  g.up_add = nullptr;
  g.ghost_ev = nullptr;
  g.ghost_terms = nullptr;
  plan->dominant_name = "k_lattice_bin";
  plan->dominant_grid = std::min<int64_t>(int64_t((lat.N + LbCfgDefault::TC - 1) / LbCfgDefault::TC) *
                                              ((lat.R + LbCfgDefault::TR - 1) / LbCfgDefault::TR),
                                          int64_t(plan->num_sms) * LbCfgDefault::kCtasPerSm);
  const float* cur = ws.cA;
  float* nxt = ws.cB;
  for (int it = 0; it < num_iters; ++it) {
    a.delta_off = it;
    if ((rc = prof_mark(plan, st, 0))) return rc;
    if ((rc = launch_lattice_bin(plan, &plan->attr_done, plan->num_sms, st, g, evidence, log_potentials + lat.first_pot, cur,
                                 nxt, a, temperature > 0.f, deltas != nullptr)))
      return rc;
    if ((rc = prof_mark(plan, st, 0))) return rc;
    const float* t = cur;
    cur = nxt;
    nxt = const_cast<float*>(t);
  }
  pgx::k_lattice_expand<<<ew_grid, pgx::kThreads, 0, st>>>(reinterpret_cast<const float4*>(cur),
                                                           reinterpret_cast<float4*>(ftov_out + lat.first_msg), cells);
  return check_launch(plan, "k_lattice_expand");
}

}  // namespace

extern "C" {

const char* pgx_last_error(void) { return g_last_error.c_str(); }

const char* pgx_build_info(void) {
  return "pgx 1 sm_100a cuda-" PGX_STR(CUDART_VERSION) " fmad=off";
}

int pgx_plan_create(const pgx_graph_desc* desc, pgx_plan** out_plan) {
  if (out_plan == nullptr || desc == nullptr) return fail(PGX_ERR_INVALID, "null argument");
  *out_plan = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(PGX_ERR_NO_DEVICE,
                "no CUDA device visible; pgx has no CPU fallback (the oracle under oracle/ is test-only)");
  pgx_plan* plan = new pgx_plan();
  auto bail = [&](int rc) {
    pgx_plan_destroy(plan);
    return rc;
  };
#define PGX_TRY(expr)                 \
  do {                                \
    int rc__ = (expr);                \
    if (rc__ != PGX_OK) return bail(rc__); \
  } while (0)
#define PGX_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return bail(fail(PGX_ERR_INVALID, __VA_ARGS__)); \
  } while (0)

  if (cudaGetDevice(&plan->device) != cudaSuccess) return bail(fail(PGX_ERR_CUDA, "cudaGetDevice failed"));
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, plan->device) != cudaSuccess)
    return bail(fail(PGX_ERR_CUDA, "cudaGetDeviceProperties failed"));
  plan->num_sms = prop.multiProcessorCount;

  PGX_REQUIRE(desc->num_vars >= 0 && desc->num_edges >= 0 && desc->num_potentials >= 0, "negative size");
  PGX_REQUIRE(desc->num_vars < INT32_MAX && desc->num_edges < INT32_MAX, "graph too large for int32 indices");
  PGX_REQUIRE(desc->num_vars == 0 || desc->var_num_states, "var_num_states is null");
  PGX_REQUIRE(desc->num_edges == 0 || (desc->edge_var_start && desc->edge_num_states), "edge table is null");
  plan->num_vars = desc->num_vars;
  plan->num_edges = desc->num_edges;
  plan->num_potentials = desc->num_potentials;

  // variables
  std::vector<int32_t> var_first_state(plan->num_vars + 1, 0);
  for (int64_t v = 0; v < plan->num_vars; ++v) {
    const int32_t ns = desc->var_num_states[v];
    PGX_REQUIRE(ns >= 0, "variable %lld has %d states", (long long)v, ns);
    const int64_t next = int64_t(var_first_state[v]) + ns;
    PGX_REQUIRE(next < INT32_MAX, "more than 2^31 variable states");
    var_first_state[v + 1] = int32_t(next);
    plan->max_var_states = std::max<int64_t>(plan->max_var_states, ns);
  }
  plan->num_var_states = var_first_state[plan->num_vars];
  std::vector<int32_t> vs_var(plan->num_var_states);
  for (int64_t v = 0; v < plan->num_vars; ++v)
    for (int32_t s = var_first_state[v]; s < var_first_state[v + 1]; ++s) vs_var[s] = int32_t(v);

  // edges, message offsets, variable -> incident edges CSR (ascending message index)
  std::vector<int64_t> edge_msg_start(plan->num_edges + 1, 0);
  std::vector<int32_t> edge_vs(desc->edge_var_start, desc->edge_var_start + plan->num_edges);
  std::vector<int32_t> edge_ns(desc->edge_num_states, desc->edge_num_states + plan->num_edges);
  std::vector<int64_t> var_ptr(plan->num_vars + 1, 0);
  for (int64_t e = 0; e < plan->num_edges; ++e) {
    const int32_t vs = edge_vs[e];
    PGX_REQUIRE(vs >= 0 && vs < plan->num_var_states, "edge %lld: var-state %d out of range", (long long)e, vs);
    const int32_t var = vs_var[vs];
    PGX_REQUIRE(var_first_state[var] == vs && desc->var_num_states[var] == edge_ns[e],
                "edge %lld does not span exactly the states of variable %d", (long long)e, var);
    edge_msg_start[e + 1] = edge_msg_start[e] + edge_ns[e];
    ++var_ptr[var + 1];
  }
  plan->num_edge_states = edge_msg_start[plan->num_edges];
  PGX_REQUIRE(plan->num_edge_states < INT32_MAX, "more than 2^31 edge states");
  for (int64_t v = 0; v < plan->num_vars; ++v) var_ptr[v + 1] += var_ptr[v];
  std::vector<int64_t> var_edge_msg(plan->num_edges);
  {
    std::vector<int64_t> cursor(var_ptr.begin(), var_ptr.end() - 1);
    for (int64_t e = 0; e < plan->num_edges; ++e) var_edge_msg[cursor[vs_var[edge_vs[e]]]++] = edge_msg_start[e];
  }
  PGX_TRY(upload(edge_vs, &plan->d_edge_vs, &plan->device_bytes));
  PGX_TRY(upload(narrow(edge_msg_start), &plan->d_edge_msg_start, &plan->device_bytes));
  {
    std::vector<int2> vs_csr(plan->num_var_states);
    for (int64_t s = 0; s < plan->num_var_states; ++s) {
      const int32_t var = vs_var[s];
      const int64_t deg = var_ptr[var + 1] - var_ptr[var], st = s - var_first_state[var];
      if (st >= (1 << pgx::kVsStateBits) || deg >= (int64_t(1) << (31 - pgx::kVsStateBits)))
        return bail(fail(PGX_ERR_UNSUPPORTED, "variable %d: %lld states / %lld incident edges exceed the packed CSR row format",
                         var, (long long)desc->var_num_states[var], (long long)deg));
      vs_csr[s] = make_int2(int(var_ptr[var]), int((deg << pgx::kVsStateBits) | st));
    }
    PGX_TRY(upload(vs_csr, &plan->d_vs_csr, &plan->device_bytes));
  }
  PGX_TRY(upload(vs_var, &plan->d_vs_var, &plan->device_bytes));
  PGX_TRY(upload(var_first_state, &plan->d_var_first_state, &plan->device_bytes));
  PGX_TRY(upload(narrow(var_ptr), &plan->d_var_ptr, &plan->device_bytes));
  PGX_TRY(upload(narrow(var_edge_msg), &plan->d_var_edge_msg, &plan->device_bytes));

  // factor types
  const int64_t desc_num_parents[3] = {desc->or_factors.num_parents, desc->and_factors.num_parents,
                                       desc->pool_factors.num_parents};
  std::vector<uint8_t> edge_covered(plan->num_edges, 0);
  PGX_REQUIRE(desc->num_enum_blocks == 0 || desc->enum_blocks, "enum_blocks is null");
  {
    // Blocks with identical (arity, configuration table, state counts) share one set of device
    // tables and one launch, except full pairwise-binary blocks (those keep their arithmetic
    // indexing for the dense-grid / pull kernels).
    std::vector<std::vector<int>> groups;
    std::unordered_map<uint64_t, std::vector<int>> by_hash;  // hash -> group indices
    for (int i = 0; i < desc->num_enum_blocks; ++i) {
      const pgx_enum_block& b = desc->enum_blocks[i];
      PGX_TRY(check_enum_block(plan, b, i, edge_msg_start, edge_ns, edge_covered));
      const size_t nbytes = size_t(b.num_configs) * b.arity * sizeof(int32_t);
      uint64_t h = 1469598103934665603ull;
      auto mix = [&h](const void* p, size_t n) {
        const unsigned char* c = static_cast<const unsigned char*>(p);
        for (size_t k = 0; k < n; ++k) h = (h ^ c[k]) * 1099511628211ull;
      };
      mix(&b.arity, sizeof(b.arity));
      mix(&b.num_configs, sizeof(b.num_configs));
      for (int a = 0; a < b.arity; ++a) mix(&edge_ns[b.first_edge + a], sizeof(int32_t));
      mix(b.configs, nbytes);
      const bool pw2 = b.arity == 2 && b.num_configs == 4 && edge_ns[b.first_edge] == 2 &&
                       edge_ns[b.first_edge + 1] == 2;
      int found = -1;
      if (!pw2)
        for (int gi : by_hash[h]) {
          const pgx_enum_block& r = desc->enum_blocks[groups[gi][0]];
          bool same = r.arity == b.arity && r.num_configs == b.num_configs &&
                      std::memcmp(r.configs, b.configs, nbytes) == 0;
          for (int a = 0; a < b.arity && same; ++a) same = edge_ns[r.first_edge + a] == edge_ns[b.first_edge + a];
          if (same) { found = gi; break; }
        }
      if (found < 0) {
        if (!pw2) by_hash[h].push_back(int(groups.size()));
        groups.push_back({i});
      } else {
        groups[found].push_back(i);
      }
    }
    plan->enum_blocks.resize(groups.size());
    for (size_t gi = 0; gi < groups.size(); ++gi)
      PGX_TRY(build_enum_block(plan, desc->enum_blocks, groups[gi], edge_msg_start, edge_ns,
                               &plan->enum_blocks[gi]));
  }
  {  // merged max-product launch: group table + work units sorted by configuration count
    std::vector<pgx::BigMaxGroup> groups;
    struct Unit { int g; int32_t f; int32_t cost; };
    std::vector<Unit> units;
    const int nwarp = pgx::kThreads / 32;
    for (EnumBlockPlan& eb : plan->enum_blocks) {
      if (eb.bigmax < 0) continue;
      const size_t smem = (size_t(2) * eb.dev.ns + 32 + size_t(nwarp) * (eb.dev.ns - eb.n0 + 32) + 32) * sizeof(float);
      if (smem > 200 * 1024 || eb.dev.num_factors >= INT32_MAX) { eb.bigmax = -1; continue; }
      eb.bigmax = int(groups.size());
      plan->bigmax_smem = std::max(plan->bigmax_smem, smem);
      plan->bigsum_smem = std::max(plan->bigsum_smem, smem + size_t(nwarp) * (eb.dev.ns - eb.n0 + 32) * sizeof(float));
      plan->bigmax_es += eb.dev.num_factors * eb.dev.ns;
      pgx::BigMaxGroup g;
      g.blk = eb.dev;
      g.rounds = eb.d_rounds;
      g.round_ptr = eb.d_round_ptr;
      g.num_groups = eb.num_groups;
      g.rounds_b = eb.d_rounds_b;
      g.num_rounds = eb.num_rounds;
      g.lane_state = eb.d_lane_state;
      g.perm_base = plan->bigmax_perm_floats;
      plan->bigmax_perm_floats += int64_t(eb.num_rounds) * 32 * eb.dev.num_factors;
      groups.push_back(g);
      for (int64_t f = 0; f < eb.dev.num_factors; ++f) units.push_back({eb.bigmax, int32_t(f), eb.dev.num_configs});
    }
    if (!groups.empty()) {
      std::stable_sort(units.begin(), units.end(), [](const Unit& x, const Unit& y) { return x.cost > y.cost; });
      std::vector<int2> u2(units.size());
      for (size_t i = 0; i < units.size(); ++i) u2[i] = make_int2(units[i].g, units[i].f);
      PGX_TRY(upload(groups, &plan->d_bigmax_groups, &plan->device_bytes));
      PGX_TRY(upload(u2, &plan->d_bigmax_units, &plan->device_bytes));
      plan->bigmax_units = int64_t(units.size());
      if (cudaMalloc(reinterpret_cast<void**>(&plan->d_bigmax_counter), 2 * sizeof(unsigned int)) != cudaSuccess)
        return bail(fail(PGX_ERR_CUDA, "cudaMalloc failed"));
    }
  }
  PGX_TRY(build_logical(plan, desc->or_factors, edge_msg_start, edge_vs, edge_ns, edge_covered, "OR factors",
                        &plan->or_f));
  PGX_TRY(build_logical(plan, desc->and_factors, edge_msg_start, edge_vs, edge_ns, edge_covered,
                        "AND factors", &plan->and_f));
  PGX_REQUIRE(desc->pool_factors.num_factors == 0 || desc->pool_factors.edge_states_offset == 1,
              "Pool factors: edge_states_offset must be +1");
  PGX_TRY(build_logical(plan, desc->pool_factors, edge_msg_start, edge_vs, edge_ns, edge_covered,
                        "Pool factors", &plan->pool_f));
  for (int64_t e = 0; e < plan->num_edges; ++e)
    PGX_REQUIRE(edge_covered[e], "edge %lld belongs to no factor description", (long long)e);
  if (plan->enum_blocks.empty() && plan->pool_f.dev.num_factors == 0 &&
      plan->or_f.dev.num_factors + plan->and_f.dev.num_factors > 0) {
    // pull wiring: per edge the message index of the variable's only other edge (-1: none,
    // -2: more than two edges, sum read from S)
    for (LogicalPlan* lg : {&plan->or_f, &plan->and_f}) {
      if (lg->dev.num_factors == 0) continue;
      const int rel = lg->dev.off > 0 ? 0 : 1;
      auto pack = [&](int32_t msg, int32_t vs) {
        const int32_t var = vs_var[vs];
        const int64_t k0 = var_ptr[var], deg = var_ptr[var + 1] - k0;
        pgx::EdgeW w{msg, vs, -2, 0};
        if (deg == 1) w.other = -1;
        else if (deg == 2) {
          const int64_t own_lo = int64_t(msg) - rel;
          const int64_t other_lo = var_edge_msg[k0] == own_lo ? var_edge_msg[k0 + 1] : var_edge_msg[k0];
          w.other = int32_t(other_lo + rel);
        }
        return w;
      };
      std::vector<pgx::EdgeW> pw(lg->h_pmsg.size()), cw(lg->h_cmsg.size());
      for (size_t i = 0; i < pw.size(); ++i) pw[i] = pack(lg->h_pmsg[i], lg->h_pvs[i]);
      for (size_t f = 0; f < cw.size(); ++f) cw[f] = pack(lg->h_cmsg[f], lg->h_cvs[f]);
      for (const pgx::EdgeW& e : pw) lg->needs_s = lg->needs_s || e.other == -2;
      for (const pgx::EdgeW& e : cw) lg->needs_s = lg->needs_s || e.other == -2;
      PGX_TRY(upload(pw, &lg->d_parents_w, &plan->device_bytes));
      PGX_TRY(upload(cw, &lg->d_children_w, &plan->device_bytes));
      lg->h_pw = pw;
      lg->h_cw = cw;
      {
        std::vector<int32_t> pf(pw.size());
        for (int64_t f = 0; f < lg->dev.num_factors; ++f)
          for (int32_t i = lg->h_ptr[f]; i < lg->h_ptr[f + 1]; ++i) pf[i] = int32_t(f);
        PGX_TRY(upload(pf, &lg->d_parent_factor, &plan->device_bytes));
        lg->num_parents = int64_t(pw.size());
      }
      lg->pull.num_factors = lg->dev.num_factors;
      lg->pull.parent_ptr = lg->d_parent_ptr;
      lg->pull.parents = lg->d_parents_w;
      lg->pull.children = lg->d_children_w;
      lg->pull.off = lg->dev.off;
      lg->pull.uniform = lg->dev.uniform;
    }
    std::vector<int32_t> hi;
    for (int64_t v = 0; v < plan->num_vars; ++v)
      if (var_ptr[v + 1] - var_ptr[v] > 2)
        for (int32_t sidx = var_first_state[v]; sidx < var_first_state[v + 1]; ++sidx) hi.push_back(sidx);
    // longest rows first: their serial chains start at once, the short ones fill in behind
    std::stable_sort(hi.begin(), hi.end(), [&](int32_t x, int32_t y) {
      const int32_t vx = vs_var[x], vy = vs_var[y];
      return var_ptr[vx + 1] - var_ptr[vx] > var_ptr[vy + 1] - var_ptr[vy];
    });
    PGX_TRY(upload(hi, &plan->d_hi_list, &plan->device_bytes));
    plan->hi_len = int64_t(hi.size());
    if (plan->num_edge_states == 2 * plan->num_edges)  // binary variables: list entries come in (state 0, state 1) pairs
      for (size_t k = 0; k + 1 < hi.size(); k += 2) {
        const int32_t var = vs_var[hi[k]];
        if (var_ptr[var + 1] - var_ptr[var] < pgx::kVsBigDegree) break;
        ++plan->hi_big;
      }
    plan->logical_pull_ok = true;
    {
      // Pairing for k_or_and_fused: every OR parent edge belongs to a variable with exactly two
      // edges whose other edge is the child edge of a two-parent AND factor (each AND factor used
      // once, all of them used), the OR edge has the smaller message index, the AND factors'
      // parents take their sums from S, and every edge has two states.
      const LogicalPlan& orp = plan->or_f;
      const LogicalPlan& andp = plan->and_f;
      bool ok = orp.dev.num_factors > 0 && andp.dev.num_factors > 0 && andp.dev.uniform == 2 &&
                int64_t(orp.h_pw.size()) == andp.dev.num_factors && plan->num_edge_states == 2 * plan->num_edges &&
                orp.max_parents <= 800;
      std::vector<pgx::FusedW> fw;
      if (ok) {
        std::unordered_map<int32_t, int32_t> and_of_child;  // state-0 message index of the AND child edge -> factor
        and_of_child.reserve(andp.h_cw.size() * 2);
        for (size_t f = 0; f < andp.h_cw.size(); ++f) and_of_child[andp.h_cw[f].msg - 1] = int32_t(f);
        std::vector<uint8_t> used(andp.h_cw.size(), 0);
        fw.resize(orp.h_pw.size());
        for (size_t i = 0; i < orp.h_pw.size() && ok; ++i) {
          const pgx::EdgeW& e = orp.h_pw[i];  // msg / other: state-0 message indices
          auto it = e.other >= 0 ? and_of_child.find(e.other) : and_of_child.end();
          ok = it != and_of_child.end() && !used[it->second] && e.msg < e.other;
          if (!ok) break;
          const int32_t f = it->second;
          used[f] = 1;
          const pgx::EdgeW &ps = andp.h_pw[2 * size_t(f)], &pw2 = andp.h_pw[2 * size_t(f) + 1];
          ok = ps.other == -2 && pw2.other == -2 && andp.h_cw[f].vs - 1 == e.vs;
          fw[i] = pgx::FusedW{e.msg >> 1, andp.h_cw[f].msg >> 1, e.vs, ps.msg >> 1, pw2.msg >> 1, ps.vs - 1, pw2.vs - 1, f};
        }
      }
      if (ok) {
        PGX_TRY(upload(fw, &plan->d_fused_w, &plan->device_bytes));
        plan->orand.num_or = orp.dev.num_factors;
        plan->orand.parent_ptr = orp.d_parent_ptr;
        plan->orand.w = plan->d_fused_w;
        plan->orand.or_children = orp.d_children_w;
        plan->orand.max_parents = int32_t(orp.max_parents);
        plan->orand_fused_ok = true;
      }
    }
    if (plan->or_f.dev.num_factors > 0 && plan->and_f.dev.num_factors > 0) {
      const int64_t es_or = plan->or_f.dev.num_factors + desc->or_factors.num_parents;
      const int64_t es_and = plan->and_f.dev.num_factors + desc->and_factors.num_parents;
      plan->aux_group = es_or <= es_and ? -1 : -2;
      plan->aux_needs_s = es_or <= es_and ? plan->or_f.needs_s : plan->and_f.needs_s;
      // highest priority: the group on the auxiliary stream is the one with long serial chains
      // (few, long-lived warps), it should get free SM slots before the wide grid beside it
      int prio_lo = 0, prio_hi = 0;
      cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
      if (cudaStreamCreateWithPriority(&plan->aux, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
          cudaEventCreateWithFlags(&plan->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&plan->ev_join, cudaEventDisableTiming) != cudaSuccess)
        return bail(fail(PGX_ERR_CUDA, "creating the auxiliary stream failed"));
    }
  }
  for (LogicalPlan* lg : {&plan->or_f, &plan->and_f, &plan->pool_f}) {
    lg->h_ptr = {}; lg->h_pmsg = {}; lg->h_pvs = {}; lg->h_cmsg = {}; lg->h_cvs = {}; lg->h_pw = {}; lg->h_cw = {};
  }
  {  // pull mode: all factors pairwise-binary, every variable of degree <= kPullMaxDegree
    constexpr int64_t kPullMaxDegree = pgx::kPullMaxDegree;
    bool ok = !plan->enum_blocks.empty() && plan->or_f.dev.num_factors == 0 &&
              plan->and_f.dev.num_factors == 0 && plan->pool_f.dev.num_factors == 0;
    for (const EnumBlockPlan& eb : plan->enum_blocks) ok = ok && eb.variant == kPw2;
    for (int64_t v = 0; v < plan->num_vars && ok; ++v) ok = var_ptr[v + 1] - var_ptr[v] <= kPullMaxDegree;
    if (ok) {
      std::vector<int2> edge_csr(plan->num_edges);
      for (int64_t e = 0; e < plan->num_edges; ++e) {
        const int32_t var = vs_var[edge_vs[e]];
        edge_csr[e] = make_int2(int(var_ptr[var]), int(var_ptr[var + 1]));
      }
      PGX_TRY(upload(edge_csr, &plan->d_edge_csr, &plan->device_bytes));
      if (cudaMalloc(reinterpret_cast<void**>(&plan->d_grid_bar), sizeof(unsigned int)) != cudaSuccess)
        return bail(fail(PGX_ERR_CUDA, "cudaMalloc failed"));
      int coop = 0;
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, plan->device);
      if (coop) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&plan->coop_blocks_per_sm[0],
                                                      pgx::k_enum_pw2_pull_resident<false, false>, pgx::kThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&plan->coop_blocks_per_sm[1],
                                                      pgx::k_enum_pw2_pull_resident<true, false>, pgx::kThreads, 0);
      }
      plan->pull_ok = true;
    }
  }
  // binary-difference storage on the generic two-pass path (k_var_sums_bin / k_enum_pw2_bin / k_enum_unary_bin)
  {
    bool ok = plan->num_edge_states == 2 * plan->num_edges && plan->num_var_states == 2 * plan->num_vars &&
              plan->num_edges > 0 && plan->or_f.dev.num_factors == 0 &&
              plan->and_f.dev.num_factors == 0 && plan->pool_f.dev.num_factors == 0;
    for (const EnumBlockPlan& eb : plan->enum_blocks)
      ok = ok && (eb.variant == kPw2 || (eb.variant == kUnary && eb.dev.ns == 2));
    for (int64_t v = 0; v < plan->num_vars && ok; ++v) ok = desc->var_num_states[v] == 2;  // var-states of v: 2v, 2v + 1
    plan->gbin_ok = ok;
  }
  // lattice mode: ONE pairwise-binary block whose factor 2u + t joins variable u with the
  // variable one row below (t = 0) / one column to the right, wrapping (t = 1), and nothing else
  if (plan->enum_blocks.size() == 1 && plan->enum_blocks[0].variant == kPw2 && plan->or_f.dev.num_factors == 0 &&
      plan->and_f.dev.num_factors == 0 && plan->pool_f.dev.num_factors == 0) {
    const EnumBlockPlan& eb = plan->enum_blocks[0];
    const int64_t F = eb.dev.num_factors, e0 = eb.dev.first_edge;
    const int64_t U = F / 2;
    const int64_t N = (F >= 2 && F % 2 == 0) ? edge_vs[e0 + 1] / 2 : 0;
    if (N >= 2 && U % N == 0 && edge_vs[e0] == 0 && (eb.dev.first_msg & 3) == 0 && (eb.dev.first_pot & 3) == 0) {
      const int64_t R = U / N;
      const bool torus = edge_vs[e0 + 4 * (R - 1) * N + 1] == 0;
      const int64_t Rv = torus ? R : R + 1;
      bool ok = plan->num_vars == Rv * N && R >= (torus ? 2 : 1) && R < INT32_MAX && N < INT32_MAX;
      for (int64_t u = 0; u < U && ok; ++u) {
        const int64_t l = u / N, j = u - l * N;
        const int64_t below = torus ? ((l + 1) % R) * N + j : u + N;
        const int64_t right = l * N + (j + 1) % N;
        ok = edge_vs[e0 + 4 * u] == 2 * u && edge_vs[e0 + 4 * u + 2] == 2 * u &&
             edge_vs[e0 + 4 * u + 1] == 2 * below && edge_vs[e0 + 4 * u + 3] == 2 * right;
      }
      if (ok) {
        plan->lattice_ok = true;
        plan->lattice.first_msg = eb.dev.first_msg;
        plan->lattice.first_pot = eb.dev.first_pot;
        plan->lattice.R = int32_t(R);
        plan->lattice.N = int32_t(N);
        plan->lattice.torus = torus ? 1 : 0;
      }
    }
  }
  {  // dense-grid pairwise blocks -> fused single-pass structures
    std::vector<uint8_t> edge_fused(plan->num_edges, 0);
    std::vector<int32_t> part_count(plan->num_vars, 0);
    struct Grid { int blk; int32_t I, J; };
    std::vector<Grid> grids;
    for (size_t bi = 0; bi < plan->enum_blocks.size(); ++bi) {
      EnumBlockPlan& eb = plan->enum_blocks[bi];
      if (eb.variant != kPw2) continue;
      const int64_t F = eb.dev.num_factors, e0 = eb.dev.first_edge;
      int64_t J = 1;
      while (J < F && edge_vs[e0 + 2 * J] == edge_vs[e0]) ++J;
      if (F % J != 0 || F / J < 2 || J < 2 || F >= INT32_MAX) continue;
      bool ok = true;
      for (int64_t f = 0; f < F && ok; ++f)
        ok = edge_vs[e0 + 2 * f] == edge_vs[e0 + 2 * (f / J * J)] && edge_vs[e0 + 2 * f + 1] == edge_vs[e0 + 2 * (f % J) + 1];
      // rows (and columns) must be distinct variables: partial slots are per (variable, block)
      if (!ok) continue;
      grids.push_back({int(bi), int32_t(F / J), int32_t(J)});
    }
    const char* env = getenv("PGX_EXACT_ORDER");
    plan->exact_order = env != nullptr && env[0] == '1';
    if (!grids.empty()) {
      constexpr int TJ = pgx::kBipTJ;
      // slots per variable, in block order
      std::vector<std::vector<int32_t>> row_slot(grids.size()), col_slot(grids.size());
      for (size_t gi = 0; gi < grids.size(); ++gi) {
        const Grid& gr = grids[gi];
        const EnumBlockPlan& eb = plan->enum_blocks[gr.blk];
        const int64_t e0 = eb.dev.first_edge;
        const int32_t NS = (gr.J + TJ - 1) / TJ;
        const int32_t NR = std::max<int32_t>(1, (gr.I + 31) / 32);
        row_slot[gi].resize(gr.I);
        col_slot[gi].resize(gr.J);
        for (int32_t i = 0; i < gr.I; ++i) {
          const int32_t var = vs_var[edge_vs[e0 + 2 * int64_t(i) * gr.J]];
          row_slot[gi][i] = part_count[var];
          part_count[var] += NS;
        }
        for (int32_t j = 0; j < gr.J; ++j) {
          const int32_t var = vs_var[edge_vs[e0 + 2 * int64_t(j) + 1]];
          col_slot[gi][j] = part_count[var];
          part_count[var] += NR;
        }
        for (int64_t e = e0; e < e0 + 2 * eb.dev.num_factors; ++e) edge_fused[e] = 1;
      }
      std::vector<int32_t> part_first(plan->num_vars, 0);
      int64_t rows = 0;
      for (int64_t v = 0; v < plan->num_vars; ++v) {
        part_first[v] = int32_t(rows);
        rows += 2 * int64_t(part_count[v]);
        PGX_REQUIRE(rows < INT32_MAX, "partial-sum buffer too large");
      }
      plan->part_rows = rows;
      plan->bips.resize(grids.size());
      if (cudaStreamCreateWithFlags(&plan->half, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&plan->ev_half_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&plan->ev_half_join, cudaEventDisableTiming) != cudaSuccess)
        return bail(fail(PGX_ERR_CUDA, "creating the half-batch stream failed"));
      if (plan->aux == nullptr && plan->enum_blocks.size() > grids.size()) {
        if (cudaStreamCreateWithFlags(&plan->aux, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&plan->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&plan->ev_join, cudaEventDisableTiming) != cudaSuccess)
          return bail(fail(PGX_ERR_CUDA, "creating the auxiliary stream failed"));
      }
      for (size_t gi = 0; gi < grids.size(); ++gi) {
        const Grid& gr = grids[gi];
        EnumBlockPlan& eb = plan->enum_blocks[gr.blk];
        BipPlan& bp = plan->bips[gi];
        const int64_t e0 = eb.dev.first_edge;
        std::vector<int32_t> row_vs(gr.I), col_vs(gr.J), row_part(gr.I), col_part(gr.J);
        for (int32_t i = 0; i < gr.I; ++i) {
          row_vs[i] = edge_vs[e0 + 2 * int64_t(i) * gr.J];
          row_part[i] = part_first[vs_var[row_vs[i]]] + 2 * row_slot[gi][i];
        }
        for (int32_t j = 0; j < gr.J; ++j) {
          col_vs[j] = edge_vs[e0 + 2 * int64_t(j) + 1];
          col_part[j] = part_first[vs_var[col_vs[j]]] + 2 * col_slot[gi][j];
        }
        PGX_TRY(upload(row_vs, &bp.d_row_vs, &plan->device_bytes));
        PGX_TRY(upload(col_vs, &bp.d_col_vs, &plan->device_bytes));
        PGX_TRY(upload(row_part, &bp.d_row_part, &plan->device_bytes));
        PGX_TRY(upload(col_part, &bp.d_col_part, &plan->device_bytes));
        bp.dev.first_msg = eb.dev.first_msg;
        bp.dev.first_pot = eb.dev.first_pot;
        bp.dev.first_cmsg = plan->c_rows;
        plan->c_rows += 2 * int64_t(gr.I) * gr.J;
        bp.dev.I = gr.I;
        bp.dev.J = gr.J;
        bp.dev.NS = (gr.J + TJ - 1) / TJ;
        bp.dev.NR = std::max<int32_t>(1, (gr.I + 31) / 32);
        bp.dev.RI = (gr.I + bp.dev.NR - 1) / bp.dev.NR;
        bp.dev.NR = (gr.I + bp.dev.RI - 1) / bp.dev.RI;  // no empty chunk
        bp.dev.row_vs = bp.d_row_vs;
        bp.dev.col_vs = bp.d_col_vs;
        bp.dev.row_part = bp.d_row_part;
        bp.dev.col_part = bp.d_col_part;
        eb.bip = int(gi);
      }
      // CSR over the remaining edges (ascending message index within a variable)
      std::vector<int64_t> rest_ptr(plan->num_vars + 1, 0);
      for (int64_t e = 0; e < plan->num_edges; ++e)
        if (!edge_fused[e]) ++rest_ptr[vs_var[edge_vs[e]] + 1];
      for (int64_t v = 0; v < plan->num_vars; ++v) rest_ptr[v + 1] += rest_ptr[v];
      std::vector<int64_t> rest_edge_msg(std::max<int64_t>(rest_ptr[plan->num_vars], 1), 0);
      {
        std::vector<int64_t> cursor(rest_ptr.begin(), rest_ptr.end() - 1);
        for (int64_t e = 0; e < plan->num_edges; ++e)
          if (!edge_fused[e]) rest_edge_msg[cursor[vs_var[edge_vs[e]]]++] = edge_msg_start[e];
      }
      PGX_TRY(upload(narrow(rest_ptr), &plan->d_rest_ptr, &plan->device_bytes));
      PGX_TRY(upload(narrow(rest_edge_msg), &plan->d_rest_edge_msg, &plan->device_bytes));
      PGX_TRY(upload(part_first, &plan->d_part_first, &plan->device_bytes));
      PGX_TRY(upload(part_count, &plan->d_part_count, &plan->device_bytes));
    }
  }
  {  // dominant launch = most edge-states
    int64_t best = -1;
    static const char* const kEnumNames[] = {"k_enum_pw2", "k_enum_small", "k_enum_big", "k_enum_unary"};
    // (with the single-pass path active the pw2 launch of a dense-grid block is k_enum_pw2_bip)
    for (size_t i = 0; i < plan->enum_blocks.size(); ++i) {
      const int64_t es = plan->enum_blocks[i].dev.num_factors * plan->enum_blocks[i].dev.ns;
      if (es > best) { best = es; plan->dominant = int(i); plan->dominant_name = kEnumNames[plan->enum_blocks[i].variant]; }
    }
    int lid = -1;
    for (const LogicalPlan* lg : {&plan->or_f, &plan->and_f, &plan->pool_f}) {
      const int id = lid--;
      const int64_t es = 2 * (lg->dev.num_factors + (lg->dev.num_factors ? desc_num_parents[-id - 1] : 0));
      if (lg->dev.num_factors && es > best) { best = es; plan->dominant = id; plan->dominant_name = id == -3 ? "k_pool" : "k_logical"; }
    }
    plan->dominant_es = std::max<int64_t>(best, 0);
    // (max-product) the dominant block may be part of the merged launch: its units are all merged groups'
    if (plan->dominant >= 0 && size_t(plan->dominant) < plan->enum_blocks.size() &&
        plan->enum_blocks[plan->dominant].bigmax >= 0)
      plan->dominant_es = plan->bigmax_es;
  }
#undef PGX_TRY
#undef PGX_REQUIRE
  if (plan->logical_pull_ok && desc->num_edges < (int64_t(1) << 24)) {
    plan->desc_copy = new DescCopy();
    plan->desc_copy->assign(*desc);
  }
  *out_plan = plan;
  return PGX_OK;
}

void pgx_plan_destroy(pgx_plan* plan) {
  if (plan == nullptr) return;
  free_dev(plan->d_vs_csr);
  free_dev(plan->d_edge_vs); free_dev(plan->d_edge_msg_start); free_dev(plan->d_vs_var);
  free_dev(plan->d_var_first_state); free_dev(plan->d_var_ptr); free_dev(plan->d_var_edge_msg);
  for (EnumBlockPlan& eb : plan->enum_blocks) {
    free_dev(eb.d_cfg_es); free_dev(eb.d_t_ptr); free_dev(eb.d_t_k); free_dev(eb.d_edge_off);
    free_dev(eb.d_fac_edge); free_dev(eb.d_fac_msg); free_dev(eb.d_fac_pot); free_dev(eb.d_rounds); free_dev(eb.d_round_ptr); free_dev(eb.d_rounds_b); free_dev(eb.d_lane_state);
  }
  free_dev(plan->d_bigmax_groups); free_dev(plan->d_bigmax_units); free_dev(plan->d_bigmax_counter);
  for (LogicalPlan* lg : {&plan->or_f, &plan->and_f, &plan->pool_f}) {
    free_dev(lg->d_parent_ptr); free_dev(lg->d_parents_msg); free_dev(lg->d_parents_vs);
    free_dev(lg->d_children_msg); free_dev(lg->d_children_vs);
    free_dev(lg->d_parents_w); free_dev(lg->d_children_w); free_dev(lg->d_parent_factor);
  }
  free_dev(plan->d_hi_list); free_dev(plan->d_fused_w);
  if (plan->ev_fork) cudaEventDestroy(plan->ev_fork);
  if (plan->ev_join) cudaEventDestroy(plan->ev_join);
  if (plan->aux) cudaStreamDestroy(plan->aux);
  if (plan->ev_half_fork) cudaEventDestroy(plan->ev_half_fork);
  if (plan->ev_half_join) cudaEventDestroy(plan->ev_half_join);
  if (plan->half) cudaStreamDestroy(plan->half);
  for (BipPlan& bp : plan->bips) {
    free_dev(bp.d_row_vs); free_dev(bp.d_col_vs); free_dev(bp.d_row_part); free_dev(bp.d_col_part);
  }
  free_dev(plan->d_edge_csr); free_dev(plan->d_grid_bar); free_dev(plan->d_energy_partial);
  free_dev(plan->d_rest_ptr); free_dev(plan->d_rest_edge_msg); free_dev(plan->d_part_first);
  free_dev(plan->d_part_count);
  free_workspace(plan->ws);
  if (plan->tail) pgx_plan_destroy(plan->tail);
  if (plan->ev_tail_fork) cudaEventDestroy(plan->ev_tail_fork);
  if (plan->ev_tail_join) cudaEventDestroy(plan->ev_tail_join);
  if (plan->tail_stream) cudaStreamDestroy(plan->tail_stream);
  delete plan->desc_copy;
  free_dev(plan->d_factor_edge_start);
  free_dev(plan->sdlp.eta); free_dev(plan->sdlp.P); free_dev(plan->sdlp.vval); free_dev(plan->sdlp.eval);
  free_dev(plan->sdlp.grad); free_dev(plan->sdlp.partial);
  for (cudaEvent_t e : plan->prof_events) cudaEventDestroy(e);
  for (pgx_plan::RunGraph& g : plan->run_graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (plan->cap_stream) cudaStreamDestroy(plan->cap_stream);
  delete plan;
}

int pgx_plan_get_info(const pgx_plan* plan, pgx_plan_info* info) {
  if (!plan || !info) return fail(PGX_ERR_INVALID, "null argument");
  info->num_vars = plan->num_vars;
  info->num_var_states = plan->num_var_states;
  info->num_edges = plan->num_edges;
  info->num_edge_states = plan->num_edge_states;
  info->num_potentials = plan->num_potentials;
  info->max_var_states = plan->max_var_states;
  info->device_bytes = plan->device_bytes;
  info->device = plan->device;
  info->num_sms = plan->num_sms;
  return PGX_OK;
}

int64_t pgx_plan_launch_count(const pgx_plan* plan) {
  return plan ? plan->launches + (plan->tail ? plan->tail->launches : 0) : 0;
}

int pgx_plan_set_exact_order(pgx_plan* plan, int enabled) {
  if (!plan) return fail(PGX_ERR_INVALID, "null plan");
  plan->exact_order = enabled != 0;
  return PGX_OK;
}

int pgx_plan_num_fused_blocks(const pgx_plan* plan) { return plan ? int(plan->bips.size()) : 0; }
int64_t pgx_plan_compressed_edges(const pgx_plan* plan) { return plan ? plan->c_rows : 0; }

int pgx_plan_disable_paths(pgx_plan* plan, uint32_t mask) {
  if (!plan) return fail(PGX_ERR_INVALID, "null plan");
  plan->disabled_paths = mask;
  return PGX_OK;
}

int pgx_plan_is_lattice(const pgx_plan* plan) { return plan && plan->lattice_ok ? 1 : 0; }

int pgx_plan_enable_graphs(pgx_plan* plan, int enabled) {
  if (!plan) return fail(PGX_ERR_INVALID, "null plan");
  plan->graphs_enabled = enabled != 0;
  return PGX_OK;
}

int64_t pgx_plan_graph_launch_count(const pgx_plan* plan) { return plan ? plan->graph_launches : 0; }

int64_t pgx_plan_dominant_edge_states(const pgx_plan* plan) { return plan ? plan->dominant_es : 0; }
int64_t pgx_plan_dominant_grid(const pgx_plan* plan) { return plan ? plan->dominant_grid : 0; }

int pgx_plan_profile_enable(pgx_plan* plan, int enabled) {
  if (!plan) return fail(PGX_ERR_INVALID, "null plan");
  plan->profiling = enabled != 0;
  plan->prof_used = 0;
  return PGX_OK;
}

int pgx_plan_profile_read(pgx_plan* plan, int64_t* num_launches, double* total_ms, const char** kernel_name) {
  if (!plan || !num_launches || !total_ms) return fail(PGX_ERR_INVALID, "null argument");
  double total = 0.0;
  const size_t pairs = plan->prof_used / 2;
  for (size_t i = 0; i < pairs; ++i) {
    PGX_CUDA(cudaEventSynchronize(plan->prof_events[2 * i + 1]));
    float ms = 0.f;
    PGX_CUDA(cudaEventElapsedTime(&ms, plan->prof_events[2 * i], plan->prof_events[2 * i + 1]));
    total += ms;
  }
  *num_launches = int64_t(pairs);
  *total_ms = total;
  if (kernel_name) *kernel_name = plan->dominant_name;
  plan->prof_used = 0;
  return PGX_OK;
}

int pgx_bp_run(pgx_plan* plan, void* stream, int64_t batch, const float* log_potentials, int lp_batched,
               const float* evidence, int ev_batched, const float* ftov_in, int msgs_batched,
               float* ftov_out, float* deltas, int32_t num_iters, float damping, float temperature) {
  return pgx_bp_run_flags(plan, stream, batch, log_potentials, lp_batched, evidence, ev_batched, ftov_in,
                          msgs_batched, ftov_out, deltas, num_iters, damping, temperature, 0);
}

static int bp_run_enqueue(pgx_plan* plan, void* stream, int64_t batch, const float* log_potentials, int lp_batched,
                          const float* evidence, int ev_batched, const float* ftov_in, int msgs_batched,
                          float* ftov_out, float* deltas, int32_t num_iters, float damping, float temperature,
                          uint32_t flags);

int pgx_bp_run_flags(pgx_plan* plan, void* stream, int64_t batch, const float* log_potentials, int lp_batched,
                     const float* evidence, int ev_batched, const float* ftov_in, int msgs_batched,
                     float* ftov_out, float* deltas, int32_t num_iters, float damping, float temperature,
                     uint32_t flags) {
  if (!plan) return fail(PGX_ERR_INVALID, "null plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
  static const bool env_off = [] {
    const char* v = getenv("PGX_GRAPH");
    return v != nullptr && v[0] == '0';
  }();
  // OR / AND graphs on the pull path run their two groups on two streams of different PRIORITY
  // (the long serial OR chains get free SM slots first); a captured graph does not keep that
  // scheduling and measured slower (deconvolution, B = 100: 0.295 ms per iteration replayed
  // against 0.260 enqueued, profiles/r02_h_graph_ab.txt): such runs stay on the direct path.
  const bool priority_streams = plan->logical_pull_ok && plan->aux != nullptr && batch >= 32 &&
                                !(plan->orand_fused_ok && !(plan->disabled_paths & (PGX_PATH_ORAND_FUSED | PGX_PATH_LOGICAL_BIN)));
  const bool eligible = plan->graphs_enabled && !env_off && !plan->profiling && !(flags & PGX_RUN_NO_GRAPH) && num_iters >= 2 &&
                        !priority_streams &&
                        plan->num_edge_states > 0 && cudaStreamIsCapturing(st, &capturing) == cudaSuccess &&
                        capturing == cudaStreamCaptureStatusNone;
  if (!eligible)
    return bp_run_enqueue(plan, stream, batch, log_potentials, lp_batched, evidence, ev_batched, ftov_in, msgs_batched,
                          ftov_out, deltas, num_iters, damping, temperature, flags);
  int rc;
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(plan))) return rc;
  const auto epoch_now = [plan] { return plan->ws_epoch + (plan->tail ? plan->tail->ws_epoch : 0); };
  pgx_plan::RunGraph* hit = nullptr;
  for (pgx_plan::RunGraph& g : plan->run_graphs)
    if (g.lp == log_potentials && g.ev == evidence && g.in == ftov_in && g.out == ftov_out && g.deltas == deltas &&
        g.batch == batch && g.iters == num_iters && g.lp_b == lp_batched && g.ev_b == ev_batched && g.in_b == msgs_batched &&
        g.damping == damping && g.temperature == temperature && g.flags == flags && g.paths == plan->disabled_paths &&
        g.exact == plan->exact_order)
      hit = &g;
  if (hit == nullptr) {
    // first sight: run directly (this also sizes the workspace) and remember the signature
    if (plan->run_graphs.size() >= 16) {
      if (plan->run_graphs.front().exec) cudaGraphExecDestroy(plan->run_graphs.front().exec);
      plan->run_graphs.erase(plan->run_graphs.begin());
    }
    plan->run_graphs.push_back(pgx_plan::RunGraph{log_potentials, evidence, ftov_in, ftov_out, deltas, batch, num_iters,
                                                  lp_batched, ev_batched, msgs_batched, damping, temperature, flags,
                                                  plan->disabled_paths, plan->exact_order, nullptr, 0, 0, 0});
    return bp_run_enqueue(plan, stream, batch, log_potentials, lp_batched, evidence, ev_batched, ftov_in, msgs_batched,
                          ftov_out, deltas, num_iters, damping, temperature, flags);
  }
  if (hit->exec != nullptr && hit->epoch != epoch_now()) {  // the workspace moved since the capture
    cudaGraphExecDestroy(hit->exec);
    hit->exec = nullptr;
  }
  if (hit->exec == nullptr) {
    if (plan->cap_stream == nullptr) PGX_CUDA(cudaStreamCreateWithFlags(&plan->cap_stream, cudaStreamNonBlocking));
    const int64_t own0 = plan->launches, tail0 = plan->tail ? plan->tail->launches : 0;
    PGX_CUDA(cudaStreamBeginCapture(plan->cap_stream, cudaStreamCaptureModeRelaxed));
    rc = bp_run_enqueue(plan, plan->cap_stream, batch, log_potentials, lp_batched, evidence, ev_batched, ftov_in,
                        msgs_batched, ftov_out, deltas, num_iters, damping, temperature, flags);
    cudaGraph_t graph = nullptr;
    const cudaError_t err = cudaStreamEndCapture(plan->cap_stream, &graph);
    const int64_t captured = (plan->launches - own0) + (plan->tail ? plan->tail->launches - tail0 : 0);
    plan->launches = own0;  // nothing ran yet: a replay adds `captured` to this plan's counter
    if (plan->tail) plan->tail->launches = tail0;
    if (rc != PGX_OK || err != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      // not capturable here: this signature stays on the direct path
      hit->flags |= 0x80000000u;  // never matches again
      if (rc != PGX_OK) return rc;
      return bp_run_enqueue(plan, stream, batch, log_potentials, lp_batched, evidence, ev_batched, ftov_in, msgs_batched,
                            ftov_out, deltas, num_iters, damping, temperature, flags);
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ierr = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ierr != cudaSuccess) {
      cudaGetLastError();
      hit->flags |= 0x80000000u;
      return bp_run_enqueue(plan, stream, batch, log_potentials, lp_batched, evidence, ev_batched, ftov_in, msgs_batched,
                            ftov_out, deltas, num_iters, damping, temperature, flags);
    }
    hit->exec = exec;
    hit->launches = captured;
    hit->epoch = epoch_now();
    hit->sums_batch = plan->final_sums_batch;
  }
  PGX_CUDA(cudaGraphLaunch(hit->exec, st));
  plan->final_sums_batch = hit->sums_batch;
  plan->launches += hit->launches;
  ++plan->graph_launches;
  return PGX_OK;
}

static int bp_run_enqueue(pgx_plan* plan, void* stream, int64_t batch, const float* log_potentials, int lp_batched,
                          const float* evidence, int ev_batched, const float* ftov_in, int msgs_batched,
                          float* ftov_out, float* deltas, int32_t num_iters, float damping, float temperature,
                          uint32_t flags) {
  PGX_CHECK(batch >= 1 && batch < (1 << 24), "batch must be in [1, 2^24), got %lld", (long long)batch);
  PGX_CHECK(num_iters >= 1, "num_iters must be >= 1, got %d", num_iters);
  PGX_CHECK(temperature >= 0.f, "temperature must be >= 0");
  PGX_CHECK(ftov_out != nullptr || plan->num_edge_states == 0 || (flags & PGX_RUN_SKIP_OUTPUT), "ftov_out is null");
  PGX_CHECK(plan->num_var_states == 0 || evidence != nullptr, "evidence is null");
  PGX_CHECK(plan->num_potentials == 0 || log_potentials != nullptr, "log_potentials is null");
  int rc;
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(plan))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const pgx::BatchMap mp = make_map(batch);
  plan->final_sums_batch = 0;
  PGX_CHECK(tiled_floats(mp, plan->num_edge_states) < (size_t(1) << 40), "workspace too large");
  const bool single = batch == 1;
  const bool evT = !single && ev_batched, lpT = !single && lp_batched;
  // Single-pass mode: dense-grid pairwise blocks emit per-tile partial sums of the new
  // messages; needs full warps of samples and potentials shared by the batch.
  const bool fused = !plan->bips.empty() && !plan->exact_order && mp.bx_log == 5 && !lpT;
  const int64_t Es = plan->num_edge_states, Vs = plan->num_var_states, C = plan->num_potentials;
  // Logical pull path: full sample tiles, 32-bit offsets inside a tile
  const bool lpull = plan->logical_pull_ok && mp.bx_log == 5 && !(plan->disabled_paths & PGX_PATH_LOGICAL_PULL) &&
                     Es < (int64_t(1) << 26) && Vs < (int64_t(1) << 26);
  // ... with the messages in binary-difference storage (such a graph has only two-state edges)
  const bool lbin = lpull && !(plan->disabled_paths & PGX_PATH_LOGICAL_BIN) && Es == 2 * plan->num_edges;
  // Pull mode (small pairwise-binary graphs on low-degree variables, see below) is decided here: the
  // choice of message storage depends on it.
  bool pull_early = false;
  if (plan->pull_ok && !fused && plan->enum_blocks.size() == 1 && !(plan->disabled_paths & PGX_PATH_PULL)) {
    const int upw0 = 32 >> mp.bx_log;
    const int64_t warps0 = ((plan->enum_blocks[0].dev.num_factors + upw0 - 1) / upw0) * mp.nbt;
    const int ks0 = temperature == 0.f ? 0 : 1;
    pull_early = warps0 * 32 <= int64_t(plan->coop_blocks_per_sm[ks0]) * plan->num_sms * pgx::kThreads;
  }
  // Generic two-pass path of an all-binary pairwise graph, full sample tiles: binary-difference storage
  // (not for plans with dense-grid blocks, even when those run two-pass - exact order: the workspace's
  // compressed buffers are sized for the single-pass path there)
  const bool gbin = plan->gbin_ok && plan->bips.empty() && mp.bx_log == 5 && !fused && !lpull && !pull_early &&
                    !(plan->disabled_paths & PGX_PATH_GENERIC_BIN);
  const bool cbin = lbin || gbin;  // messages live in ws.cA / ws.cB, one float per edge
  // Batch tail: a few samples beyond the last full tile run through the tail plan on its own stream.
  const int64_t tail_n = batch & 31;
  // (not with the fused OR + AND launch: there one more - partial - tile costs less than the tail
  // plan's generic kernels, whose serial wide-OR update then bounds the iteration: deconvolution
  // B = 100, 0.235 ms per iteration split against 0.227 as four tiles)
  const bool fused_orand_run = lbin && plan->orand_fused_ok && !(plan->disabled_paths & PGX_PATH_ORAND_FUSED);
  if (lpull && batch > 32 && tail_n != 0 && tail_n <= pgx::kTailMaxSamples && plan->desc_copy != nullptr &&
      !fused_orand_run && !(plan->disabled_paths & PGX_PATH_TAIL_SPLIT)) {
    const int64_t main_n = batch - tail_n;
    if (plan->tail == nullptr) {
      if ((rc = pgx_plan_create(&plan->desc_copy->desc, &plan->tail))) return rc;
      delete plan->tail->desc_copy;  // the tail never splits again
      plan->tail->desc_copy = nullptr;
      PGX_CUDA(cudaStreamCreateWithFlags(&plan->tail_stream, cudaStreamNonBlocking));
      PGX_CUDA(cudaEventCreateWithFlags(&plan->ev_tail_fork, cudaEventDisableTiming));
      PGX_CUDA(cudaEventCreateWithFlags(&plan->ev_tail_join, cudaEventDisableTiming));
    }
    plan->tail->disabled_paths = plan->disabled_paths;
    plan->tail->exact_order = plan->exact_order;
    PGX_CUDA(cudaEventRecord(plan->ev_tail_fork, st));
    PGX_CUDA(cudaStreamWaitEvent(plan->tail_stream, plan->ev_tail_fork, 0));
    auto at = [&](const float* p, int batched, int64_t rows) { return (p && batched) ? p + main_n * rows : p; };
    if ((rc = pgx_bp_run_flags(plan->tail, plan->tail_stream, tail_n, at(log_potentials, lp_batched, C), lp_batched,
                               at(evidence, ev_batched, Vs), ev_batched, at(ftov_in, msgs_batched, Es), msgs_batched,
                               ftov_out + main_n * Es, deltas ? deltas + main_n * num_iters : nullptr, num_iters, damping,
                               temperature, flags)))
      return rc;
    if ((rc = pgx_bp_run_flags(plan, stream, main_n, log_potentials, lp_batched, evidence, ev_batched, ftov_in,
                               msgs_batched, ftov_out, deltas, num_iters, damping, temperature, flags)))
      return rc;
    PGX_CUDA(cudaEventRecord(plan->ev_tail_join, plan->tail_stream));
    PGX_CUDA(cudaStreamWaitEvent(st, plan->ev_tail_join, 0));
    return PGX_OK;
  }
  if (Es == 0) {  // a graph without factors: nothing to update, every delta is 0
    if (deltas) PGX_CUDA(cudaMemsetAsync(deltas, 0, size_t(batch) * num_iters * sizeof(float), static_cast<cudaStream_t>(stream)));
    return PGX_OK;
  }
  // Large single-sample lattices run on binary-difference storage (k_lattice_bin): decided here
  // because the compressed buffers belong to the workspace.
  bool lat_bin = false;
  if (plan->lattice_ok && single &&
      !(plan->disabled_paths & (PGX_PATH_LATTICE | PGX_PATH_LATTICE_STREAM | PGX_PATH_LATTICE_BIN))) {
    const pgx::LatticeDev& g = plan->lattice;
    const int64_t num_tiles = int64_t((g.N + pgx::kLsTC - 1) / pgx::kLsTC) * ((g.R + pgx::kLsTR - 1) / pgx::kLsTR);
    const auto misaligned = [](const void* p, uintptr_t mask) { return (reinterpret_cast<uintptr_t>(p) & mask) != 0; };
    lat_bin = num_tiles >= 4 * int64_t(plan->num_sms) && !misaligned(ftov_in, 15) && !misaligned(ftov_out, 15) &&
              !misaligned(log_potentials, 15) && !misaligned(evidence, 7);
  }
  if ((rc = ensure_workspace(plan, batch, evT, lpT, fused, cbin || lat_bin))) return rc;
  Workspace& ws = plan->ws;
  if (lat_bin) return run_lattice_bin(plan, st, log_potentials, evidence, ftov_in, ftov_out, deltas, num_iters, damping,
                                      temperature);

  // ---- inputs -> tile-blocked workspace ----------------------------------------------------
  pgx::View ev{evidence, Vs, 0}, lp{log_potentials, C, 0};
  if (evT) {
    if ((rc = to_tiles(plan, st, evidence, ws.evT, Vs, mp))) return rc;
    ev = pgx::View{ws.evT, Vs, 1};
  }
  if (lpT) {
    if ((rc = to_tiles(plan, st, log_potentials, ws.lpT, C, mp))) return rc;
    lp = pgx::View{ws.lpT, C, 1};
  }
  // Messages the caller guarantees to be normalised already (output of a previous run) are,
  // for one sample, read in place: no staging copy, no normalisation pass.
  const bool in_place = single && ftov_in != nullptr && (flags & PGX_RUN_INPUT_NORMALIZED) != 0 &&
                        ftov_in != ftov_out;
  const float* cur = ws.mA;
  float* nxt = ws.mB;
  // Single-pass mode with initial messages shared by the batch (or absent): ONE [Es] vector is
  // normalised; the fused blocks' part goes straight to binary-difference storage, only the
  // other edges' rows are broadcast, and the first variable sums read the shared vector.
  const bool shared_init = fused && !(ftov_in != nullptr && msgs_batched);
  if (shared_init) {
    if (ftov_in == nullptr) {
      PGX_CUDA(cudaMemsetAsync(ws.row, 0, size_t(Es) * sizeof(float), st));
    } else {
      PGX_CUDA(cudaMemcpyAsync(ws.row, ftov_in, size_t(Es) * sizeof(float), cudaMemcpyDeviceToDevice, st));
      if ((flags & PGX_RUN_INPUT_NORMALIZED) == 0) {
        if ((rc = normalize_edges(plan, st, make_map(1), ws.row))) return rc;
      }
    }
    int64_t done = 0;
    for (const BipPlan* bpp : bips_by_msg(plan)) {
      const pgx::BipDev& g = bpp->dev;
      const int64_t count = 2 * int64_t(g.I) * g.J;
      if (g.first_msg > done) {
        pgx::k_broadcast_rows_range<<<plan->num_sms * 4, pgx::kThreads, 0, st>>>(ws.row, ws.mA, Es, done, g.first_msg, mp);
        if ((rc = check_launch(plan, "k_broadcast_rows_range"))) return rc;
      }
      pgx::k_broadcast_bin<<<plan->num_sms * 8, pgx::kThreads, 0, st>>>(ws.row, g.first_msg, ws.cB, plan->c_rows, g.first_cmsg,
                                                                   count, mp.nbt);
      if ((rc = check_launch(plan, "k_broadcast_bin"))) return rc;
      done = g.first_msg + 2 * count;
    }
    if (Es > done) {
      pgx::k_broadcast_rows_range<<<plan->num_sms * 4, pgx::kThreads, 0, st>>>(ws.row, ws.mA, Es, done, Es, mp);
      if ((rc = check_launch(plan, "k_broadcast_rows_range"))) return rc;
    }
  } else if (in_place) {
    cur = ftov_in;
    nxt = ws.mA;
  } else if (ftov_in == nullptr) {
    if (cbin)
      PGX_CUDA(cudaMemsetAsync(ws.cA, 0, tiled_floats(mp, Es / 2) * sizeof(float), st));
    else
      PGX_CUDA(cudaMemsetAsync(ws.mA, 0, tiled_floats(mp, Es) * sizeof(float), st));  // NC(0) = 0
  } else {
    if (single) {
      PGX_CUDA(cudaMemcpyAsync(ws.mA, ftov_in, size_t(Es) * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else if (msgs_batched) {
      if ((rc = to_tiles(plan, st, ftov_in, ws.mA, Es, mp))) return rc;
    } else {
      pgx::k_broadcast_rows<<<plan->num_sms * 8, pgx::kThreads, 0, st>>>(ftov_in, ws.mA, Es, mp);
      if ((rc = check_launch(plan, "k_broadcast_rows"))) return rc;
    }
    if ((flags & PGX_RUN_INPUT_NORMALIZED) == 0) {
      if ((rc = normalize_edges(plan, st, mp, ws.mA))) return rc;
    }
  }
  // the two buffers the generic loop below ping-pongs between
  float* const bufA = cbin ? ws.cA : ws.mA;
  float* const bufB = cbin ? ws.cB : ws.mB;
  if (cbin) {
    if (ftov_in != nullptr) {  // normalised full layout -> one float per edge
      pgx::k_compress_bin<<<plan->num_sms * 8, pgx::kThreads, 0, st>>>(ws.mA, ws.cA, Es / 2, mp.nbt);
      if ((rc = check_launch(plan, "k_compress_bin"))) return rc;
    }
    cur = ws.cA;
    nxt = ws.cB;
  }
  if (deltas) PGX_CUDA(cudaMemsetAsync(deltas, 0, size_t(batch) * num_iters * sizeof(float), st));

  // ---- iterations ------------------------------------------------------------------------
  pgx::RunArgs a{};
  a.d = damping;
  a.one_minus_d = 1.0f - damping;
  a.T = temperature;
  if (temperature > 0.f) {
    a.c_exp = float(1.4426950408889634 / double(temperature));
    a.c_log = float(0.6931471805599453 * double(temperature));
  }
  a.deltas = deltas;
  a.delta_stride = num_iters;
  a.Es = Es;
  a.Vs = Vs;
  // Pull mode: graphs made only of pairwise-binary factors on low-degree variables (grids)
  // need no var-sum array: one kernel per block and iteration, bit-identical to the
  // two-pass path.  Small such graphs with a single block run ALL iterations in one
  // cooperative launch.
  // (Large grids keep the two-pass path: re-deriving S per edge costs 4x the gathers and
  // measured slower than k_var_sums + k_enum_pw2 once the graph no longer fits in cache.)
  // auxiliary stream: the smaller logical group beside the larger one (pull path), or the
  // non-fused enum blocks (unary factors of an RBM) beside the fused dense-grid kernel - they
  // read the same old messages and sums and write disjoint message ranges
  const bool side_blocks = fused && plan->enum_blocks.size() > plan->bips.size() && plan->or_f.dev.num_factors == 0 &&
                           plan->and_f.dev.num_factors == 0 && plan->pool_f.dev.num_factors == 0;
  // half-batch pipeline (see pgx_plan::half); not while the profiling events bracket whole launches
  const bool split = fused && mp.nbt >= 16 && plan->half != nullptr && !plan->profiling &&
                     !(plan->disabled_paths & PGX_PATH_HALF_BATCH);
  const bool orand_fused = lpull && lbin && plan->orand_fused_ok && !(plan->disabled_paths & PGX_PATH_ORAND_FUSED);
  const cudaStream_t aux = ((lpull || (side_blocks && !split)) && plan->aux != nullptr && !orand_fused &&
                            !(plan->disabled_paths & PGX_PATH_AUX_STREAM))
                               ? plan->aux : nullptr;
  const bool aux_after_s = lpull ? plan->aux_needs_s : true;
  bool pull = pull_early;
  // Lattice mode (one sample): one index-free kernel per iteration, bit-identical to the
  // two-pass path.  Graphs small enough for the resident kernels keep those (no launches).
  bool lattice = false;
  const bool use_resident = pull && num_iters >= 2 && !plan->profiling && !(plan->disabled_paths & PGX_PATH_RESIDENT);
  if (plan->lattice_ok && single && !(plan->disabled_paths & PGX_PATH_LATTICE) && !use_resident) {
    const auto misaligned = [](const void* p, uintptr_t mask) { return (reinterpret_cast<uintptr_t>(p) & mask) != 0; };
    lattice = !misaligned(cur, 15) && !misaligned(ws.mA, 15) && !misaligned(ws.mB, 15) && !misaligned(ftov_out, 15) &&
              !misaligned(log_potentials, 15) && !misaligned(evidence, 7);
  }
  if (lattice) {
    const pgx::LatticeDev& g = plan->lattice;
    const dim3 grid(unsigned((g.N + pgx::kLatTC - 1) / pgx::kLatTC), unsigned((g.R + pgx::kLatTR - 1) / pgx::kLatTR));
    // large lattices stream through the persistent TMA kernel, small ones (a few tiles per SM)
    // keep the one-tile-per-CTA kernel
    const int64_t num_tiles = int64_t((g.N + pgx::kLsTC - 1) / pgx::kLsTC) * ((g.R + pgx::kLsTR - 1) / pgx::kLsTR);  // streaming tiles
    const bool stream = num_tiles >= 4 * int64_t(plan->num_sms) && !(plan->disabled_paths & PGX_PATH_LATTICE_STREAM);
    const bool want_delta = deltas != nullptr;
    const int variant = (temperature == 0.f ? 0 : 2) + (want_delta ? 1 : 0);
    using LatFn = void (*)(pgx::LatticeDev, const float*, const float*, const float*, float*, pgx::RunArgs);
    static const LatFn tile_fn[4] = {pgx::k_lattice<false, false>, pgx::k_lattice<false, true>, pgx::k_lattice<true, false>,
                                     pgx::k_lattice<true, true>};
    static const LatFn stream_fn[4] = {pgx::k_lattice_stream<false, false>, pgx::k_lattice_stream<false, true>,
                                       pgx::k_lattice_stream<true, false>, pgx::k_lattice_stream<true, true>};
    if (!((plan->attr_done >> kAttrLattice) & 1u)) {
      plan->attr_done |= 1u << kAttrLattice;
      for (int v = 0; v < 4; ++v) {
        PGX_CUDA(cudaFuncSetAttribute(tile_fn[v], cudaFuncAttributeMaxDynamicSharedMemorySize, int(pgx::lattice_smem_bytes())));
        PGX_CUDA(cudaFuncSetAttribute(stream_fn[v], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(pgx::lattice_stream_smem_bytes())));
      }
    }
    plan->dominant_name = stream ? "k_lattice_stream" : "k_lattice";
    plan->dominant_grid = stream ? std::min<int64_t>(num_tiles, plan->num_sms) : int64_t(grid.x) * grid.y;
    for (int it = 0; it < num_iters; ++it) {
      a.delta_off = it;
      float* dst = (it == num_iters - 1) ? ftov_out : nxt;
      if ((rc = prof_mark(plan, st, 0))) return rc;
      if (stream) {
        const unsigned sgrid = unsigned(std::min<int64_t>(num_tiles, plan->num_sms));
        stream_fn[variant]<<<sgrid, pgx::kLsThreads, pgx::lattice_stream_smem_bytes(), st>>>(g, evidence, log_potentials, cur,
                                                                                          dst, a);
      } else {
        tile_fn[variant]<<<grid, pgx::kLatThreads, pgx::lattice_smem_bytes(), st>>>(g, evidence, log_potentials, cur, dst, a);
      }
      if ((rc = check_launch(plan, "k_lattice"))) return rc;
      if ((rc = prof_mark(plan, st, 0))) return rc;
      nxt = (dst == ws.mA) ? ws.mB : ws.mA;
      cur = dst;
    }
    pull = false;
  }
  if (pull) {
    auto pull_args = [&](const EnumBlockPlan& eb) {
      pgx::PullArgs g;
      g.num_factors = eb.dev.num_factors;
      g.first_edge = eb.dev.first_edge;
      g.first_msg = eb.dev.first_msg;
      g.first_pot = eb.dev.first_pot;
      g.edge_vs = plan->d_edge_vs;
      g.edge_csr = plan->d_edge_csr;
      g.var_edge_msg = plan->d_var_edge_msg;
      return g;
    };
    const int ks = temperature == 0.f ? 0 : 1;
    const EnumBlockPlan& eb0 = plan->enum_blocks[0];
    // resident kernels: one thread per (factor, sample)
    const int upw = 32 >> mp.bx_log;
    const int64_t res_warps = ((eb0.dev.num_factors + upw - 1) / upw) * mp.nbt;
    const bool res_ok = use_resident;
    const bool cluster = res_ok && res_warps * 32 <= int64_t(pgx::kResidentClusterCtas) * pgx::kResidentClusterThreads;
    const int64_t coop_cap = int64_t(plan->coop_blocks_per_sm[ks]) * plan->num_sms;
    const int64_t coop_blocks = (res_warps * 32 + pgx::kThreads - 1) / pgx::kThreads;
    bool cluster_done = false;
    pgx::PullArgs g0 = pull_args(eb0);
    pgx::BatchMap mpv = mp;
    float* out = single ? ftov_out : nullptr;
    int iters = num_iters;
    if (cluster) {
      const int ctas = int(std::min<int64_t>(pgx::kResidentClusterCtas, (res_warps + 3) / 4));
      const int threads = int((res_warps + ctas - 1) / ctas) * 32;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(ctas);
      cfg.blockDim = dim3(threads);
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = ctas;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      unsigned int* no_bar = nullptr;
      cudaError_t err;
      if (ks) {
        cudaFuncSetAttribute(pgx::k_enum_pw2_pull_resident<true, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        err = cudaLaunchKernelEx(&cfg, pgx::k_enum_pw2_pull_resident<true, true>, mpv, g0, ev, lp, cur, ws.mA, ws.mB,
                                 out, iters, a, no_bar);
      } else {
        cudaFuncSetAttribute(pgx::k_enum_pw2_pull_resident<false, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        err = cudaLaunchKernelEx(&cfg, pgx::k_enum_pw2_pull_resident<false, true>, mpv, g0, ev, lp, cur, ws.mA, ws.mB,
                                 out, iters, a, no_bar);
      }
      if (err == cudaSuccess) {
        if ((rc = check_launch(plan, "k_enum_pw2_pull_resident<cluster>"))) return rc;
        cluster_done = true;
      } else {
        cudaGetLastError();  // cluster shape not schedulable here: use the cooperative variant
      }
    }
    const bool coop = res_ok && !cluster_done && coop_blocks <= coop_cap;
    if (coop) {
      PGX_CUDA(cudaMemsetAsync(plan->d_grid_bar, 0, sizeof(unsigned int), st));
      void* args[] = {&mpv, &g0, &ev, &lp, &cur, &ws.mA, &ws.mB, &out, &iters, &a, &plan->d_grid_bar};
      const void* fn = ks ? reinterpret_cast<const void*>(pgx::k_enum_pw2_pull_resident<true, false>)
                          : reinterpret_cast<const void*>(pgx::k_enum_pw2_pull_resident<false, false>);
      PGX_CUDA(cudaLaunchCooperativeKernel(fn, dim3(unsigned(coop_blocks)), dim3(pgx::kThreads), args, 0, st));
      if ((rc = check_launch(plan, "k_enum_pw2_pull_resident<coop>"))) return rc;
    }
    if (cluster_done || coop) {
      // where the kernel left the final messages (same rule as in the kernel)
      float* nx = (cur == ws.mA) ? ws.mB : ws.mA;
      for (int it = 0; it < num_iters; ++it) {
        float* dst = (it == num_iters - 1 && out != nullptr) ? out : nx;
        nx = (dst == ws.mA) ? ws.mB : ws.mA;
        cur = dst;
      }
    } else {
      for (int it = 0; it < num_iters; ++it) {
        a.delta_off = it;
        float* dst = (single && it == num_iters - 1) ? ftov_out : nxt;
        for (size_t bi = 0; bi < plan->enum_blocks.size(); ++bi) {
          const EnumBlockPlan& eb = plan->enum_blocks[bi];
          const dim3 grid = grid_for(plan, mp, eb.dev.num_factors);
          if ((rc = prof_mark(plan, st, int(bi)))) return rc;
          if (ks)
            pgx::k_enum_pw2_pull<true><<<grid, pgx::kThreads, 0, st>>>(mp, pull_args(eb), ev, lp, cur, dst, a);
          else
            pgx::k_enum_pw2_pull<false><<<grid, pgx::kThreads, 0, st>>>(mp, pull_args(eb), ev, lp, cur, dst, a);
          if ((rc = check_launch(plan, "k_enum_pw2_pull"))) return rc;
          if (int(bi) == plan->dominant) plan->dominant_name = "k_enum_pw2_pull";
          if ((rc = prof_mark(plan, st, int(bi)))) return rc;
        }
        nxt = (dst == ws.mA) ? ws.mB : ws.mA;
        cur = dst;
      }
    }
  }
  // Merged max-product launch (RCN-size factors): with potentials shared by the batch and enough
  // iterations to amortise it, one pass copies the potentials into round order first.
  plan->bigmax_perm_active = false;
  if (plan->bigmax_units > 0 && !(plan->disabled_paths & (PGX_PATH_MERGED_MAX | PGX_PATH_PERM_POTENTIALS)) &&
      !pull && !lattice && lp.kind == 0 && num_iters >= 3 && plan->bigmax_perm_floats > 0) {
    if (ws.lpR == nullptr) {
      PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(&ws.lpR), size_t(plan->bigmax_perm_floats) * sizeof(float)));
      ++plan->ws_epoch;
      plan->lpR_src = nullptr;
    }
    // the copy is kept across runs: a caller that promises unchanged potentials skips the pass
    if (!((flags & PGX_RUN_POTENTIALS_UNCHANGED) && plan->lpR_src == log_potentials)) {
      const dim3 grid(8, unsigned(std::min<int64_t>(plan->bigmax_units, 65535)));
      pgx::k_bigmax_permute<true><<<grid, pgx::kThreads, 0, st>>>(mp, plan->d_bigmax_groups, plan->d_bigmax_units,
                                                                plan->bigmax_units, lp, ws.lpR);
      if ((rc = check_launch(plan, "k_bigmax_permute"))) return rc;
      plan->lpR_src = log_potentials;
    }
    plan->bigmax_perm_active = true;
  }
  for (int it = 0; it < ((pull || lattice) ? 0 : num_iters); ++it) {
    a.delta_off = it;
    if (aux != nullptr && !aux_after_s) {
      PGX_CUDA(cudaEventRecord(plan->ev_fork, st));
      PGX_CUDA(cudaStreamWaitEvent(aux, plan->ev_fork, 0));
    }
    if (lpull) {
      if (plan->hi_len > 0) {
        const int64_t per_tile = std::min<int64_t>(plan->hi_len, (int64_t(1) << 30) / mp.nbt);
        if (lbin) {
          // the list is sorted by degree, longest first: its first hi_big variables (>= kVsBigDegree edges)
          // take the cooperative kernel, the rest a warp each
          const int64_t big = (plan->disabled_paths & PGX_PATH_VARSUM_COOP) ? 0 : plan->hi_big;
          // ... on the auxiliary stream beside the warp-per-variable launch when that stream is free
          // (fused OR + AND path: nothing else uses it)
          const bool side = big > 0 && plan->aux != nullptr && plan->orand_fused_ok &&
                            !(plan->disabled_paths & (PGX_PATH_ORAND_FUSED | PGX_PATH_AUX_STREAM));
          if (big > 0) {
            const cudaStream_t sb = side ? plan->aux : st;
            if (side) {
              PGX_CUDA(cudaEventRecord(plan->ev_fork, st));
              PGX_CUDA(cudaStreamWaitEvent(plan->aux, plan->ev_fork, 0));
            }
            pgx::k_var_sums_big_bin<<<unsigned(big * mp.nbt), pgx::kVsBigWarps * 32, 0, sb>>>(
                mp.batch, mp.nbt, Es / 2, Vs, plan->d_vs_csr, plan->d_var_edge_msg, plan->d_hi_list, ev, cur, ws.S);
            if ((rc = check_launch(plan, "k_var_sums_big_bin"))) return rc;
            if (side) PGX_CUDA(cudaEventRecord(plan->ev_join, plan->aux));
          }
          const int64_t rest = plan->hi_len - 2 * big;
          if (rest > 0) {
            pgx::k_var_sums_list_bin<<<unsigned(std::max<int64_t>(std::min<int64_t>(per_tile, rest) / 2, 1) * mp.nbt), 32, 0, st>>>(
                mp.batch, mp.nbt, Es / 2, Vs, plan->d_vs_csr, plan->d_var_edge_msg, plan->d_hi_list + 2 * big, rest, ev,
                cur, ws.S);
            if ((rc = check_launch(plan, "k_var_sums_list"))) return rc;
          }
          if (side) PGX_CUDA(cudaStreamWaitEvent(st, plan->ev_join, 0));
        } else {
          pgx::k_var_sums_list<<<unsigned(per_tile * mp.nbt), 32, 0, st>>>(mp.batch, mp.nbt, Es, Vs, plan->d_vs_csr,
                                                                        plan->d_var_edge_msg, plan->d_hi_list, plan->hi_len,
                                                                        ev, cur, ws.S);
          if ((rc = check_launch(plan, "k_var_sums_list"))) return rc;
        }
      }
    } else if (gbin) {
      pgx::k_var_sums_bin<<<grid_for(plan, mp, plan->num_vars), pgx::kThreads, 0, st>>>(
          mp, plan->num_vars, Es / 2, Vs, plan->d_vs_csr, plan->d_var_edge_msg, ev, cur, ws.S);
      if ((rc = check_launch(plan, "k_var_sums_bin"))) return rc;
    } else if (!fused || it == 0) {
      pgx::k_var_sums<<<grid_for(plan, mp, Vs), pgx::kThreads, 0, st>>>(
          mp, Vs, Es, plan->d_vs_csr, plan->d_var_edge_msg, ev, shared_init ? ws.row : cur, ws.S, shared_init ? 1 : 0);
      if ((rc = check_launch(plan, "k_var_sums"))) return rc;
    }
    if (aux != nullptr && aux_after_s) {
      PGX_CUDA(cudaEventRecord(plan->ev_fork, st));
      PGX_CUDA(cudaStreamWaitEvent(aux, plan->ev_fork, 0));
    }
    // With one sample the last iteration writes straight into the caller's buffer.
    float* dst = (single && it == num_iters - 1) ? ftov_out : nxt;
    // fused blocks keep their messages compressed (ws.cA / ws.cB) after the first iteration
    const float* c_old = (fused && (it > 0 || shared_init)) ? ((it & 1) ? ws.cA : ws.cB) : nullptr;
    float* c_new = fused ? ((it & 1) ? ws.cB : ws.cA) : nullptr;
    if (split) {
      if (it == 0) {  // the first variable sums (whole batch, main stream) feed both halves
        PGX_CUDA(cudaEventRecord(plan->ev_half_fork, st));
        PGX_CUDA(cudaStreamWaitEvent(plan->half, plan->ev_half_fork, 0));
      }
      for (int h = 0; h < 2; ++h) {
        const cudaStream_t sh = h ? plan->half : st;
        const int tile0 = h ? mp.nbt / 2 : 0, nt = h ? mp.nbt - mp.nbt / 2 : mp.nbt / 2;
        const pgx::BatchMap mph{int(std::min<int64_t>(batch - int64_t(tile0) * 32, int64_t(nt) * 32)), 5, nt};
        auto off = [&](int64_t rows) { return size_t(tile0) * size_t(rows) * 32; };
        pgx::RunArgs ah = a;
        if (deltas) ah.deltas = deltas + size_t(tile0) * 32 * num_iters;
        pgx::View evh = ev;
        if (ev.kind == 1) evh.p += off(Vs);
        float* S_h = ws.S + off(Vs);
        float* part_h = ws.part + off(plan->part_rows);
        const float* c_old_h = c_old ? c_old + off(plan->c_rows) : nullptr;
        if (temperature == 0.f)
          rc = launch_f2v<false>(plan, sh, mph, lp, S_h, cur + off(Es), dst + off(Es), ah, fused, false, evh, nullptr,
                                 c_old_h, c_new + off(plan->c_rows), false, part_h, h);
        else
          rc = launch_f2v<true>(plan, sh, mph, lp, S_h, cur + off(Es), dst + off(Es), ah, fused, false, evh, nullptr,
                                c_old_h, c_new + off(plan->c_rows), false, part_h, h);
        if (rc) return rc;
        if (it + 1 < num_iters) {
          pgx::k_var_reduce<<<grid_for(plan, mph, Vs), pgx::kThreads, 0, sh>>>(
              mph, Vs, Es, plan->part_rows, plan->d_vs_var, plan->d_var_first_state, plan->d_rest_ptr,
              plan->d_rest_edge_msg, plan->d_part_first, plan->d_part_count, evh, dst + off(Es), part_h, S_h);
          if ((rc = check_launch(plan, "k_var_reduce"))) return rc;
        }
      }
      if (it == num_iters - 1) {
        PGX_CUDA(cudaEventRecord(plan->ev_half_join, plan->half));
        PGX_CUDA(cudaStreamWaitEvent(st, plan->ev_half_join, 0));
      }
      nxt = (dst == bufA) ? bufB : bufA;
      cur = dst;
      continue;
    }
    if (temperature == 0.f)
      rc = launch_f2v<false>(plan, st, mp, lp, ws.S, cur, dst, a, fused, lpull, ev, aux, c_old, c_new, lbin, nullptr, 0, gbin);
    else
      rc = launch_f2v<true>(plan, st, mp, lp, ws.S, cur, dst, a, fused, lpull, ev, aux, c_old, c_new, lbin, nullptr, 0, gbin);
    if (rc) return rc;
    if (aux != nullptr) {
      PGX_CUDA(cudaEventRecord(plan->ev_join, aux));
      PGX_CUDA(cudaStreamWaitEvent(st, plan->ev_join, 0));
    }
    if (fused && it + 1 < num_iters) {
      // next iteration's variable sums from the partial sums the fused blocks just wrote
      pgx::k_var_reduce<<<grid_for(plan, mp, Vs), pgx::kThreads, 0, st>>>(
          mp, Vs, Es, plan->part_rows, plan->d_vs_var, plan->d_var_first_state, plan->d_rest_ptr,
          plan->d_rest_edge_msg, plan->d_part_first, plan->d_part_count, ev, dst, ws.part, ws.S);
      if ((rc = check_launch(plan, "k_var_reduce"))) return rc;
    }
    // ping-pong between the two workspace buffers; the caller's input is never written
    nxt = (dst == bufA) ? bufB : bufA;
    cur = dst;
  }
  // PGX_RUN_FINAL_SUMS: the variable sums of the FINAL messages - the beliefs - are left in the
  // workspace for pgx_decode_last_run (batched two-pass / single-pass / pull paths; the others
  // simply do not set final_sums_batch and the caller decodes from the messages)
  plan->final_sums_batch = 0;
  if ((flags & PGX_RUN_FINAL_SUMS) && !single && !lpull && !lattice) {
    if (fused) {
      pgx::k_var_reduce<<<grid_for(plan, mp, Vs), pgx::kThreads, 0, st>>>(
          mp, Vs, Es, plan->part_rows, plan->d_vs_var, plan->d_var_first_state, plan->d_rest_ptr,
          plan->d_rest_edge_msg, plan->d_part_first, plan->d_part_count, ev, cur, ws.part, ws.S);
      if ((rc = check_launch(plan, "k_var_reduce"))) return rc;
    } else if (gbin) {
      pgx::k_var_sums_bin<<<grid_for(plan, mp, plan->num_vars), pgx::kThreads, 0, st>>>(
          mp, plan->num_vars, Es / 2, Vs, plan->d_vs_csr, plan->d_var_edge_msg, ev, cur, ws.S);
      if ((rc = check_launch(plan, "k_var_sums_bin"))) return rc;
    } else {
      pgx::k_var_sums<<<grid_for(plan, mp, Vs), pgx::kThreads, 0, st>>>(mp, Vs, Es, plan->d_vs_csr, plan->d_var_edge_msg,
                                                                       ev, cur, ws.S, 0);
      if ((rc = check_launch(plan, "k_var_sums"))) return rc;
    }
    plan->final_sums_batch = batch;
    if (flags & PGX_RUN_SKIP_OUTPUT) return PGX_OK;  // the caller only wants the decoding
  }
  PGX_CHECK(ftov_out != nullptr, "ftov_out is null");
  if (cbin) {
    dim3 grid((unsigned)((Es / 2 + 31) / 32), (unsigned)((mp.batch + 31) / 32)), block(32, 8);
    pgx::k_expand_bin<<<grid, block, 0, st>>>(cur, Es / 2, 0, Es / 2, ftov_out, Es, 0, mp);
    if ((rc = check_launch(plan, "k_expand_bin"))) return rc;
  } else if (!single && !fused) {
    if ((rc = from_tiles(plan, st, cur, ftov_out, Es, mp))) return rc;
  } else if (!single) {
    // the fused blocks' messages are expanded from the compressed array, the rest comes from
    // the full-layout buffer (message ranges of the blocks are disjoint and ascending)
    const float* c_fin = ((num_iters - 1) & 1) ? ws.cB : ws.cA;
    int64_t done = 0;
    for (const BipPlan* bpp : bips_by_msg(plan)) {
      const BipPlan& bp = *bpp;
      const int64_t count = 2 * int64_t(bp.dev.I) * bp.dev.J;
      if ((rc = from_tiles(plan, st, cur, ftov_out, Es, mp, done, bp.dev.first_msg))) return rc;
      dim3 grid((unsigned)((count + 31) / 32), (unsigned)((mp.batch + 31) / 32)), block(32, 8);
      pgx::k_expand_bin<<<grid, block, 0, st>>>(c_fin, plan->c_rows, bp.dev.first_cmsg, count, ftov_out, Es,
                                                bp.dev.first_msg, mp);
      if ((rc = check_launch(plan, "k_expand_bin"))) return rc;
      done = bp.dev.first_msg + 2 * count;
    }
    if ((rc = from_tiles(plan, st, cur, ftov_out, Es, mp, done, Es))) return rc;
  }
  return PGX_OK;
}

static int decode_impl(pgx_plan* plan, cudaStream_t st, int64_t batch, const float* evidence, int ev_batched,
                       const float* ftov_msgs, int msgs_batched, float* beliefs, int32_t* map_out,
                       float* marginals, int32_t* ties) {
  PGX_CHECK(batch >= 1, "batch must be >= 1");
  PGX_CHECK(plan->num_var_states == 0 || evidence != nullptr, "evidence is null");
  PGX_CHECK(plan->num_edge_states == 0 || ftov_msgs != nullptr, "ftov_msgs is null");
  int rc;
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(plan))) return rc;
  if (ties) PGX_CUDA(cudaMemsetAsync(ties, 0, size_t(batch) * sizeof(int32_t), st));
  if (plan->num_vars == 0) return PGX_OK;
  const pgx::BatchMap mp = make_map(batch);
  // The ABI arrays are read in place through strided views (batch-major).
  pgx::View ev{evidence, plan->num_var_states, ev_batched ? 2 : 0};
  pgx::View m{ftov_msgs, plan->num_edge_states, msgs_batched ? 2 : 0};
  pgx::k_decode<<<grid_for(plan, mp, plan->num_vars), pgx::kThreads, 0, st>>>(
      mp, plan->num_vars, plan->num_var_states, plan->d_var_first_state, plan->d_var_ptr,
      plan->d_var_edge_msg, ev, m, beliefs, map_out, marginals, ties);
  return check_launch(plan, "k_decode");
}

int pgx_beliefs(pgx_plan* plan, void* stream, int64_t batch, const float* evidence, int ev_batched,
                const float* ftov_msgs, int msgs_batched, float* beliefs_out) {
  if (!plan) return fail(PGX_ERR_INVALID, "null plan");
  PGX_CHECK(beliefs_out != nullptr || plan->num_var_states == 0, "beliefs_out is null");
  return decode_impl(plan, static_cast<cudaStream_t>(stream), batch, evidence, ev_batched, ftov_msgs,
                     msgs_batched, beliefs_out, nullptr, nullptr, nullptr);
}

int pgx_decode(pgx_plan* plan, void* stream, int64_t batch, const float* evidence, int ev_batched,
               const float* ftov_msgs, int msgs_batched, int32_t* map_out, float* marginals_out,
               int32_t* tie_count_out) {
  if (!plan) return fail(PGX_ERR_INVALID, "null plan");
  return decode_impl(plan, static_cast<cudaStream_t>(stream), batch, evidence, ev_batched, ftov_msgs,
                     msgs_batched, nullptr, map_out, marginals_out, tie_count_out);
}

int pgx_decode_last_run(pgx_plan* plan, void* stream, int64_t batch, int32_t* map_out, float* marginals_out,
                        int32_t* tie_count_out) {
  if (!plan) return fail(PGX_ERR_INVALID, "null plan");
  PGX_CHECK(plan->final_sums_batch == batch && batch >= 1,
            "the last pgx_bp_run on this plan left no final sums for a batch of %lld (PGX_RUN_FINAL_SUMS)", (long long)batch);
  int rc;
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(plan))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (tie_count_out) PGX_CUDA(cudaMemsetAsync(tie_count_out, 0, size_t(batch) * sizeof(int32_t), st));
  if (plan->num_vars == 0) return PGX_OK;
  const pgx::BatchMap mp = make_map(batch);
  pgx::k_decode_sums<<<grid_for(plan, mp, plan->num_vars), pgx::kThreads, 0, st>>>(
      mp, plan->num_vars, plan->num_var_states, plan->d_var_first_state, plan->ws.S, map_out, marginals_out, tie_count_out);
  return check_launch(plan, "k_decode_sums");
}

int pgx_energy(pgx_plan* plan, void* stream, int64_t batch, const float* log_potentials, int lp_batched,
               const float* evidence, int ev_batched, const int32_t* map_states, int map_batched,
               float* energy_out) {
  if (!plan) return fail(PGX_ERR_INVALID, "null plan");
  PGX_CHECK(batch >= 1 && batch <= 65535, "batch must be in [1, 65535], got %lld", (long long)batch);
  PGX_CHECK(energy_out != nullptr, "energy_out is null");
  PGX_CHECK(plan->num_vars == 0 || map_states != nullptr, "map_states is null");
  PGX_CHECK(plan->num_var_states == 0 || evidence != nullptr, "evidence is null");
  PGX_CHECK(plan->num_potentials == 0 || log_potentials != nullptr, "log_potentials is null");
  int rc;
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(plan))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int slots = 1 + int(plan->enum_blocks.size()) + 3;
  const int64_t need = batch * slots * pgx::kEnergyChunks;
  if (plan->energy_partial_floats < need) {
    free_dev(plan->d_energy_partial);
    plan->d_energy_partial = nullptr;
    PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(&plan->d_energy_partial), size_t(need) * sizeof(float)));
    plan->energy_partial_floats = need;
  }
  float* partial = plan->d_energy_partial;
  PGX_CUDA(cudaMemsetAsync(partial, 0, size_t(need) * sizeof(float), st));  // absent factor types contribute 0
  pgx::EnergyArgs e{};
  e.map = map_states;
  e.map_stride = map_batched ? plan->num_vars : 0;
  e.ev = evidence;
  e.ev_stride = ev_batched ? plan->num_var_states : 0;
  e.lp = log_potentials;
  e.lp_stride = lp_batched ? plan->num_potentials : 0;
  e.var_first_state = plan->d_var_first_state;
  e.vs_var = plan->d_vs_var;
  e.edge_vs = plan->d_edge_vs;
  const dim3 grid(pgx::kEnergyChunks, unsigned(batch));
  if (plan->num_vars > 0) {
    pgx::k_energy_vars<<<grid, pgx::kThreads, 0, st>>>(e, plan->num_vars, partial, slots);
    if ((rc = check_launch(plan, "k_energy_vars"))) return rc;
  }
  int slot = 1;
  for (const EnumBlockPlan& eb : plan->enum_blocks) {
    pgx::k_energy_enum<<<grid, pgx::kThreads, 0, st>>>(e, eb.dev, partial, slots, slot++);
    if ((rc = check_launch(plan, "k_energy_enum"))) return rc;
  }
  for (const LogicalPlan* lg : {&plan->or_f, &plan->and_f}) {
    if (lg->dev.num_factors > 0) {
      pgx::k_energy_logical<false><<<grid, pgx::kThreads, 0, st>>>(e, lg->dev, partial, slots, slot);
      if ((rc = check_launch(plan, "k_energy_logical"))) return rc;
    }
    ++slot;
  }
  if (plan->pool_f.dev.num_factors > 0) {
    pgx::k_energy_logical<true><<<grid, pgx::kThreads, 0, st>>>(e, plan->pool_f.dev, partial, slots, slot);
    if ((rc = check_launch(plan, "k_energy_logical"))) return rc;
  }
  pgx::k_energy_sum<<<unsigned((batch + 255) / 256), 256, 0, st>>>(partial, slots * pgx::kEnergyChunks, batch, energy_out);
  return check_launch(plan, "k_energy_sum");
}

}  // extern "C"

namespace {
template <typename T>
int grow(T** p, int64_t* have, int64_t need) {
  if (need <= *have) return PGX_OK;
  free_dev(*p);
  *p = nullptr;
  *have = 0;
  PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(p), size_t(need) * sizeof(T)));
  *have = need;
  return PGX_OK;
}
}  // namespace

extern "C" {

int pgx_infer_host(pgx_plan* plan, void* stream, int64_t batch, const float* lp_h, int lp_batched,
                   const float* ev_h, int ev_batched, const float* msgs_h, int msgs_batched,
                   int32_t num_iters, float damping, float temperature, int32_t* map_h, float* marg_h,
                   int32_t* ties_h, float* msgs_out_h, float* deltas_h) {
  if (!plan) return fail(PGX_ERR_INVALID, "null plan");
  PGX_CHECK(batch >= 1, "batch must be >= 1");
  int rc;
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(plan))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace& ws = plan->ws;
  const int64_t n_lp = plan->num_potentials * (lp_batched ? batch : 1);
  const int64_t n_ev = plan->num_var_states * (ev_batched ? batch : 1);
  const int64_t n_in = msgs_h ? plan->num_edge_states * (msgs_batched ? batch : 1) : 0;
  const int64_t n_out = plan->num_edge_states * batch;
  if ((rc = grow(&ws.h_lp, &ws.n_lp, std::max<int64_t>(n_lp, 1)))) return rc;
  if ((rc = grow(&ws.h_ev, &ws.n_ev, std::max<int64_t>(n_ev, 1)))) return rc;
  if ((rc = grow(&ws.h_msgs_in, &ws.n_msgs_in, std::max<int64_t>(n_in, 1)))) return rc;
  if ((rc = grow(&ws.h_msgs_out, &ws.n_msgs_out, std::max<int64_t>(n_out, 1)))) return rc;
  if ((rc = grow(&ws.h_map, &ws.n_map, std::max<int64_t>(plan->num_vars * batch, 1)))) return rc;
  if ((rc = grow(&ws.h_ties, &ws.n_ties, batch))) return rc;
  if (marg_h && (rc = grow(&ws.h_marg, &ws.n_marg, std::max<int64_t>(plan->num_var_states * batch, 1))))
    return rc;
  if (deltas_h && (rc = grow(&ws.h_deltas, &ws.n_deltas, batch * num_iters))) return rc;
  if (n_lp) PGX_CUDA(cudaMemcpyAsync(ws.h_lp, lp_h, size_t(n_lp) * 4, cudaMemcpyHostToDevice, st));
  if (n_ev) PGX_CUDA(cudaMemcpyAsync(ws.h_ev, ev_h, size_t(n_ev) * 4, cudaMemcpyHostToDevice, st));
  if (n_in) PGX_CUDA(cudaMemcpyAsync(ws.h_msgs_in, msgs_h, size_t(n_in) * 4, cudaMemcpyHostToDevice, st));
  // the decoding comes from the variable sums of the final messages where the path leaves them
  // (no second read of the messages); the messages themselves are only materialised in the ABI
  // layout when the caller asks for them
  const uint32_t flags = PGX_RUN_FINAL_SUMS | (msgs_out_h ? 0u : PGX_RUN_SKIP_OUTPUT);
  if ((rc = pgx_bp_run_flags(plan, stream, batch, ws.h_lp, lp_batched, ws.h_ev, ev_batched,
                             msgs_h ? ws.h_msgs_in : nullptr, msgs_batched, ws.h_msgs_out,
                             deltas_h ? ws.h_deltas : nullptr, num_iters, damping, temperature, flags)))
    return rc;
  if (map_h || marg_h || ties_h) {
    if (plan->final_sums_batch == batch)
      rc = pgx_decode_last_run(plan, stream, batch, map_h ? ws.h_map : nullptr, marg_h ? ws.h_marg : nullptr,
                               ties_h ? ws.h_ties : nullptr);
    else
      rc = pgx_decode(plan, stream, batch, ws.h_ev, ev_batched, ws.h_msgs_out, 1, map_h ? ws.h_map : nullptr,
                      marg_h ? ws.h_marg : nullptr, ties_h ? ws.h_ties : nullptr);
    if (rc) return rc;
  }
  if (map_h)
    PGX_CUDA(cudaMemcpyAsync(map_h, ws.h_map, size_t(plan->num_vars) * batch * 4, cudaMemcpyDeviceToHost, st));
  if (marg_h)
    PGX_CUDA(cudaMemcpyAsync(marg_h, ws.h_marg, size_t(plan->num_var_states) * batch * 4,
                             cudaMemcpyDeviceToHost, st));
  if (ties_h) PGX_CUDA(cudaMemcpyAsync(ties_h, ws.h_ties, size_t(batch) * 4, cudaMemcpyDeviceToHost, st));
  if (msgs_out_h)
    PGX_CUDA(cudaMemcpyAsync(msgs_out_h, ws.h_msgs_out, size_t(n_out) * 4, cudaMemcpyDeviceToHost, st));
  if (deltas_h)
    PGX_CUDA(cudaMemcpyAsync(deltas_h, ws.h_deltas, size_t(batch) * num_iters * 4, cudaMemcpyDeviceToHost, st));
  PGX_CUDA(cudaStreamSynchronize(st));
  return PGX_OK;
}

}  // extern "C"

#include "pgx_sdlp.cuh"
#include "pgx_strip.cuh"
#include "pgx_vjp.cuh"
