/*
 * pgx.h — C ABI of the B200-native loopy belief propagation engine.
 *
 * This is the drop-in boundary for PGMax's inference path.  Every entry point
 * replaces one piece of the reference's Python/JAX surface (citations are
 * relative to the reference tree, pgmax 0.6.1):
 *
 *   pgx_plan_create      InfererContext.__post_init__           pgmax/infer/inferer.py:65-118
 *                        (+ the Wiring arrays it concatenates:  pgmax/factor/factor.py:132-192,
 *                         enum.py:54-72, logical.py:76-104, pool.py:62-83)
 *   pgx_bp_run           BP.run / run_with_diffs / run_bp       pgmax/infer/bp.py:63-176
 *                        (pass_var_to_fac_messages :191-219, FAC_TO_VAR_UPDATES loop :107-123,
 *                         damping + normalize_and_clip_msgs :127-133,222-260, msgs_delta :136;
 *                         per-type updates enum.py:398-475, logical.py:492-779, pool.py:276-474)
 *   pgx_beliefs          InfererContext.get_beliefs             pgmax/infer/inferer.py:211-225
 *   pgx_decode           decode_map_states + get_marginals      pgmax/infer/inferer.py:251-264,
 *                                                               pgmax/infer/bp.py:263-288
 *   pgx_infer_host       the benchmark's timed region           benchmark/rbm_lib.py:173-187
 *                        (init -> run -> get_beliefs -> decode_map_states) with HOST buffers
 *
 * Shape of the ABI: XLA-FFI-like — plain device pointers, sizes and scalar
 * attributes plus a stream; stateless apart from the immutable plan handle.
 * No torch / jax types appear in any signature.  Functions return 0 on
 * success and a negative pgx_status otherwise, never throw, never exit; the
 * message of the last error on the calling thread is pgx_last_error().
 *
 * Data contract (identical to BPArrays, pgmax/infer/bp_state.py:30-55):
 *   log_potentials [C] or [B, C], ftov_msgs [E_s] or [B, E_s],
 *   evidence [V_s] or [B, V_s]; contiguous fp32, batch-major.  The leading
 *   batch axis replaces jax.vmap.  Inputs are never written.
 */
#ifndef PGX_H_
#define PGX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGX_VERSION 1

/* Numerics policy, reference pgmax/utils/__init__.py:26-37 and logical.py:33. */
#define PGX_MSG_NEG_INF (-1e32f)
#define PGX_LOG_POTENTIAL_MAX_ABS (1e6f)
#define PGX_TEMPERATURE_STABILITY_THRE (0.5f)

typedef enum pgx_status {
  PGX_OK = 0,
  PGX_ERR_INVALID = -1,     /* bad argument / inconsistent graph description */
  PGX_ERR_CUDA = -2,        /* a CUDA runtime call failed */
  PGX_ERR_UNSUPPORTED = -3, /* valid graph the kernels do not cover yet */
  PGX_ERR_NO_DEVICE = -4    /* no CUDA device visible: there is no CPU fallback */
} pgx_status;

/* One group of EnumFactors sharing a configuration table
 * (pgmax/fgroup/enum.py:30-95; replaces the reference's expanded
 * factor_configs_edge_states rows, pgmax/factor/enum.py:364-394).
 * Factor f (0 <= f < num_factors) of the block owns edges
 * [first_edge + f*arity, +arity), whose message slots are contiguous from
 * first_msg on, and potentials [first_potential + f*num_configs, +num_configs). */
typedef struct pgx_enum_block {
  int64_t num_factors;
  int32_t arity;
  int32_t num_configs;
  const int32_t* configs; /* host, [num_configs * arity]: state of variable a in configuration k */
  int64_t first_edge;
  int64_t first_msg;
  int64_t first_potential;
} pgx_enum_block;

/* All OR factors, all AND factors, or all Pool factors of the graph
 * (LogicalWiring pgmax/factor/logical.py:39-104, PoolWiring pool.py:34-83).
 * Message indices are GLOBAL (slice start already added).  parents_factor is
 * ascending and every factor has >= 1 parent, as in every wiring the
 * reference compiles (logical.py:472-483). */
typedef struct pgx_logical_desc {
  int64_t num_factors;
  int64_t num_parents;
  const int32_t* parents_factor; /* host, [num_parents] */
  const int32_t* parents_msg;    /* host, [num_parents]: state 0 (OR, Pool) / state 1 (AND) */
  const int32_t* children_msg;   /* host, [num_factors] */
  int32_t edge_states_offset;    /* +1 OR and Pool, -1 AND */
} pgx_logical_desc;

/* Flat layout of a compiled factor graph.  An "edge" is a
 * (factor, variable) pair; its states are contiguous in the message vector and
 * in the evidence vector, so the incidence is given per edge. */
typedef struct pgx_graph_desc {
  int64_t num_vars;
  const int32_t* var_num_states;  /* host, [num_vars]; sum = V_s */
  int64_t num_edges;
  const int32_t* edge_var_start;  /* host, [num_edges]: var-state index of the edge's state 0 */
  const int32_t* edge_num_states; /* host, [num_edges]; sum = E_s */
  int64_t num_potentials;         /* C */
  int32_t num_enum_blocks;
  const pgx_enum_block* enum_blocks;
  pgx_logical_desc or_factors;
  pgx_logical_desc and_factors;
  pgx_logical_desc pool_factors;
} pgx_graph_desc;

typedef struct pgx_plan pgx_plan;

/* Sizes a caller needs to allocate buffers. */
typedef struct pgx_plan_info {
  int64_t num_vars, num_var_states, num_edges, num_edge_states, num_potentials;
  int64_t max_var_states;       /* largest num_states of any variable */
  int64_t device_bytes;         /* index structures resident on the device */
  int32_t device;               /* CUDA device the plan lives on */
  int32_t num_sms;
} pgx_plan_info;

/* Builds the device-resident plan (index structures incl. the variable->edges
 * CSR the reference never builds) on the current CUDA device.  The
 * description is copied; the caller may free it on return. */
int pgx_plan_create(const pgx_graph_desc* desc, pgx_plan** out_plan);
void pgx_plan_destroy(pgx_plan* plan);
int pgx_plan_get_info(const pgx_plan* plan, pgx_plan_info* out_info);

/* num_iters iterations of damped loopy BP, enqueued on `stream` (a
 * cudaStream_t; NULL = legacy default stream).  All pointers are DEVICE
 * pointers.  batch >= 1.  lp_batched / ev_batched / msgs_batched say whether
 * the corresponding input carries the leading batch axis; ftov_out is
 * [batch, E_s].  ftov_in may be NULL (zero messages, as bp.init gives).
 * deltas may be NULL, else [batch, num_iters] receives max|m' - m| per
 * iteration (run_with_diffs).  temperature == 0 selects max-product.
 * ftov_out may alias ftov_in.  No host synchronisation is performed unless the
 * plan's workspace has to grow (first call for a given batch size). */
int pgx_bp_run(pgx_plan* plan, void* stream, int64_t batch,
               const float* log_potentials, int lp_batched,
               const float* evidence, int ev_batched,
               const float* ftov_in, int msgs_batched,
               float* ftov_out, float* deltas,
               int32_t num_iters, float damping, float temperature);

/* pgx_bp_run with flags.  PGX_RUN_INPUT_NORMALIZED: the caller guarantees that
 * ftov_in is the output of a previous run (every edge already has max 0 and is
 * clipped), so the initial normalisation pass (bp.py:92-96, idempotent) is
 * skipped; for batch == 1 and ftov_in != ftov_out the input is then read in
 * place with no staging copy.  This is the step function of callers that must
 * touch the evidence between iterations (multi-GPU halo exchange, dist.py). */
#define PGX_RUN_INPUT_NORMALIZED 1u
/* One launch per run (the reference's loop is one device-side lax.scan, bp.py:142-146): the
 * SECOND call with the same signature - buffers, batch, iteration count, damping, temperature,
 * flags, path mask - captures the launch sequence of all iterations in a CUDA graph, and that
 * call and every later one with the signature is a single cudaGraphLaunch on `stream`.  (The
 * first call runs directly: it sizes the workspace, and a signature that never repeats never pays
 * for a capture.)  Skipped when `stream` is itself being captured (the caller's graph then holds
 * the launches), under pgx_plan_profile_enable, with PGX_GRAPH=0 in the environment, or with: */
#define PGX_RUN_NO_GRAPH 2u
/* The caller guarantees that log_potentials (same pointer) holds the same values as in this
 * plan's previous run: per-run preprocessing of the potentials (the round-ordered copy of the
 * merged max-product launch) is reused instead of redone. */
#define PGX_RUN_POTENTIALS_UNCHANGED 4u
/* Leave the variable sums of the FINAL messages (= the beliefs) in the plan's workspace for
 * pgx_decode_last_run: the fused decode then reads no message at all.  Honoured by the batched
 * two-pass / single-pass / pull paths (pgx_decode_last_run fails for the others: decode from the
 * messages then).  With PGX_RUN_SKIP_OUTPUT a run that left the sums does not write ftov_out. */
#define PGX_RUN_FINAL_SUMS 8u
#define PGX_RUN_SKIP_OUTPUT 16u
int pgx_bp_run_flags(pgx_plan* plan, void* stream, int64_t batch,
                     const float* log_potentials, int lp_batched,
                     const float* evidence, int ev_batched,
                     const float* ftov_in, int msgs_batched,
                     float* ftov_out, float* deltas,
                     int32_t num_iters, float damping, float temperature, uint32_t flags);

/* beliefs_out[batch, V_s] = evidence + sum of incoming messages. */
int pgx_beliefs(pgx_plan* plan, void* stream, int64_t batch,
                const float* evidence, int ev_batched,
                const float* ftov_msgs, int msgs_batched,
                float* beliefs_out);

/* Fused beliefs + MAP decode (+ marginals).  map_out[batch, num_vars] int32 =
 * first arg-max state of every variable; marginals_out (may be NULL)
 * [batch, V_s] = softmax of the beliefs per variable; tie_count_out (may be
 * NULL) [batch] int32 = number of variables whose two largest beliefs are
 * exactly equal (the north-star's tie rule). */
int pgx_decode(pgx_plan* plan, void* stream, int64_t batch,
               const float* evidence, int ev_batched,
               const float* ftov_msgs, int msgs_batched,
               int32_t* map_out, float* marginals_out, int32_t* tie_count_out);

/* pgx_decode on the sums the last pgx_bp_run_flags(..., PGX_RUN_FINAL_SUMS) of this plan left behind
 * (same outputs and rules as pgx_decode; PGX_ERR_INVALID if that run left none for `batch`). */
int pgx_decode_last_run(pgx_plan* plan, void* stream, int64_t batch,
                        int32_t* map_out, float* marginals_out, int32_t* tie_count_out);

/* Energy of a decoding (replaces infer.compute_energy, pgmax/infer/energy.py:53-148, and the
 * per-type compute_energy of pgmax/factor/enum.py:276-323, logical.py:295-358, pool.py:184-239):
 *   energy = - sum_v evidence[v, state(v)] - sum over EnumFactors of the log potential of the
 *   configuration the decoding selects; +inf if an EnumFactor has no valid configuration for
 *   the decoding or an OR / AND / Pool constraint is violated.  Potentials are NOT clipped (as
 *   in the reference).  map_states: [batch][num_vars] int32 device array in the flat variable
 *   order (pgx_decode's map_out), or one shared row (map_batched = 0); energy_out: [batch]
 *   float32.  Sums are formed in a fixed order (deterministic; no float atomics). */
int pgx_energy(pgx_plan* plan, void* stream, int64_t batch,
               const float* log_potentials, int lp_batched,
               const float* evidence, int ev_batched,
               const int32_t* map_states, int map_batched, float* energy_out);

/* ---- Smooth dual LP-MAP (replaces SDLP, pgmax/infer/dual_lp.py:60-463) ------------------
 *
 * Factor membership of the edges (the reference's factor_indices_for_edge_states,
 * pgmax/infer/inferer.py:76-98; edges of one factor are contiguous in every wiring the
 * reference compiles): edges [factor_edge_start[f], factor_edge_start[f + 1]) belong to
 * factor f.  host array, [num_factors + 1].  Required once before any pgx_sdlp_* call; the
 * BP entry points do not need it. */
int pgx_plan_set_factors(pgx_plan* plan, int64_t num_factors, const int32_t* factor_edge_start);

/* One evaluation of smooth_dual_objval_and_grad (dual_lp.py:67-237) at the dual messages
 * ftov_msgs: the BP updates on vtof = -ftov_msgs with normalize=False and the UNCLIPPED
 * potentials, per-variable and per-edge softmax / logsumexp at logsumexp_temp (arg-max / max
 * with ties to the largest index for logsumexp_temp == 0, update_utils.py:26-64,102-131).
 *   objval_out     [batch]             sum_v L_v + sum_f max over the factor's edges of L_e
 *   grad_out       [batch, E_s] / NULL (sub)gradient with respect to the dual messages
 *   bp_updates_out [batch, E_s] / NULL the BP updates   (get_bp_updates, dual_lp.py:411-445)
 *   edge_vals_out  [batch, num_edges] / NULL  L_e: the per-edge logsumexp (max)
 * get_primal_upper_bound (dual_lp.py:366-378) is this call at logsumexp_temp == 0. */
int pgx_sdlp_objval_and_grad(pgx_plan* plan, void* stream, int64_t batch,
                             const float* log_potentials, int lp_batched,
                             const float* evidence, int ev_batched,
                             const float* ftov_msgs, int msgs_batched, float logsumexp_temp,
                             float* objval_out, float* grad_out, float* bp_updates_out,
                             float* edge_vals_out);

/* run_with_objvals (dual_lp.py:239-324): num_iters steps of accelerated gradient descent
 * (logsumexp_temp > 0) or subgradient descent (== 0) on the dual messages, all enqueued on
 * `stream` with no host synchronisation:
 *   eta' = m - steps[it] * grad;   m' = eta' + momenta[it] * (eta' - eta);   eta starts as m.
 * steps / momenta are HOST arrays [num_iters] holding the reference's fp32 scalars
 * (lr or lr / sqrt(it + 1), and (it + 1) / (it + 4); dual_lp.py:293-309) - the caller owns the
 * learning-rate policy and its argument checks, as the Python reference does.
 * ftov_in may be NULL (zeros).  objvals (may be NULL) [batch, num_iters] receives the
 * objective BEFORE each step.  ftov_out [batch, E_s] may alias ftov_in. */
int pgx_sdlp_run(pgx_plan* plan, void* stream, int64_t batch,
                 const float* log_potentials, int lp_batched,
                 const float* evidence, int ev_batched,
                 const float* ftov_in, int msgs_batched, float* ftov_out, float* objvals,
                 int32_t num_iters, const float* steps, const float* momenta,
                 float logsumexp_temp);

/* End-to-end call with HOST buffers: H2D copies, pgx_bp_run, pgx_decode, D2H
 * copies, then a stream synchronise.  ftov_in_host may be NULL (zeros);
 * ftov_out_host, marginals_out_host, tie_count_out_host, deltas_out_host may be
 * NULL (not copied back).  Host buffers should be page-locked for the copies
 * to overlap; pageable memory works but is slower. */
int pgx_infer_host(pgx_plan* plan, void* stream, int64_t batch,
                   const float* log_potentials_host, int lp_batched,
                   const float* evidence_host, int ev_batched,
                   const float* ftov_in_host, int msgs_batched,
                   int32_t num_iters, float damping, float temperature,
                   int32_t* map_out_host, float* marginals_out_host,
                   int32_t* tie_count_out_host, float* ftov_out_host,
                   float* deltas_out_host);

/* Number of kernels this library launched on behalf of `plan` since creation
 * (bench.py reports it as gpu_launches). */
int64_t pgx_plan_launch_count(const pgx_plan* plan);

/* Summation order of the variable sums S_v = ev_v + sum of incoming messages.
 * Default (0): blocks of pairwise binary factors that form a dense rows x columns
 * grid (RBM-like) are updated in ONE pass per iteration when batch > 16; their
 * contribution to S_v is then added as per-tile partial sums in a fixed tree
 * order - deterministic, but not the serial ascending order of a CPU
 * scatter-add (the reference's own GPU scatter-add is unordered atomics).
 * 1: always use the two-pass path that accumulates S_v serially in ascending
 * message index, bit-compatible with the CPU oracle for max-product.  The
 * environment variable PGX_EXACT_ORDER=1 sets the default for new plans. */
int pgx_plan_set_exact_order(pgx_plan* plan, int enabled);
/* Number of enum blocks for which the single-pass path is available. */
int pgx_plan_num_fused_blocks(const pgx_plan* plan);
/* Edges (per sample) whose messages the single-pass path keeps in binary-difference
 * storage between iterations: a normalised message of a two-state edge is (n0, n1) with
 * max(n0, n1) == 0 exactly, so the one float x = n1 - n0 carries both states without loss
 * (n0 = min(-x, 0), n1 = min(x, 0)).  The ABI arrays keep the reference's layout; only the
 * workspace between iterations is compressed (half the message traffic, bit-identical
 * values).  bench.py uses the count for the bytes the kernels actually have to move. */
int64_t pgx_plan_compressed_edges(const pgx_plan* plan);

/* Specialised launch paths pgx_bp_run may choose from the graph's structure.  The mask only
 * exists so that tests and profiles can pin a path.  Default: nothing disabled.  Max-product
 * (temperature 0): every path performs the same operations in the same order - bit-identical
 * messages - except the single-pass dense-grid path (tiled partial sums, see
 * pgx_plan_set_exact_order).  Sum-product: additionally PGX_PATH_ENUM_CONFIG_MAJOR (and the two
 * paths below it) evaluates exp as ex2 on the hardware unit, and the T > 0 side of
 * PGX_PATH_MERGED_MAX forms logsumexp in one pass with running maxima, adding the terms in
 * the order of its round schedule (deterministic, not the ascending configuration order); both
 * stay within the sum-product tolerance (1e-5) of the serial fp32 evaluation (tests/).
 *   PGX_PATH_LATTICE   index-free stencil kernel for a graph that is one 2-D lattice of
 *                      pairwise binary factors (Ising grids, batch == 1)
 *   PGX_PATH_RESIDENT  all iterations of a small pairwise graph in one cluster /
 *                      cooperative launch
 *   PGX_PATH_PULL      pairwise kernels that re-derive the variable sums per factor
 *                      instead of materialising them
 *   PGX_PATH_MERGED_MAX  all large pairwise max-product groups (RCN) in one balanced launch */
#define PGX_PATH_LATTICE 1u
#define PGX_PATH_RESIDENT 2u
#define PGX_PATH_PULL 4u
#define PGX_PATH_MERGED_MAX 8u /* one launch for all large sorted pairwise groups (k_enum_big_maxprod_all; T > 0 on round-ordered potentials: k_enum_big_sumprod_all) */
#define PGX_PATH_LATTICE_STREAM 32u /* persistent TMA variant of the lattice kernel (large lattices) */
#define PGX_PATH_AUX_STREAM 64u /* the smaller of the OR / AND groups on an auxiliary stream */
#define PGX_PATH_WIDE_SPLIT 128u /* wide OR / AND update as a serial reduce launch + a parent-parallel emit launch */
#define PGX_PATH_LOGICAL_PULL 16u /* OR / AND kernels that re-derive the sums of variables with <= 2 edges */
#define PGX_PATH_PERM_POTENTIALS 512u /* merged max-product launch reads a round-ordered copy of the potentials */
#define PGX_PATH_HALF_BATCH 1024u /* single-pass mode: the two halves of a batch of >= 16 sample tiles as two pipelined chains on two streams */
#define PGX_PATH_STAGED_WIRING 2048u /* uniform OR / AND groups: wiring of a CTA's factor range staged in shared memory */
#define PGX_PATH_TAIL_SPLIT 4096u /* OR / AND graphs: a short batch tail beyond the last full tile of 32 samples runs beside the full tiles (fused path: packed k_or_and_fused launch, <= 16 samples; separate kernels: second plan, <= 8 samples) */
#define PGX_PATH_ORAND_FUSED 16384u /* OR factors over degree-2 children of two-parent AND factors: one fused launch (k_or_and_fused) */
#define PGX_PATH_VARSUM_COOP 32768u /* sums of very high-degree variables: a CTA per variable gathers through shared memory */
#define PGX_PATH_ENUM_UNARY 65536u /* one-variable EnumFactors over all states: closed form (k_enum_unary) instead of k_enum_small */
#define PGX_PATH_ENUM_CONFIG_MAJOR 131072u /* small EnumFactors (<= 32 edge-states): configuration-major walk (k_enum_small_cm) instead of k_enum_small */
#define PGX_PATH_ENUM_DENSE_PAIR 524288u /* complete n0 x n1 pairwise tables (<= 32 edge-states): nested-loop kernel k_enum_pair_dense instead of k_enum_small_cm */
#define PGX_PATH_ENUM_PAIR_FEW 1048576u /* complete pairwise tables with a 2 ... 4-state side: that side in registers (k_enum_pair_few) instead of k_enum_pair_dense */
#define PGX_PATH_GENERIC_BIN 262144u /* all-binary pairwise graphs on the generic two-pass path: binary-difference storage (k_var_sums_bin + k_enum_pw2_bin) */
#define PGX_PATH_LATTICE_BIN 8192u /* large single-sample lattices on binary-difference storage (k_lattice_bin) */
#define PGX_PATH_LOGICAL_BIN 256u /* ... with the messages in binary-difference storage (one float per edge) */
int pgx_plan_disable_paths(pgx_plan* plan, uint32_t mask);
/* 1 if the lattice path is available for this plan. */
int pgx_plan_is_lattice(const pgx_plan* plan);
/* CUDA-graph replay of repeated runs (see PGX_RUN_NO_GRAPH): on by default. */
int pgx_plan_enable_graphs(pgx_plan* plan, int enabled);
int64_t pgx_plan_graph_launch_count(const pgx_plan* plan); /* cudaGraphLaunch calls so far */

/* Device-time instrumentation for the roofline figure (bench.py).  While
 * enabled, pgx_bp_run brackets, in every iteration, the launch of the plan's
 * dominant kernel (the factor->variable kernel that covers the most
 * edge-states) with CUDA events recorded on the run's stream.
 * pgx_plan_profile_read synchronises those events, returns the number of
 * bracketed launches and their summed device time since the last read, the
 * kernel's name, and resets the counters.  Off by default (no events). */
int pgx_plan_profile_enable(pgx_plan* plan, int enabled);
/* Edge-states (per sample) the dominant launch updates: the units of its roofline figure. */
int64_t pgx_plan_dominant_edge_states(const pgx_plan* plan);
/* CTAs of the dominant kernel's most recent launch (0 before the first run). */
int64_t pgx_plan_dominant_grid(const pgx_plan* plan);
int pgx_plan_profile_read(pgx_plan* plan, int64_t* num_launches, double* total_ms,
                          const char** kernel_name);

/* Message of the last failure on the calling thread ("" if none). */
const char* pgx_last_error(void);

/* Library / build information: "pgx <version> sm_100a ..." */
const char* pgx_build_info(void);

/* ---- Reverse mode through bp.run ---------------------------------------------------------------
 *
 * What jax.grad gives the reference (pgmax/infer/bp.py:98 @jax.checkpoint on the update;
 * examples/grid_mrf.ipynb cells 15-16: value_and_grad of a loss of the marginals with respect to
 * the log potentials): the vector-Jacobian product of
 *     ftov_out = run(log_potentials, evidence, ftov_in; num_iters, damping, temperature)
 * at the cotangent g_ftov_out [batch, E_s].  Sum-product only (temperature > 0), graphs of
 * EnumFactors with at most 64 edge-states per factor; PGX_ERR_UNSUPPORTED otherwise.  Outputs (any
 * may be NULL): g_lp_out [batch, C] if lp_batched else [C] (summed over the batch), g_ev_out
 * likewise with V_s, g_ftov_in_out likewise with E_s.  The iterations are re-run keeping every
 * iterate ((num_iters + 1) x batch x E_s floats of scratch); synchronises `stream` before
 * returning. */
int pgx_bp_run_vjp(pgx_plan* plan, void* stream, int64_t batch,
                   const float* log_potentials, int lp_batched,
                   const float* evidence, int ev_batched,
                   const float* ftov_in, int msgs_batched,
                   const float* g_ftov_out, int32_t num_iters, float damping, float temperature,
                   float* g_lp_out, float* g_ev_out, float* g_ftov_in_out);

/* ---- Row strips of ONE 2-D lattice across the GPUs of a box ----------------------------------
 *
 * BASELINE.json configs[4] / SURVEY.md 8(e) row 2 (the reference has no multi-GPU path; the graph
 * is the torus examples/ising_model.ipynb cell 12 builds, scaled): one process per GPU, rank g owns
 * `rows` consecutive rows of the n_rows x n_cols torus and their factors (variable (i, j) owns the
 * vertical factor to (i + 1, j) and the horizontal factor to (i, j + 1)); each iteration exchanges
 * the boundary messages with NCCL send/recv (ring) on a side stream while the interior rows are
 * updated; all iterations of a run are ONE CUDA graph launch.  world == 1: the whole torus, no NCCL.
 *
 * Arrays (device, fp32, the reference's flat layout restricted to the strip):
 *   log_potentials [8 * rows * n_cols]  4 per factor, factor 2 * (l * n_cols + j) + t (t = 0 vertical)
 *   evidence_own   [2 * rows * n_cols]
 *   msgs_in / out  [8 * rows * n_cols]  (msgs_in may be NULL: zero messages; normalised on entry)
 * The NCCL communicator is created from a 128-byte unique id the caller obtains on rank 0
 * (pgx_nccl_unique_id) and broadcasts with whatever it has (torch.distributed in dist.py). */
typedef struct pgx_strip pgx_strip;
#define PGX_NCCL_ID_BYTES 128
#define PGX_STRIP_NO_GRAPH 1u   /* enqueue the launches directly instead of replaying a CUDA graph */
#define PGX_STRIP_NO_OVERLAP 2u /* wait for the halo before any row is updated (A/B of the overlap) */
int pgx_nccl_unique_id(void* id_out);
int pgx_strip_create(int64_t n_cols, int64_t rows, int rank, int world, const void* nccl_id,
                     pgx_strip** out_strip);
void pgx_strip_destroy(pgx_strip* strip);
int pgx_strip_run(pgx_strip* strip, void* stream, const float* log_potentials,
                  const float* evidence_own, const float* msgs_in, float* msgs_out,
                  int32_t num_iters, float damping, float temperature, uint32_t flags);
/* beliefs_out [2 * rows * n_cols] = evidence + incoming messages of the owned variables (one more
 * boundary exchange). */
int pgx_strip_beliefs(pgx_strip* strip, void* stream, const float* evidence_own, const float* msgs,
                      float* beliefs_out);
int64_t pgx_strip_launch_count(const pgx_strip* strip);       /* kernels + exchanges enqueued so far */
int64_t pgx_strip_graph_launch_count(const pgx_strip* strip); /* cudaGraphLaunch calls so far */

#ifdef __cplusplus
}
#endif
#endif /* PGX_H_ */
