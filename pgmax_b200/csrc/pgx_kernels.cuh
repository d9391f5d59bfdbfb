// pgx_kernels.cuh — sm_100a kernels of the loopy-BP hot path.
//
// Internal data layout ("tile-blocked batch-inner"): every per-sample vector
// x[b][n] of the ABI (batch-major, as jax.vmap produces) is held as
// X[tile][n][TW]: samples are grouped in tiles of TW = min(32, pow2ceil(B))
// consecutive samples; inside a tile the sample index is fastest and element n
// of the tile starts at float offset (tile * N + n) * TW.  A warp covers the TW
// samples of 32 / TW graph elements, so
//   * every index load (incidence, wiring, config tables) is warp-uniform per
//     element, every message / evidence / var-sum access is a full coalesced
//     row, whatever the graph's structure;
//   * consecutive elements of one tile are CONTIGUOUS in memory (128 B apart
//     for TW = 32): a factor's edge-states, and a run of factors, form one
//     contiguous span that a single TMA bulk copy can move, and all address
//     arithmetic is "lane pointer + (n << log2 TW)".
// For batch == 1 this degenerates to the reference's flat vectors, one graph
// element per lane.
//
// Arithmetic follows SURVEY.md App. A operation by operation (same order of
// additions, expf/logf/log1pf/expm1f, no FMA contraction: the library is compiled
// with -fmad=false) so that max-product results are bit-comparable with the CPU
// oracle.  The one deliberate exception is the sum-product path of the
// pairwise-binary kernels (see lse2 below).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace pgx {

constexpr float kMsgNegInf = -1e32f;   // pgmax/utils/__init__.py:26
constexpr float kLpMaxAbs = 1e6f;      // pgmax/utils/__init__.py:32
constexpr float kTempStabThre = 0.5f;  // pgmax/factor/logical.py:33
constexpr float kLn2 = 0.69314718055994530942f;
constexpr int kThreads = 256;
constexpr int kSmallMaxNS = 64;        // enum "small" kernel: edge-states per factor

// How threads map onto (graph element, sample) pairs.
struct BatchMap {
  int batch;   // B
  int bx_log;  // log2 TW, TW = samples per tile = samples covered side by side in a warp
  int nbt;     // number of tiles
};

// Offset (in floats) of (element 0, sample b) in a tile-blocked array of n_rows elements;
// element n of that sample is at  off + (n << bx_log).
__device__ __forceinline__ int64_t lane_off(const BatchMap& mp, int64_t n_rows, int b) {
  const int bt = b >> mp.bx_log, bl = b & ((1 << mp.bx_log) - 1);
  return ((int64_t(bt) * n_rows) << mp.bx_log) + bl;
}

// A per-sample vector as the kernels see it.  kind 0: shared by the whole batch,
// read in place (x[n]); kind 1: tile-blocked workspace array; kind 2: the ABI's
// batch-major array read in place (x[b * n_rows + n]).
struct View {
  const float* p;
  int64_t n_rows;
  int kind;
};

// The view of ONE sample: element n is q[n << sh].
struct LaneView {
  const float* q;
  int sh;
  __device__ __forceinline__ float at(int64_t n) const { return q[n << sh]; }
};

__device__ __forceinline__ LaneView lane_view(const View& v, const BatchMap& mp, int b) {
  if (v.kind == 1) return LaneView{v.p + lane_off(mp, v.n_rows, b), mp.bx_log};
  if (v.kind == 2) return LaneView{v.p + int64_t(b) * v.n_rows, 0};
  return LaneView{v.p, 0};
}

struct UnitLoop {
  int b;
  bool b_ok;
  int64_t u, u_end;
  int64_t step;
};

// Splits `num_units` graph elements over the grid.  blockIdx.y is the sample tile;
// within a tile the warps sweep the elements together (grid-stride): at any moment the
// whole grid works on one contiguous window of the tile's arrays, so neighbouring
// elements' data (gathers into adjacent grid rows, shared index entries) is still in
// L2 when it is needed again, and re-reads of the (small) index arrays by the other
// tiles hit L2.
__device__ __forceinline__ UnitLoop unit_loop(const BatchMap& mp, int64_t num_units) {
  UnitLoop L;
  const int lane = threadIdx.x & 31;
  const int64_t gwarp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int bx = 1 << mp.bx_log;
  const int upw = 32 >> mp.bx_log;  // elements handled side by side in one warp
  L.b = blockIdx.y * bx + (lane & (bx - 1));
  L.b_ok = L.b < mp.batch;
  L.u = gwarp * upw + (lane >> mp.bx_log);
  L.u_end = num_units;
  L.step = nwarps * upw;
  return L;
}

__device__ __forceinline__ float clip_lp(float x) {
  return fminf(fmaxf(x, -kLpMaxAbs), kLpMaxAbs);  // pgmax/infer/bp.py:85-87
}

// max|m' - m| of one sample, accumulated with an integer atomicMax (valid for
// non-negative floats; NaNs are skipped).
__device__ __forceinline__ void publish_delta(float* deltas, int64_t idx, float d) {
  if (deltas != nullptr && d > 0.f) atomicMax(reinterpret_cast<int*>(deltas) + idx, __float_as_int(d));
}

// ---------------------------------------------------------------------------
// update_utils.py restated (pgmax/factor/update_utils.py:135-190)
// ---------------------------------------------------------------------------
__device__ __forceinline__ float logaddexp_t(float x, float y, float T) {
  const float mx = fmaxf(x, y), mn = fminf(x, y);
  return T * log1pf(expf((mn - mx) / T)) + mx;
}
__device__ __forceinline__ float log1mexp(float u) {
  return (u <= kLn2) ? logf(-expm1f(-u)) : log1pf(-expf(-u));
}
__device__ __forceinline__ float logminusexp_t(float x, float y, float T, float eps) {
  return (x >= y + eps) ? (T * log1mexp((x - y) / T) + x) : -INFINITY;
}

// ---------------------------------------------------------------------------
// Layout conversion: ABI batch-major [B][N]  <->  tile-blocked [tile][N][TW]
// ---------------------------------------------------------------------------
__global__ void k_to_tiles(const float* __restrict__ src, float* __restrict__ dst, int64_t N,
                           BatchMap mp) {
  __shared__ float tile[32][33];
  const int64_t n0 = int64_t(blockIdx.x) * 32;
  const int b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int b = b0 + r;
    const int64_t n = n0 + threadIdx.x;
    tile[r][threadIdx.x] = (b < mp.batch && n < N) ? src[int64_t(b) * N + n] : 0.f;
  }
  __syncthreads();
  const int tw = 1 << mp.bx_log;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t n = n0 + r;
    const int b = b0 + threadIdx.x;
    if (n < N && b < mp.nbt * tw) dst[lane_off(mp, N, b) + (n << mp.bx_log)] = tile[threadIdx.x][r];
  }
}

__global__ void k_from_tiles(const float* __restrict__ src, float* __restrict__ dst, int64_t N,
                             int64_t n_begin, int64_t n_end, BatchMap mp) {
  // rows [n_begin, n_end) of the N-row arrays
  __shared__ float tile[32][33];
  const int64_t n0 = n_begin + int64_t(blockIdx.x) * 32;
  const int b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t n = n0 + r;
    const int b = b0 + threadIdx.x;
    tile[r][threadIdx.x] = (n < n_end && b < mp.batch) ? src[lane_off(mp, N, b) + (n << mp.bx_log)] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int b = b0 + r;
    const int64_t n = n0 + threadIdx.x;
    if (b < mp.batch && n < n_end) dst[int64_t(b) * N + n] = tile[threadIdx.x][r];
  }
}

// Compressed binary messages (one float per edge, rows [c_begin, c_begin + count) of the
// c_rows-row tile-blocked array) -> the ABI's batch-major array: edge c -> message rows
// first_msg + 2c (pointed state), + 2c + 1.  Full sample tiles only (bx_log == 5).
__global__ void k_expand_bin(const float* __restrict__ src, int64_t c_rows, int64_t c_begin, int64_t count,
                             float* __restrict__ dst, int64_t N, int64_t first_msg, BatchMap mp) {
  __shared__ float tile[32][33];
  const int64_t c0 = int64_t(blockIdx.x) * 32;
  const int b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t c = c0 + r;
    const int b = b0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < count && b < mp.batch) ? src[lane_off(mp, c_rows, b) + ((c_begin + c) << 5)] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int b = b0 + r;
    if (b >= mp.batch) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int e = h * 32 + threadIdx.x;  // element of the 64 output floats of this tile
      const int64_t c = c0 + (e >> 1);
      if (c < count) {
        const float x = tile[e >> 1][r];
        dst[int64_t(b) * N + first_msg + 2 * c0 + e] = (x != x) ? kMsgNegInf : ((e & 1) ? fminf(x, 0.f) : fminf(-x, 0.f));
      }
    }
  }
}

// Broadcast a shared [N] vector into every sample of a tile-blocked array.
__global__ void k_broadcast_rows(const float* __restrict__ src, float* __restrict__ dst,
                                 int64_t N, BatchMap mp) {
  const int64_t per_tile = N << mp.bx_log;
  const int64_t total = per_tile * mp.nbt;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x)
    dst[i] = src[(i % per_tile) >> mp.bx_log];
}

// Rows [n_begin, n_end) only.
__global__ void k_broadcast_rows_range(const float* __restrict__ src, float* __restrict__ dst, int64_t N,
                                       int64_t n_begin, int64_t n_end, BatchMap mp) {
  const int64_t per_tile = (n_end - n_begin) << mp.bx_log;
  const int64_t total = per_tile * mp.nbt;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t tile = i / per_tile, r = i - tile * per_tile;
    const int64_t n = n_begin + (r >> mp.bx_log);
    dst[((tile * N + n) << mp.bx_log) + (r & ((1 << mp.bx_log) - 1))] = src[n];
  }
}

// A shared, normalised [N] message vector -> binary-difference storage of every sample:
// compressed rows [c_begin, c_begin + count) <- src[first_msg + 2c + 1] - src[first_msg + 2c].
__global__ void k_broadcast_bin(const float* __restrict__ src, int64_t first_msg, float* __restrict__ dst,
                                int64_t c_rows, int64_t c_begin, int64_t count, int nbt) {
  const int64_t per_tile = count << 5;
  const int64_t total = per_tile * nbt;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t tile = i / per_tile, r = i - tile * per_tile;
    const int64_t c = r >> 5;
    dst[((tile * c_rows + c_begin + c) << 5) + (r & 31)] = src[first_msg + 2 * c + 1] - src[first_msg + 2 * c];
  }
}

// ---------------------------------------------------------------------------
// normalize_and_clip_msgs applied to the INPUT messages (pgmax/infer/bp.py:92-96,
// 249-259): per edge subtract the max over its states, clip below at -1e32.
// In place on the tile-blocked buffer.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_normalize_edges(BatchMap mp, int64_t num_edges, int64_t Es,
                  const int32_t* __restrict__ edge_msg_start, float* __restrict__ m) {
  UnitLoop L = unit_loop(mp, num_edges);
  if (!L.b_ok) return;
  float* mL = m + lane_off(mp, Es, L.b);
  const int sh = mp.bx_log;
  for (int64_t e = L.u; e < L.u_end; e += L.step) {
    const int64_t s0 = edge_msg_start[e], s1 = edge_msg_start[e + 1];
    float mx = -INFINITY;
    for (int64_t s = s0; s < s1; ++s) mx = fmaxf(mx, mL[s << sh]);
    for (int64_t s = s0; s < s1; ++s) mL[s << sh] = fmaxf(mL[s << sh] - mx, kMsgNegInf);
  }
}

// ---------------------------------------------------------------------------
// K1: variable sums  S_v = ev_v + sum_{e incident to v} m_e, accumulated in
// ascending message index starting from the evidence (the order of a serial
// scatter-add, pgmax/infer/bp.py:217).  One thread per (var-state, sample)
// walking the variable's incident-edge list (CSR built by the plan).
// ---------------------------------------------------------------------------
constexpr int kVsStateBits = 12;  // vs_csr packing: states per variable < 4096, degree < 2^19
constexpr int kVsUnits = 4;    // var-states processed together by one thread
constexpr int kVsLowDeg = 4;   // ... when each has at most this many incident edges

__global__ void __launch_bounds__(kThreads)
k_var_sums(BatchMap mp, int64_t num_var_states, int64_t Es, const int2* __restrict__ vs_csr,
           const int32_t* __restrict__ var_edge_msg, View ev, const float* __restrict__ m,
           float* __restrict__ S, int m_shared = 0) {
  // m_shared: `m` is ONE [Es] vector shared by every sample (initial messages not batched)
  // vs_csr[v] = (CSR begin, degree << kVsStateBits | state offset within the variable): one
  // 8-byte index load per var-state instead of the chain var-state -> variable -> CSR row.  A thread takes
  // kVsUnits var-states per iteration: their rows are loaded together, and when all of them
  // are low-degree (the common case in sparse graphs) so are all their gathers, which keeps
  // 4 x more bytes in flight per thread than one short dependent chain at a time.
  UnitLoop L = unit_loop(mp, num_var_states);
  if (!L.b_ok) return;
  const LaneView evL = lane_view(ev, mp, L.b);
  const float* mL = m_shared ? m : m + lane_off(mp, Es, L.b);
  const int msh = m_shared ? 0 : mp.bx_log;
  float* SL = S + lane_off(mp, num_var_states, L.b);
  const int sh = mp.bx_log;
  for (int64_t v0 = L.u; v0 < L.u_end; v0 += kVsUnits * L.step) {
    int4 row[kVsUnits];  // (begin, end, state offset)
    bool low = true;
#pragma unroll
    for (int u = 0; u < kVsUnits; ++u) {
      const int64_t v = v0 + u * L.step;
      const int2 r = v < L.u_end ? vs_csr[v] : make_int2(0, 0);
      row[u] = make_int4(r.x, r.x + (r.y >> kVsStateBits), r.y & ((1 << kVsStateBits) - 1), 0);
      low = low && (row[u].y - row[u].x <= kVsLowDeg);
    }
    if (low) {
      float acc[kVsUnits], x[kVsUnits][kVsLowDeg];
#pragma unroll
      for (int u = 0; u < kVsUnits; ++u) {
        const int64_t v = v0 + u * L.step;
        acc[u] = v < L.u_end ? evL.at(v) : 0.f;
#pragma unroll
        for (int j = 0; j < kVsLowDeg; ++j)
          x[u][j] = (row[u].x + j < row[u].y) ? mL[(int64_t(var_edge_msg[row[u].x + j]) + row[u].z) << msh] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kVsUnits; ++u) {
#pragma unroll
        for (int j = 0; j < kVsLowDeg; ++j)
          if (row[u].x + j < row[u].y) acc[u] += x[u][j];
        const int64_t v = v0 + u * L.step;
        if (v < L.u_end) SL[v << sh] = acc[u];
      }
      continue;
    }
#pragma unroll 1
    for (int u = 0; u < kVsUnits; ++u) {
      const int64_t v = v0 + u * L.step;
      if (v >= L.u_end) break;
      const int64_t st = row[u].z, k1 = row[u].y;
      float acc = evL.at(v);
      int64_t k = row[u].x;
      // loads are independent of the running sum: issue 16 / 4 at a time (high-degree
      // variables - RBM units, shared deconvolution features - would otherwise serialise
      // one DRAM latency per edge), add in ascending order
      for (; k + 16 <= k1; k += 16) {
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = mL[(var_edge_msg[k + j] + st) << msh];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc += x[j];
      }
      for (; k + 4 <= k1; k += 4) {
        const float a0 = mL[(var_edge_msg[k] + st) << msh];
        const float a1 = mL[(var_edge_msg[k + 1] + st) << msh];
        const float a2 = mL[(var_edge_msg[k + 2] + st) << msh];
        const float a3 = mL[(var_edge_msg[k + 3] + st) << msh];
        acc += a0; acc += a1; acc += a2; acc += a3;
      }
      for (; k < k1; ++k) acc += mL[(var_edge_msg[k] + st) << msh];
      SL[v << sh] = acc;
    }
  }
}

// K1-list: the same sums for a LIST of var-states only (full sample tiles, TW = 32): the
// high-degree var-states when the factor kernels re-derive the sums of low-degree variables
// themselves (k_logical_pull_*).  A warp = the 32 samples of one var-state; 32 gathers in
// flight per lane, added in ascending message index.
constexpr int kVsListChunk = 32;

__global__ void __launch_bounds__(32)
k_var_sums_list(int batch, int nbt, int64_t Es, int64_t Vs, const int2* __restrict__ vs_csr,
                const int32_t* __restrict__ var_edge_msg, const int32_t* __restrict__ list, int64_t list_len,
                View ev, const float* __restrict__ m, float* __restrict__ S) {
  // launched with ONE warp per CTA: a long serial chain (a variable with hundreds of edges) then
  // holds only its own warp's resources, not a whole CTA of finished warps.  The list is sorted
  // by degree, longest first, and the sample tile is the FASTEST block coordinate, so the long
  // chains of all tiles start at once and the short rows fill in behind.
  const int lane = threadIdx.x & 31;
  const int tile_i = int(blockIdx.x % unsigned(nbt));
  const bool live = tile_i * 32 + lane < batch;
  const int ll = live ? lane : 0;  // dead lanes shadow sample 0 of the tile (they stay for the shuffles)
  const size_t tile = tile_i;
  const float* mL = m + tile * size_t(Es) * 32 + ll;
  float* SL = S + tile * size_t(Vs) * 32 + ll;
  const float* evq = ev.kind == 1 ? ev.p + tile * size_t(ev.n_rows) * 32 + ll : ev.p;
  const int esh = ev.kind == 1 ? 5 : 0;
  const int64_t gwarp = blockIdx.x / unsigned(nbt);
  const int64_t nwarps = gridDim.x / unsigned(nbt);
  for (int64_t i = gwarp; i < list_len; i += nwarps) {
    const int v = list[i];
    const int2 r = vs_csr[v];
    const int st = r.y & ((1 << kVsStateBits) - 1);
    const int k1 = r.x + (r.y >> kVsStateBits);
    float acc = evq[uint32_t(v) << esh];
    // 32 incident edges per round: ONE coalesced index load (lane j: edge k + j), indices handed
    // out by shuffles, 32 gathers in flight per lane, added in ascending message index
    int mine = r.x + lane < k1 ? var_edge_msg[r.x + lane] : 0;
    for (int k = r.x; k < k1; k += kVsListChunk) {
      const int held = mine;
      if (k + kVsListChunk + lane < k1) mine = var_edge_msg[k + kVsListChunk + lane];
      float x[kVsListChunk];
#pragma unroll
      for (int j = 0; j < kVsListChunk; ++j) {
        const int idx = __shfl_sync(0xffffffffu, held, j);
        x[j] = (k + j < k1) ? mL[uint32_t(idx + st) << 5] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < kVsListChunk; ++j)
        if (k + j < k1) acc += x[j];
    }
    if (live) SL[uint32_t(v) << 5] = acc;
  }
}

// K1-list on binary-difference storage: `list` holds the var-states of the listed variables,
// state 0 and state 1 of a variable adjacent; a warp = the 32 samples of ONE variable and
// accumulates both sums in one walk (each stored difference is read once).
__global__ void __launch_bounds__(32)
k_var_sums_list_bin(int batch, int nbt, int64_t E, int64_t Vs, const int2* __restrict__ vs_csr,
                    const int32_t* __restrict__ var_edge_msg, const int32_t* __restrict__ list, int64_t list_len,
                    View ev, const float* __restrict__ c, float* __restrict__ S) {
  const int lane = threadIdx.x & 31;
  const int tile_i = int(blockIdx.x % unsigned(nbt));
  const bool live = tile_i * 32 + lane < batch;
  const int ll = live ? lane : 0;
  const size_t tile = tile_i;
  const float* cL = c + tile * size_t(E) * 32 + ll;
  float* SL = S + tile * size_t(Vs) * 32 + ll;
  const float* evq = ev.kind == 1 ? ev.p + tile * size_t(ev.n_rows) * 32 + ll : ev.p;
  const int esh = ev.kind == 1 ? 5 : 0;
  const int64_t gwarp = blockIdx.x / unsigned(nbt);
  const int64_t nwarps = gridDim.x / unsigned(nbt);
  for (int64_t i = gwarp; 2 * i < list_len; i += nwarps) {
    const int v = list[2 * i];  // var-state of state 0; state 1 is v + 1
    const int2 r = vs_csr[v];
    const int k1 = r.x + (r.y >> kVsStateBits);
    float acc0 = evq[uint32_t(v) << esh], acc1 = evq[uint32_t(v + 1) << esh];
    int mine = r.x + lane < k1 ? var_edge_msg[r.x + lane] : 0;
    for (int k = r.x; k < k1; k += kVsListChunk) {
      const int held = mine;
      if (k + kVsListChunk + lane < k1) mine = var_edge_msg[k + kVsListChunk + lane];
      float x[kVsListChunk];
#pragma unroll
      for (int j = 0; j < kVsListChunk; ++j) {
        const int idx = __shfl_sync(0xffffffffu, held, j);
        x[j] = (k + j < k1) ? cL[(uint32_t(idx) >> 1) << 5] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < kVsListChunk; ++j)
        if (k + j < k1) {
          const bool fl = x[j] != x[j];  // both states at the floor (load_msg)
          acc0 += fl ? kMsgNegInf : fminf(-x[j], 0.f);
          acc1 += fl ? kMsgNegInf : fminf(x[j], 0.f);
        }
    }
    if (live) {
      SL[uint32_t(v) << 5] = acc0;
      SL[uint32_t(v + 1) << 5] = acc1;
    }
  }
}

// Full tile-blocked messages (normalised, every edge two states) -> binary-difference storage.
__global__ void __launch_bounds__(kThreads)
k_compress_bin(const float* __restrict__ m, float* __restrict__ c, int64_t E, int nbt) {
  const int64_t total = E * 32 * nbt;  // one float per (tile, edge, sample)
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t row = i >> 5;  // tile * E + e
    const int l = int(i & 31);
    c[i] = m[(2 * row + 1) * 32 + l] - m[(2 * row) * 32 + l];
  }
}

// ---------------------------------------------------------------------------
// Shared epilogue: damping, per-edge max-normalisation, clip, delta
// (pgmax/infer/bp.py:127-136).  `one_minus_d` is computed on the host in fp32.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float damp(float m_old, float f, float d, float one_minus_d) {
  return d * m_old + one_minus_d * f;
}

// Scalars of one run shared by every factor->variable kernel.
struct RunArgs {
  float d, one_minus_d;  // damping
  float T;               // temperature
  float c_exp, c_log;    // log2(e) / T and T * ln(2) (fast pairwise sum-product path)
  float* deltas;         // [batch][delta_stride] or null
  int64_t delta_stride, delta_off;
  int64_t Es, Vs;        // rows of the message / var-sum arrays
};

// ---------------------------------------------------------------------------
// Pairwise binary EnumFactor with all 4 configurations valid (PairwiseFactorGroup
// over binary variables: Ising, RBM).  Everything in registers.
//   s_k = (q_a + q_b) + lp_k;  M_e = max over the 2 configs containing e;
//   T = 0: f_e = M_e - q_e;  T > 0: f_e = (T log sum exp((s_k - M_e)/T) + M_e) - q_e
// (pgmax/factor/enum.py:451-475, update_utils.py:68-98.)  With two terms the sum is
// exp(0) + exp((min - max)/T) = 1 + e, exactly as the reference forms it.
// This is the one place the library trades the last bits for speed: the pair
// kernels are bandwidth-bound only if the four softplus terms per factor are
// cheap, so e and log(1 + e) use the hardware ex2 / lg2 units
// (e = ex2((min - max) * log2(e)/T), relative error 2^-22; lg2 on (1, 2] has
// absolute error <= 2^-22, i.e. <= 1.7e-7 * T on the message) instead of the
// ~33-instruction expf / logf pair.  Messages are O(1..10), where one fp32 ulp is
// 1e-7..1e-6, and the north-star tolerance for sum-product is 1e-5.
// Max-product (T = 0) involves no transcendental and stays bit-exact.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool kSumProduct>
__device__ __forceinline__ float lse2(float a, float b, float c_exp, float c_log) {
  const float mx = fmaxf(a, b);
  if (!kSumProduct) return mx;
  // min - max == -|a - b| exactly (one subtraction either way): one FADD, the sign and the
  // absolute value ride on the FMUL as operand modifiers
  return __fmaf_rn(c_log, lg2_approx(1.0f + ex2_approx(-fabsf(a - b) * c_exp)), mx);
}

// In: old messages m[4] = (v0s0, v0s1, v1s0, v1s1), var sums S[4] of the same
// var-states, clipped potentials lp[4] in config order (0,0),(0,1),(1,0),(1,1).
// Out: n[4] damped + normalised + clipped; returns max|n - m| (kDelta; else 0).
// Damping: max-product rounds d*m and (1-d)*f separately, as the reference does (bit-exact
// with the oracle); sum-product, which already carries the 1e-7-level ex2/lg2 error, fuses
// the second product into an FMA.
template <bool kSumProduct, bool kDelta = true>
__device__ __forceinline__ float pw2_update(const float (&m)[4], const float (&Sv)[4],
                                            const float (&lp)[4], const RunArgs& a,
                                            float (&n)[4]) {
  const float q0 = Sv[0] - m[0], q1 = Sv[1] - m[1], q2 = Sv[2] - m[2], q3 = Sv[3] - m[3];
  const float s00 = (q0 + q2) + lp[0], s01 = (q0 + q3) + lp[1];
  const float s10 = (q1 + q2) + lp[2], s11 = (q1 + q3) + lp[3];
  const float f0 = lse2<kSumProduct>(s00, s01, a.c_exp, a.c_log) - q0;
  const float f1 = lse2<kSumProduct>(s10, s11, a.c_exp, a.c_log) - q1;
  const float f2 = lse2<kSumProduct>(s00, s10, a.c_exp, a.c_log) - q2;
  const float f3 = lse2<kSumProduct>(s01, s11, a.c_exp, a.c_log) - q3;
  float n0, n1, n2, n3;
  if (kSumProduct) {
    n0 = __fmaf_rn(a.d, m[0], a.one_minus_d * f0); n1 = __fmaf_rn(a.d, m[1], a.one_minus_d * f1);
    n2 = __fmaf_rn(a.d, m[2], a.one_minus_d * f2); n3 = __fmaf_rn(a.d, m[3], a.one_minus_d * f3);
  } else {
    n0 = damp(m[0], f0, a.d, a.one_minus_d); n1 = damp(m[1], f1, a.d, a.one_minus_d);
    n2 = damp(m[2], f2, a.d, a.one_minus_d); n3 = damp(m[3], f3, a.d, a.one_minus_d);
  }
  const float mxa = fmaxf(n0, n1), mxb = fmaxf(n2, n3);
  n[0] = fmaxf(n0 - mxa, kMsgNegInf); n[1] = fmaxf(n1 - mxa, kMsgNegInf);
  n[2] = fmaxf(n2 - mxb, kMsgNegInf); n[3] = fmaxf(n3 - mxb, kMsgNegInf);
  if (!kDelta) return 0.f;
  return fmaxf(fmaxf(fabsf(n[0] - m[0]), fabsf(n[1] - m[1])),
               fmaxf(fabsf(n[2] - m[2]), fabsf(n[3] - m[3])));
}

// K2a: one thread per (factor, sample).
template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_enum_pw2(BatchMap mp, int64_t num_factors, int64_t first_edge, int64_t first_msg,
           int64_t first_pot, const int32_t* __restrict__ edge_vs, View lp,
           const float* __restrict__ S, const float* __restrict__ m_old,
           float* __restrict__ m_new, RunArgs a) {
  UnitLoop L = unit_loop(mp, num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int sh = mp.bx_log;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float* mo = m_old + moff;
  float* mn = m_new + moff;
  const float* SL = S + lane_off(mp, a.Vs, L.b);
  const LaneView lpL = lane_view(lp, mp, L.b);
  // One sample (flat vectors, one factor per lane): the factor's 4 messages, its 4 potentials
  // and each variable's 2 sums are contiguous -> 128-bit / 64-bit accesses when aligned.
  const bool vec = sh == 0 && lpL.sh == 0 && ((first_msg | first_pot) & 3) == 0 && (first_edge & 1) == 0 &&
                   ((reinterpret_cast<uintptr_t>(mo) | reinterpret_cast<uintptr_t>(mn) |
                     reinterpret_cast<uintptr_t>(lpL.q)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(SL) & 7) == 0;
  if (vec) {
    for (int64_t f = L.u; f < L.u_end; f += L.step) {
      const int2 vs = *reinterpret_cast<const int2*>(edge_vs + first_edge + 2 * f);
      const int64_t mb = first_msg + 4 * f;
      const float4 m4 = *reinterpret_cast<const float4*>(mo + mb);
      const float4 l4 = *reinterpret_cast<const float4*>(lpL.q + first_pot + 4 * f);
      float Sv[4];
      if (((vs.x | vs.y) & 1) == 0) {
        const float2 s0 = *reinterpret_cast<const float2*>(SL + vs.x);
        const float2 s1 = *reinterpret_cast<const float2*>(SL + vs.y);
        Sv[0] = s0.x; Sv[1] = s0.y; Sv[2] = s1.x; Sv[3] = s1.y;
      } else {
        Sv[0] = SL[vs.x]; Sv[1] = SL[vs.x + 1]; Sv[2] = SL[vs.y]; Sv[3] = SL[vs.y + 1];
      }
      const float m[4] = {m4.x, m4.y, m4.z, m4.w};
      const float lpv[4] = {clip_lp(l4.x), clip_lp(l4.y), clip_lp(l4.z), clip_lp(l4.w)};
      float n[4];
      dmax = fmaxf(dmax, pw2_update<kSumProduct>(m, Sv, lpv, a, n));
      *reinterpret_cast<float4*>(mn + mb) = make_float4(n[0], n[1], n[2], n[3]);
    }
  } else {
    for (int64_t f = L.u; f < L.u_end; f += L.step) {
      const int64_t e = first_edge + 2 * f;
      const int64_t vs0 = edge_vs[e], vs1 = edge_vs[e + 1];
      const int64_t mb = first_msg + 4 * f;
      const float m[4] = {mo[mb << sh], mo[(mb + 1) << sh], mo[(mb + 2) << sh], mo[(mb + 3) << sh]};
      const float Sv[4] = {SL[vs0 << sh], SL[(vs0 + 1) << sh], SL[vs1 << sh], SL[(vs1 + 1) << sh]};
      const int64_t pb = first_pot + 4 * f;
      const float lpv[4] = {clip_lp(lpL.at(pb)), clip_lp(lpL.at(pb + 1)), clip_lp(lpL.at(pb + 2)),
                            clip_lp(lpL.at(pb + 3))};
      float n[4];
      dmax = fmaxf(dmax, pw2_update<kSumProduct>(m, Sv, lpv, a, n));
      mn[mb << sh] = n[0]; mn[(mb + 1) << sh] = n[1]; mn[(mb + 2) << sh] = n[2]; mn[(mb + 3) << sh] = n[3];
    }
  }
  publish_delta(a.deltas, int64_t(L.b) * a.delta_stride + a.delta_off, dmax);
}

// ---------------------------------------------------------------------------
// K2a-pull: pairwise-binary block on LOW-DEGREE variables (grids: Ising).  One
// pass per iteration without a var-sum array: every (factor, sample) thread
// re-derives the sums of its two variables by walking their incident-edge lists
// (evidence first, then messages in ascending index: the same serial order as
// k_var_sums, hence bit-identical results).  The degree-fold re-reads hit L1/L2
// (neighbouring factors share variables); HBM sees the messages once.
// edge_csr[e] = (begin, end) of the CSR row of edge e's variable.
// ---------------------------------------------------------------------------
constexpr int kPullMaxDegree = 6;

struct PullArgs {
  int64_t num_factors, first_edge, first_msg, first_pot;
  const int32_t* edge_vs;
  const int2* edge_csr;
  const int32_t* var_edge_msg;
};

// kViaL2: message loads bypass L1 (ld.global.cg) - required inside the persistent kernel,
// where other SMs rewrite the buffers between iterations.
template <bool kSumProduct, bool kViaL2>
__device__ __forceinline__ float pull_factors(const BatchMap& mp, const UnitLoop& L, const PullArgs& g,
                                              const LaneView& evL, const LaneView& lpL,
                                              const float* mo_, float* mn, const RunArgs& a) {
  const int sh = mp.bx_log;
  float dmax = 0.f;
  struct Loader {
    const float* p;
    __device__ __forceinline__ float operator[](int64_t i) const { return kViaL2 ? __ldcg(p + i) : p[i]; }
  } mo{mo_};
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t e = g.first_edge + 2 * f;
    const int64_t vs0 = g.edge_vs[e], vs1 = g.edge_vs[e + 1];
    const int2 c0 = g.edge_csr[e], c1 = g.edge_csr[e + 1];
    float Sv[4] = {evL.at(vs0), evL.at(vs0 + 1), evL.at(vs1), evL.at(vs1 + 1)};
    // incident-edge lists (degree <= kPullMaxDegree): all index loads first, then all
    // message loads, then the additions in ascending message order
    int32_t i0[kPullMaxDegree], i1[kPullMaxDegree];
#pragma unroll
    for (int k = 0; k < kPullMaxDegree; ++k) {
      i0[k] = (c0.x + k < c0.y) ? g.var_edge_msg[c0.x + k] : -1;
      i1[k] = (c1.x + k < c1.y) ? g.var_edge_msg[c1.x + k] : -1;
    }
    float g0[kPullMaxDegree][2], g1[kPullMaxDegree][2];
#pragma unroll
    for (int k = 0; k < kPullMaxDegree; ++k) {
      g0[k][0] = i0[k] >= 0 ? mo[int64_t(i0[k]) << sh] : 0.f;
      g0[k][1] = i0[k] >= 0 ? mo[int64_t(i0[k] + 1) << sh] : 0.f;
      g1[k][0] = i1[k] >= 0 ? mo[int64_t(i1[k]) << sh] : 0.f;
      g1[k][1] = i1[k] >= 0 ? mo[int64_t(i1[k] + 1) << sh] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kPullMaxDegree; ++k) {
      if (i0[k] >= 0) { Sv[0] += g0[k][0]; Sv[1] += g0[k][1]; }
      if (i1[k] >= 0) { Sv[2] += g1[k][0]; Sv[3] += g1[k][1]; }
    }
    const int64_t mb = g.first_msg + 4 * f;
    const float m[4] = {mo[mb << sh], mo[(mb + 1) << sh], mo[(mb + 2) << sh], mo[(mb + 3) << sh]};
    const int64_t pb = g.first_pot + 4 * f;
    const float lpv[4] = {clip_lp(lpL.at(pb)), clip_lp(lpL.at(pb + 1)), clip_lp(lpL.at(pb + 2)),
                          clip_lp(lpL.at(pb + 3))};
    float n[4];
    dmax = fmaxf(dmax, pw2_update<kSumProduct>(m, Sv, lpv, a, n));
    mn[mb << sh] = n[0]; mn[(mb + 1) << sh] = n[1]; mn[(mb + 2) << sh] = n[2]; mn[(mb + 3) << sh] = n[3];
  }
  return dmax;
}

template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_enum_pw2_pull(BatchMap mp, PullArgs g, View ev, View lp, const float* __restrict__ m_old,
                float* __restrict__ m_new, RunArgs a) {
  UnitLoop L = unit_loop(mp, g.num_factors);
  if (!L.b_ok) return;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float dmax = pull_factors<kSumProduct, false>(mp, L, g, lane_view(ev, mp, L.b),
                                                      lane_view(lp, mp, L.b), m_old + moff, m_new + moff, a);
  publish_delta(a.deltas, int64_t(L.b) * a.delta_stride + a.delta_off, dmax);
}

// Resident variant for graphs that are ONE pull block and small enough that every
// (factor, sample) pair gets its own thread (Ising 50x50: 5 000 factors, 1000
// iterations: latency-bound, the working set lives in L2): ALL iterations in one
// launch.  Each thread keeps its factor's indices, evidence and potentials in
// registers across iterations; per iteration it issues its (<= 2 * degree + 4)
// message loads at once (ld.global.cg: other SMs rewrite the buffers), updates,
// stores, and joins one barrier:
//   kCluster = true : the grid is ONE thread-block cluster (<= 16 CTAs); the barrier is
//                     the hardware cluster barrier (arrive.release / wait.acquire);
//   kCluster = false: cooperative launch; barrier = monotonic counter in global memory.
// Buffers: iteration 0 reads `src0`; iteration `it` writes `out` if it is the last one
// and out != null, else it ping-pongs between bufA and bufB (never writing src0).
constexpr int kResidentClusterThreads = 384;
constexpr int kResidentClusterCtas = 16;  // non-portable cluster size (B200 allows 16)

template <bool kSumProduct, bool kCluster>
__global__ void __launch_bounds__(kCluster ? kResidentClusterThreads : kThreads)
k_enum_pw2_pull_resident(BatchMap mp, PullArgs g, View ev, View lp, const float* src0, float* bufA,
                         float* bufB, float* out, int num_iters, RunArgs a, unsigned int* bar) {
  const int lane = threadIdx.x & 31;
  const int64_t gwarp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int bx = 1 << mp.bx_log, upw = 32 >> mp.bx_log, sh = mp.bx_log;
  const int64_t wpt = (g.num_factors + upw - 1) / upw;  // warps per sample tile
  const int tile = int(gwarp / wpt);
  const int64_t f = (gwarp - int64_t(tile) * wpt) * upw + (lane >> mp.bx_log);
  const int b = tile * bx + (lane & (bx - 1));
  const bool ok = tile < mp.nbt && f < g.num_factors && b < mp.batch;

  // loop-invariant state of this thread's factor
  int32_t i0[kPullMaxDegree], i1[kPullMaxDegree];
  float ev4[4] = {0.f, 0.f, 0.f, 0.f}, lpv[4] = {0.f, 0.f, 0.f, 0.f};
  int64_t moff = 0, mb = 0;
#pragma unroll
  for (int k = 0; k < kPullMaxDegree; ++k) i0[k] = i1[k] = -1;
  if (ok) {
    const int64_t e = g.first_edge + 2 * f;
    const int64_t vs0 = g.edge_vs[e], vs1 = g.edge_vs[e + 1];
    const int2 c0 = g.edge_csr[e], c1 = g.edge_csr[e + 1];
#pragma unroll
    for (int k = 0; k < kPullMaxDegree; ++k) {
      if (c0.x + k < c0.y) i0[k] = g.var_edge_msg[c0.x + k];
      if (c1.x + k < c1.y) i1[k] = g.var_edge_msg[c1.x + k];
    }
    const LaneView evL = lane_view(ev, mp, b), lpL = lane_view(lp, mp, b);
    ev4[0] = evL.at(vs0); ev4[1] = evL.at(vs0 + 1); ev4[2] = evL.at(vs1); ev4[3] = evL.at(vs1 + 1);
    const int64_t pb = g.first_pot + 4 * f;
#pragma unroll
    for (int k = 0; k < 4; ++k) lpv[k] = clip_lp(lpL.at(pb + k));
    moff = lane_off(mp, a.Es, b);
    mb = g.first_msg + 4 * f;
  }
  const unsigned int nblocks = gridDim.x;
  const float* cur = src0;
  float* nxt = (src0 == bufA) ? bufB : bufA;
  for (int it = 0; it < num_iters; ++it) {
    float* dst = (it == num_iters - 1 && out != nullptr) ? out : nxt;
    if (ok) {
      const float* mo = cur + moff;
      float g0[kPullMaxDegree][2], g1[kPullMaxDegree][2], m[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) m[k] = __ldcg(mo + ((mb + k) << sh));
#pragma unroll
      for (int k = 0; k < kPullMaxDegree; ++k) {
        g0[k][0] = i0[k] >= 0 ? __ldcg(mo + (int64_t(i0[k]) << sh)) : 0.f;
        g0[k][1] = i0[k] >= 0 ? __ldcg(mo + (int64_t(i0[k] + 1) << sh)) : 0.f;
        g1[k][0] = i1[k] >= 0 ? __ldcg(mo + (int64_t(i1[k]) << sh)) : 0.f;
        g1[k][1] = i1[k] >= 0 ? __ldcg(mo + (int64_t(i1[k] + 1) << sh)) : 0.f;
      }
      float Sv[4] = {ev4[0], ev4[1], ev4[2], ev4[3]};
#pragma unroll
      for (int k = 0; k < kPullMaxDegree; ++k) {
        if (i0[k] >= 0) { Sv[0] += g0[k][0]; Sv[1] += g0[k][1]; }
        if (i1[k] >= 0) { Sv[2] += g1[k][0]; Sv[3] += g1[k][1]; }
      }
      float n[4];
      const float dmax = pw2_update<kSumProduct>(m, Sv, lpv, a, n);
      float* mn = dst + moff;
#pragma unroll
      for (int k = 0; k < 4; ++k) mn[(mb + k) << sh] = n[k];
      publish_delta(a.deltas, int64_t(b) * a.delta_stride + it, dmax);
    }
    nxt = (dst == bufA) ? bufB : bufA;
    cur = dst;
    if (it + 1 < num_iters) {
      if (kCluster) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
      } else {
        __syncthreads();
        if (threadIdx.x == 0) {
          __threadfence();
          atomicAdd(bar, 1u);
          const unsigned int target = (unsigned int)(it + 1) * nblocks;
          while (*reinterpret_cast<volatile unsigned int*>(bar) < target) {
          }
          __threadfence();
        }
        __syncthreads();
      }
    }
  }
}

// ---------------------------------------------------------------------------
// K2a-lattice: ONE pass per iteration, no index array at all, for a graph that is a 2-D
// nearest-neighbour lattice of binary variables built as examples/ising_model.ipynb cell 12
// builds it: variable (l, j) = l * N + j owns the vertical factor 2 * var (to (l + 1, j))
// and the horizontal factor 2 * var + 1 (to (l, j + 1) mod N), so that its 8 messages are
// the 32 contiguous bytes m[8 var .. 8 var + 7] = (V.v0s0, V.v0s1, V.v1s0, V.v1s1,
// H.v0s0, H.v0s1, H.v1s0, H.v1s1).  torus = 1: rows wrap (the notebook's graph);
// torus = 0: R owner rows plus a ghost row R that only receives (the row strips of
// dist.py).  The structure is detected from the generic edge table at plan time.
//
// A CTA owns a TR x TC tile of owner variables: it stages the messages of the tile plus a
// one-variable halo in shared memory (128-bit loads), forms the variable sums of the
// (TR + 1) x (TC + 1) variables its factors touch - evidence first, then the incident
// messages in ASCENDING MESSAGE INDEX, i.e. the order of the serial scatter-add of
// pgmax/infer/bp.py:217 and of k_var_sums, wrap-around neighbours included - and updates
// its 2 TR TC factors with 128-bit loads of the potentials and 128-bit stores.  Per
// iteration HBM sees the messages once in and once out, the potentials and the evidence:
// 13 bytes per edge-state, against 17 "algorithmic" ones (which include the incidence
// index this kernel does not need); halo re-reads hit L2 (neighbouring tiles run
// concurrently).  One sample only (flat vectors).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p);

struct LatticeDev {
  int64_t first_msg, first_pot;
  int32_t R, N;   // owner rows, columns
  int32_t torus;  // 1: rows wrap; 0: ghost row R below the last owner row
};

constexpr int kLatTR = 16, kLatTC = 64, kLatThreads = 256;
constexpr int kLatMR = kLatTR + 2, kLatMC = kLatTC + 2;
__host__ __device__ constexpr size_t lattice_smem_bytes() {
  return size_t(kLatMR) * kLatMC * 2 * sizeof(float4) + size_t(kLatTR + 1) * (kLatTC + 1) * sizeof(float2);
}

template <bool kSumProduct, bool kDelta>
__global__ void __launch_bounds__(kLatThreads)
k_lattice(LatticeDev g, const float* __restrict__ ev, const float* __restrict__ lp,
          const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  constexpr int TR = kLatTR, TC = kLatTC, MR = kLatMR, MC = kLatMC;
  extern __shared__ float4 lat_smem[];
  float4* sm = lat_smem;                                      // [MR][MC][2]: (V, H) messages
  float2* Ss = reinterpret_cast<float2*>(sm + MR * MC * 2);   // [TR + 1][TC + 1] variable sums
  const int l0 = blockIdx.y * TR, j0 = blockIdx.x * TC;
  const int R = g.R, N = g.N;
  const bool torus = g.torus != 0;
  const float4* mo4 = reinterpret_cast<const float4*>(m_old + g.first_msg);
  float4* mn4 = reinterpret_cast<float4*>(m_new + g.first_msg);
  const float4* lp4 = reinterpret_cast<const float4*>(lp + g.first_pot);
  const float2* ev2 = reinterpret_cast<const float2*>(ev);

  // ---- phase 1: messages of rows l0 - 1 .. l0 + TR, columns j0 - 1 .. j0 + TC ----------
  // cp.async (LDGSTS, 16 B, L2 only): every load of the thread is in flight at once and no
  // register is held for it; the potentials of the thread's factors (needed in phase 3) are
  // requested now as well, so that their latency hides behind phases 1 and 2.
  constexpr int kLoads = (MR * MC * 2 + kLatThreads - 1) / kLatThreads;
#pragma unroll
  for (int k = 0; k < kLoads; ++k) {
    const int t = threadIdx.x + k * kLatThreads;
    const int cell = t >> 1;
    const int rr = cell / MC, cc = cell - rr * MC;
    int l = l0 - 1 + rr, j = j0 - 1 + cc;
    bool ok = j >= -1 && j <= N;
    j = (j < 0) ? N - 1 : (j == N ? 0 : j);
    if (torus) {
      ok = ok && l <= R;
      l = (l < 0) ? R - 1 : (l == R ? 0 : l);
    } else {
      ok = ok && l >= 0 && l < R;  // the ghost row owns no factor
    }
    if (t < MR * MC * 2) {
      if (ok)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sm + t)),
                     "l"(mo4 + (int64_t(l) * N + j) * 2 + (t & 1)) : "memory");
      else
        sm[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  constexpr int kFac = TR * TC * 2 / kLatThreads;
  float4 lq[kFac];
#pragma unroll
  for (int k = 0; k < kFac; ++k) {
    const int t = threadIdx.x + k * kLatThreads;
    const int cell = t >> 1;
    const int rr = cell / TC, cc = cell - rr * TC;
    const int l = min(l0 + rr, R - 1), j = min(j0 + cc, N - 1);
    const float4* src = lp4 + (int64_t(l) * N + j) * 2 + (t & 1);
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(lq[k].x), "=f"(lq[k].y), "=f"(lq[k].z), "=f"(lq[k].w) : "l"(src));
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- phase 2: variable sums of rows l0 .. l0 + TR, columns j0 .. j0 + TC -------------
  for (int t = threadIdx.x; t < (TR + 1) * (TC + 1); t += kLatThreads) {
    const int rr = t / (TC + 1), cc = t - rr * (TC + 1);
    int l = l0 + rr, j = j0 + cc;
    if (l > R || j > N) continue;
    if (j == N) j = 0;
    if (torus && l == R) l = 0;
    const bool has_own = l < R;              // owner row (always true on the torus)
    const bool has_up = torus || l > 0;
    const bool up_wrap = torus && l == 0;    // the upper neighbour is row R - 1: largest index
    const bool left_wrap = j == 0;           // the left neighbour is column N - 1: after the own ones
    const float2 e = __ldg(ev2 + (int64_t(l) * N + j));
    const float4 own_v = sm[((rr + 1) * MC + cc + 1) * 2], own_h = sm[((rr + 1) * MC + cc + 1) * 2 + 1];
    const float4 up = sm[(rr * MC + cc + 1) * 2];            // V factor of the row above: v1 slot
    const float4 left = sm[((rr + 1) * MC + cc) * 2 + 1];    // H factor of the left neighbour: v1 slot
    float s0 = e.x, s1 = e.y;
    if (has_up && !up_wrap) { s0 += up.z; s1 += up.w; }
    if (has_own && !left_wrap) { s0 += left.z; s1 += left.w; }
    if (has_own) { s0 += own_v.x; s1 += own_v.y; s0 += own_h.x; s1 += own_h.y; }
    if (has_own && left_wrap) { s0 += left.z; s1 += left.w; }
    if (up_wrap) { s0 += up.z; s1 += up.w; }
    Ss[t] = make_float2(s0, s1);
  }
  __syncthreads();

  // ---- phase 3: the 2 TR TC factors of the tile ---------------------------------------------
  float dmax = 0.f;
#pragma unroll
  for (int k = 0; k < kFac; ++k) {
    const int t = threadIdx.x + k * kLatThreads;
    const int tt = t & 1, cell = t >> 1;
    const int rr = cell / TC, cc = cell - rr * TC;
    const int l = l0 + rr, j = j0 + cc;
    if (l < R && j < N) {
      const int64_t f = (int64_t(l) * N + j) * 2 + tt;
      const float4 l4 = lq[k];
      const float4 m4 = sm[((rr + 1) * MC + cc + 1) * 2 + tt];
      const float2 sa = Ss[rr * (TC + 1) + cc];
      const float2 sb = tt == 0 ? Ss[(rr + 1) * (TC + 1) + cc] : Ss[rr * (TC + 1) + cc + 1];
      const float m[4] = {m4.x, m4.y, m4.z, m4.w};
      const float Sv[4] = {sa.x, sa.y, sb.x, sb.y};
      const float lpv[4] = {clip_lp(l4.x), clip_lp(l4.y), clip_lp(l4.z), clip_lp(l4.w)};
      float n[4];
      dmax = fmaxf(dmax, pw2_update<kSumProduct, kDelta>(m, Sv, lpv, a, n));
      mn4[f] = make_float4(n[0], n[1], n[2], n[3]);
    }
  }
  if (kDelta) {
    for (int o = 16; o > 0; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    if ((threadIdx.x & 31) == 0) publish_delta(a.deltas, a.delta_off, dmax);
  }
}

// ---------------------------------------------------------------------------
// TMA (bulk async copy) + mbarrier helpers, sm_90+ PTX.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared, completion signalled on `bar` (complete_tx::bytes)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global, tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------
// K2a-lattice, streaming variant: the same update as k_lattice (identical arithmetic, same
// summation order: bit-identical) as a PERSISTENT, warp-specialised kernel - one CTA per SM,
// tiles handed out round-robin in row-major order (the CTAs sweep the grid together, so a
// tile's halo rows are still in L2 when the next tile-row needs them):
//   * a producer warp moves every tile with TMA bulk copies: the message rows of the tile +
//     halo and the potential rows global -> shared on an mbarrier (<= 4 copies per row, one
//     lane per row), and the updated rows shared -> global as bulk groups; no LSU
//     instruction touches global memory for messages or potentials;
//   * 16 consumer warps wait for the tile, form the variable sums, update the factors IN
//     PLACE in shared memory and hand the tile back; two stages, so the loads of tile i + 1
//     and the stores of tile i - 1 overlap the arithmetic of tile i.
// Shared memory: 2 x (message tile 6 x 258 cells + potential tile 4 x 256 cells) x 32 B
// + 2 sum buffers = 185 KB.
// ---------------------------------------------------------------------------
constexpr int kLsConsumers = 512;               // 16 warps
constexpr int kLsThreads = kLsConsumers + 32;   // + producer warp
// Tile shape: 4 rows x 256 columns, two stages.  The producer's cost is the NUMBER of bulk
// copies (one per tile row and array, ~2 KB each at 64 columns), not their bytes: measured on
// Ising 8192^2 (ms per iteration) 16x64x2 stages 1.73, 16x48x3 2.10, 8x64x4 2.08, 8x128x2 1.45,
// 6x160x2 1.42, 4x256x2 1.415, 4x192x3 1.50, 2x512x2 1.45 - wide, flat tiles (8 KB copies) win;
// their extra halo rows (6 loaded per 4 updated) are L2 hits, the CTAs sweep the grid together.
constexpr int kLsStages = 2;
constexpr int kLsTR = 4, kLsTC = 256, kLsMR = kLsTR + 2, kLsMC = kLsTC + 2;
constexpr int kLsMsgF4 = kLsMR * kLsMC * 2;   // float4 per message stage
constexpr int kLsLpF4 = kLsTR * kLsTC * 2;    // float4 per potential stage
constexpr int kLsSumF2 = (kLsTR + 1) * (kLsTC + 1);
__host__ __device__ constexpr size_t lattice_stream_smem_bytes() {
  return size_t(kLsStages) * (kLsMsgF4 + kLsLpF4) * sizeof(float4) + size_t(2) * kLsSumF2 * sizeof(float2) +
         2 * kLsStages * sizeof(uint64_t);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <bool kSumProduct, bool kDelta>
__global__ void __launch_bounds__(kLsThreads, 1)
k_lattice_stream(LatticeDev g, const float* __restrict__ ev, const float* __restrict__ lp,
                 const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  constexpr int TR = kLsTR, TC = kLsTC, MC = kLsMC;
  extern __shared__ __align__(128) unsigned char ls_raw[];
  float4* msg_s = reinterpret_cast<float4*>(ls_raw);                       // [stage][MR][MC][2]
  float4* lp_s = msg_s + kLsStages * kLsMsgF4;                             // [stage][TR][TC][2]
  float2* sum_s = reinterpret_cast<float2*>(lp_s + kLsStages * kLsLpF4);   // [2][TR + 1][TC + 1]
  uint64_t* full = reinterpret_cast<uint64_t*>(sum_s + 2 * kLsSumF2);      // [stage] tile landed
  uint64_t* done = full + kLsStages;                                       // [stage] tile updated
  const int R = g.R, N = g.N;
  const bool torus = g.torus != 0;
  const int tiles_x = (N + TC - 1) / TC, tiles_y = (R + TR - 1) / TR;
  const int64_t num_tiles = int64_t(tiles_x) * tiles_y;
  const int64_t my_tiles = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kLsStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&done[s], kLsConsumers);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const float4* mo4 = reinterpret_cast<const float4*>(m_old + g.first_msg);
  float4* mn4 = reinterpret_cast<float4*>(m_new + g.first_msg);
  const float4* lp4 = reinterpret_cast<const float4*>(lp + g.first_pot);

  if (threadIdx.x >= kLsConsumers) {
    // ------------------------------- producer warp ---------------------------------------
    const int lane = threadIdx.x & 31;
    auto load_tile = [&](int64_t i) {
      const int64_t tile = blockIdx.x + i * gridDim.x;
      const int ty = int(tile / tiles_x), tx = int(tile - int64_t(ty) * tiles_x);
      const int l0 = ty * TR, j0 = tx * TC;
      const int stage = int(i % kLsStages);
      float4* ms = msg_s + stage * kLsMsgF4;
      float4* ls = lp_s + stage * kLsLpF4;
      const int ncols = min(TC + 1, N - j0);        // cells from column j0 on (incl. the right halo if inside)
      const bool wrap_right = j0 + TC >= N;         // right halo of the last valid column is column 0
      const int lcols = min(TC, N - j0);
      // rows: sm row rr <-> lattice row l0 - 1 + rr
      int l = l0 - 1 + lane;
      bool row_ok = lane < kLsMR;
      if (torus) { row_ok = row_ok && l <= R; l = l < 0 ? R - 1 : (l == R ? 0 : l); }
      else row_ok = row_ok && l >= 0 && l < R;
      const bool lp_ok = lane < TR && l0 + lane < R;
      const uint32_t row_bytes = uint32_t(32 + ncols * 32 + (wrap_right ? 32 : 0));
      const uint32_t my_bytes = (row_ok ? row_bytes : 0u) + (lp_ok ? uint32_t(lcols) * 32u : 0u);
      uint32_t total = my_bytes;
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
      if (lane == 0) mbar_expect_tx(&full[stage], total);
      __syncwarp();
      if (row_ok) {
        const float4* src = mo4 + int64_t(l) * N * 2;
        float4* dst = ms + lane * MC * 2;
        bulk_g2s(dst, src + int64_t(j0 == 0 ? N - 1 : j0 - 1) * 2, 32, &full[stage]);
        bulk_g2s(dst + 2, src + int64_t(j0) * 2, uint32_t(ncols) * 32u, &full[stage]);
        if (wrap_right) bulk_g2s(dst + 2 + ncols * 2, src, 32, &full[stage]);
      }
      if (lp_ok)
        bulk_g2s(ls + lane * TC * 2, lp4 + (int64_t(l0 + lane) * N + j0) * 2, uint32_t(lcols) * 32u, &full[stage]);
    };
    for (int s = 0; s < kLsStages - 1; ++s)
      if (s < my_tiles) load_tile(s);
    for (int64_t i = 0; i < my_tiles; ++i) {
      // the stage of tile i - 1 has been drained below: refill it with tile i + kLsStages - 1
      if (i + kLsStages - 1 < my_tiles) load_tile(i + kLsStages - 1);
      const int stage = int(i % kLsStages);
      mbar_wait(&done[stage], uint32_t(i / kLsStages) & 1u);
      const int64_t tile = blockIdx.x + i * gridDim.x;
      const int ty = int(tile / tiles_x), tx = int(tile - int64_t(ty) * tiles_x);
      const int l0 = ty * TR, j0 = tx * TC;
      const int lcols = min(TC, N - j0);
      if (lane < TR && l0 + lane < R) {
        bulk_s2g(mn4 + (int64_t(l0 + lane) * N + j0) * 2, msg_s + stage * kLsMsgF4 + ((lane + 1) * MC + 1) * 2,
                 uint32_t(lcols) * 32u);
        bulk_commit();
      }
      bulk_wait_read<0>();  // this stage's shared memory may be overwritten from here on
      __syncwarp();
    }
    return;
  }

  // --------------------------------- consumer warps -----------------------------------------
  const float2* ev2 = reinterpret_cast<const float2*>(ev);
  float dmax = 0.f;
  for (int64_t i = 0; i < my_tiles; ++i) {
    const int64_t tile = blockIdx.x + i * gridDim.x;
    const int ty = int(tile / tiles_x), tx = int(tile - int64_t(ty) * tiles_x);
    const int l0 = ty * TR, j0 = tx * TC;
    const int stage = int(i % kLsStages);
    float4* sm = msg_s + stage * kLsMsgF4;
    const float4* lq = lp_s + stage * kLsLpF4;
    float2* Ss = sum_s + (i & 1) * kLsSumF2;
    // evidence of this thread's variables: requested before the wait
    constexpr int kVars = (kLsSumF2 + kLsConsumers - 1) / kLsConsumers;
    float2 e[kVars];
#pragma unroll
    for (int k = 0; k < kVars; ++k) {
      const int t = threadIdx.x + k * kLsConsumers;
      const int rr = t / (TC + 1), cc = t - rr * (TC + 1);
      int l = l0 + rr, j = j0 + cc;
      e[k] = make_float2(0.f, 0.f);
      if (t < kLsSumF2 && l <= R && j <= N) {
        if (j == N) j = 0;
        if (torus && l == R) l = 0;
        e[k] = __ldg(ev2 + (int64_t(l) * N + j));
      }
    }
    mbar_wait(&full[stage], uint32_t(i / kLsStages) & 1u);
    // ---- variable sums (same order as k_lattice) -------------------------------------------
#pragma unroll
    for (int k = 0; k < kVars; ++k) {
      const int t = threadIdx.x + k * kLsConsumers;
      const int rr = t / (TC + 1), cc = t - rr * (TC + 1);
      int l = l0 + rr, j = j0 + cc;
      if (t >= kLsSumF2 || l > R || j > N) continue;
      if (j == N) j = 0;
      if (torus && l == R) l = 0;
      const bool has_own = l < R;
      const bool has_up = torus || l > 0;
      const bool up_wrap = torus && l == 0;
      const bool left_wrap = j == 0;
      // cell of column j0 + cc sits at sm column cc + 1, except the wrapped right halo
      const int col = (j0 + cc == N) ? (N - j0) + 1 : cc + 1;
      const float4 own_v = sm[((rr + 1) * MC + col) * 2], own_h = sm[((rr + 1) * MC + col) * 2 + 1];
      const float4 up = sm[(rr * MC + col) * 2];
      const float4 left = sm[((rr + 1) * MC + col - 1) * 2 + 1];
      float s0 = e[k].x, s1 = e[k].y;
      if (has_up && !up_wrap) { s0 += up.z; s1 += up.w; }
      if (has_own && !left_wrap) { s0 += left.z; s1 += left.w; }
      if (has_own) { s0 += own_v.x; s1 += own_v.y; s0 += own_h.x; s1 += own_h.y; }
      if (has_own && left_wrap) { s0 += left.z; s1 += left.w; }
      if (up_wrap) { s0 += up.z; s1 += up.w; }
      Ss[t] = make_float2(s0, s1);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kLsConsumers) : "memory");
    // ---- factors, in place --------------------------------------------------------------------
    constexpr int kFac = TR * TC * 2 / kLsConsumers;
#pragma unroll
    for (int k = 0; k < kFac; ++k) {
      const int t = threadIdx.x + k * kLsConsumers;
      const int tt = t & 1, cell = t >> 1;
      const int rr = cell / TC, cc = cell - rr * TC;
      if (l0 + rr < R && j0 + cc < N) {
        const float4 l4 = lq[t];
        float4* slot = sm + ((rr + 1) * MC + cc + 1) * 2 + tt;
        const float4 m4 = *slot;
        const float2 sa = Ss[rr * (TC + 1) + cc];
        const float2 sb = tt == 0 ? Ss[(rr + 1) * (TC + 1) + cc] : Ss[rr * (TC + 1) + cc + 1];
        const float m[4] = {m4.x, m4.y, m4.z, m4.w};
        const float Sv[4] = {sa.x, sa.y, sb.x, sb.y};
        const float lpv[4] = {clip_lp(l4.x), clip_lp(l4.y), clip_lp(l4.z), clip_lp(l4.w)};
        float n[4];
        dmax = fmaxf(dmax, pw2_update<kSumProduct, kDelta>(m, Sv, lpv, a, n));
        *slot = make_float4(n[0], n[1], n[2], n[3]);
      }
    }
    fence_proxy_async();
    mbar_arrive(&done[stage]);
  }
  if (kDelta) {
    for (int o = 16; o > 0; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    if ((threadIdx.x & 31) == 0) publish_delta(a.deltas, a.delta_off, dmax);
  }
}

// ---------------------------------------------------------------------------
// K2a-fused: a pairwise-binary block whose factors form a dense I x J grid,
// factor (i, j) = row variable i x column variable j, stored row-major (the RBM
// of benchmark/rbm_lib.py:138-169: i = hidden unit, j = visible unit).  One pass
// per iteration: besides the new messages the kernel produces, per warp tile, the
// partial sums of the NEW messages per variable, so that the next iteration's
// variable sums need no second read of the message array (k_var_reduce adds the
// partials in a fixed order: deterministic, but a tree order rather than the
// serial ascending order of k_var_sums).
//
// A warp owns (one tile of 32 samples) x (strip of TJ columns) x (chunk of RI
// rows).  In the tile-blocked layout the messages of TJ consecutive factors of
// one row are ONE contiguous span of TJ*4*128 B (8 KiB for TJ = 16): the warp
// streams its chunk row by row through a private ring of kBipStages shared-memory
// buffers with TMA bulk copies (global -> shared on an mbarrier; shared -> global
// as a bulk group), updates each row in place in shared memory, and never holds a
// message in a long-latency register load.  Column sums S_v and the column
// accumulators live in registers for the whole chunk, the row accumulator for
// one row.  The four warps of a CTA share strip and chunk (their potentials are
// staged once in shared memory) and cover four sample tiles.
// ---------------------------------------------------------------------------
struct BipDev {
  int64_t first_msg, first_pot;
  int64_t first_cmsg;    // first row of the block in the compressed (one float per edge) message array
  int32_t I, J;          // rows, columns
  int32_t NS, NR, RI;    // column strips, row chunks, rows per chunk
  const int32_t* row_vs;   // [I] var-state of state 0 of row variable i
  const int32_t* col_vs;   // [J]
  const int32_t* row_part; // [I] partial-buffer row of (row var i, state 0, strip 0); state s, strip k at +2k+s
  const int32_t* col_part; // [J] same for column variables / row chunks
};

constexpr int kBipTJ = 16;
constexpr int kBipStages = 3;  // ring depth
// warps (= sample tiles) per CTA: 8 with compressed input rows (two CTAs of 104 KiB per SM, 16
// warps: the kernel is issue-latency-bound, not bandwidth-bound, below that), 4 with full rows
__host__ __device__ constexpr int bip_warps(bool in_full) { return in_full ? 4 : 8; }

// dynamic shared memory of k_enum_pw2_bip
__host__ __device__ constexpr size_t bip_smem_bytes(int RI, int TJ, bool in_full) {
  return size_t(bip_warps(in_full)) * kBipStages * (in_full ? 4 : 2) * TJ * 32 * sizeof(float)  // rings
         + size_t(RI) * TJ * 4 * sizeof(float)                                                  // potentials
         + size_t(bip_warps(in_full)) * kBipStages * sizeof(uint64_t);                          // mbarriers
}

// Binary-difference storage.  A normalised message of a two-state edge is (n_p, n_r) with
// max(n_p, n_r) == 0 exactly, so the single float x = n_r - n_p carries both states without
// loss: n_p = min(-x, 0), n_r = min(x, 0) (one of the two is the exact zero, the other is
// +-x; the clip at -1e32 commutes).  Between iterations the fused kernel keeps only x: half
// the message traffic of the reference layout, bit-identical values.
__device__ __forceinline__ void bin_expand(float x, float& n_p, float& n_r) {
  n_p = fminf(-x, 0.f);
  n_r = fminf(x, 0.f);
}


// ---- packed fp32x2 arithmetic (sm_100a FADD2 / FMUL2 / FFMA2: two IEEE fp32 operations per
// issued instruction; each half rounds exactly like the scalar instruction) -----------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// Scalars of a run in packed form.
struct RunArgs2 {
  f32x2 d, one_minus_d, c_exp, c_log, one;
};
__device__ __forceinline__ RunArgs2 make_args2(const RunArgs& a) {
  RunArgs2 r;
  r.d = pk2(a.d, a.d);
  r.one_minus_d = pk2(a.one_minus_d, a.one_minus_d);
  r.c_exp = pk2(a.c_exp, a.c_exp);
  r.c_log = pk2(a.c_log, a.c_log);
  r.one = pk2(1.0f, 1.0f);
  return r;
}

// Two two-term logsumexps at once: lse(a_i, b_i) given the pairs (a - b) and max(a, b).
template <bool kSumProduct>
__device__ __forceinline__ f32x2 lse2_x2(f32x2 diff, f32x2 mx, const RunArgs2& c) {
  if (!kSumProduct) return mx;
  float t0, t1;
  upk2(mul2(diff, c.c_exp), t0, t1);
  float l0, l1;
  upk2(add2(pk2(ex2_approx(-fabsf(t0)), ex2_approx(-fabsf(t1))), c.one), l0, l1);
  return fma2(c.c_log, pk2(lg2_approx(l0), lg2_approx(l1)), mx);
}

// pw2_update on binary-difference storage, packed: xa / xb are the stored differences of the
// two edges, Sa / Sb the (state 0, state 1) variable sums, lp01 / lp23 the clipped potentials
// (0,0),(0,1) / (1,0),(1,1).  Returns the new differences and the normalised new messages
// na = (n0, n1), nb = (n2, n3) (for the partial sums); same operations and roundings as
// pw2_update followed by n1 - n0.
template <bool kSumProduct, bool kDelta>
__device__ __forceinline__ float pw2_update_bin(float xa, float xb, f32x2 Sa, f32x2 Sb, f32x2 lp01, f32x2 lp23,
                                                const RunArgs2& c, float& xa_new, float& xb_new, f32x2& na,
                                                f32x2& nb) {
  const f32x2 ma = pk2(fminf(-xa, 0.f), fminf(xa, 0.f)), mb = pk2(fminf(-xb, 0.f), fminf(xb, 0.f));
  const f32x2 qa = sub2(Sa, ma), qb = sub2(Sb, mb);
  float q0, q1, q2, q3;
  upk2(qa, q0, q1);
  upk2(qb, q2, q3);
  const f32x2 P = add2(pk2(q0 + q2, q0 + q3), lp01);  // (s00, s01)
  const f32x2 Q = add2(pk2(q1 + q2, q1 + q3), lp23);  // (s10, s11)
  float s00, s01, s10, s11;
  upk2(P, s00, s01);
  upk2(Q, s10, s11);
  // messages to variable b: lse over the state of a, element-wise on (P, Q)
  const f32x2 fb = sub2(lse2_x2<kSumProduct>(sub2(P, Q), pk2(fmaxf(s00, s10), fmaxf(s01, s11)), c), qb);
  // messages to variable a: lse over the state of b, within P and within Q
  const f32x2 fa = sub2(lse2_x2<kSumProduct>(pk2(s00 - s01, s10 - s11), pk2(fmaxf(s00, s01), fmaxf(s10, s11)), c), qa);
  f32x2 da, db;
  if (kSumProduct) {
    da = fma2(c.d, ma, mul2(c.one_minus_d, fa));
    db = fma2(c.d, mb, mul2(c.one_minus_d, fb));
  } else {
    da = add2(mul2(c.d, ma), mul2(c.one_minus_d, fa));
    db = add2(mul2(c.d, mb), mul2(c.one_minus_d, fb));
  }
  float n0, n1, n2, n3;
  upk2(da, n0, n1);
  upk2(db, n2, n3);
  // (n1 - mx) - (n0 - mx) with mx = max(n0, n1) is n1 - n0 exactly (one term is the exact 0);
  // the clip of the smaller state at -1e32 becomes a clamp of the difference
  xa_new = fminf(fmaxf(n1 - n0, kMsgNegInf), -kMsgNegInf);
  xb_new = fminf(fmaxf(n3 - n2, kMsgNegInf), -kMsgNegInf);
  na = pk2(fminf(-xa_new, 0.f), fminf(xa_new, 0.f));
  nb = pk2(fminf(-xb_new, 0.f), fminf(xb_new, 0.f));
  if (!kDelta) return 0.f;
  float e0, e1, e2, e3;
  upk2(sub2(na, ma), e0, e1);
  upk2(sub2(nb, mb), e2, e3);
  return fmaxf(fmaxf(fabsf(e0), fabsf(e1)), fmaxf(fabsf(e2), fabsf(e3)));
}

// kInFull: the input rows are in the full tile-blocked layout (first iteration of a run);
// the output is always compressed.
template <bool kSumProduct, int TJ, bool kDelta, bool kInFull>
__global__ void __launch_bounds__(bip_warps(kInFull) * 32)
k_enum_pw2_bip(int batch, int nbt_groups, BipDev g, const float* __restrict__ lp,
               const float* __restrict__ S, const float* __restrict__ m_old, int64_t old_rows,
               float* __restrict__ c_new, int64_t c_rows, float* __restrict__ part, int64_t part_rows,
               RunArgs a) {
  constexpr int kIn = kInFull ? 4 : 2;           // floats per factor and sample in the input rows
  constexpr int kStages = kBipStages;
  constexpr int kBipWarps = bip_warps(kInFull);
  constexpr int kRowFloats = TJ * kIn * 32;      // one input row of a strip for one sample tile
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);
  float* lp_s = ring + kBipWarps * kStages * kRowFloats;  // [RI][TJ][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(lp_s + g.RI * TJ * 4);

  // blockIdx.x = (chunk * NS + strip) * nbt_groups + sample-tile group
  const int grp = blockIdx.x % nbt_groups;
  const int sc = blockIdx.x / nbt_groups;
  const int js = sc % g.NS, rc = sc / g.NS;
  const int j0 = js * TJ, i0 = rc * g.RI;
  const int i1 = min(i0 + g.RI, g.I);
  const int nj = min(TJ, g.J - j0);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int bt = grp * kBipWarps + w;        // sample tile of this warp
  const bool active = bt * 32 < batch;       // whole warp in or out
  const int b = bt * 32 + lane;

  if (threadIdx.x < kBipWarps * kStages) mbar_init(&bars[threadIdx.x], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  float* my_ring = ring + w * kStages * kRowFloats;
  uint64_t* my_bar = bars + w * kStages;
  const uint32_t in_bytes = uint32_t(nj) * kIn * 32 * sizeof(float);
  const uint32_t out_bytes = uint32_t(nj) * 2 * 32 * sizeof(float);
  // global float offset of (row i, first factor of the strip) for this sample tile
  const int64_t in_base = (int64_t(bt) * old_rows + (kInFull ? g.first_msg : g.first_cmsg)) * 32;
  const int64_t out_base = (int64_t(bt) * c_rows + g.first_cmsg) * 32;
  auto in_off = [&](int i) { return in_base + (int64_t(i) * g.J + j0) * (kIn * 32); };
  auto out_off = [&](int i) { return out_base + (int64_t(i) * g.J + j0) * (2 * 32); };
  const int nrows = i1 - i0;
  if (active && lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s)
      if (s < nrows) {
        mbar_expect_tx(&my_bar[s], in_bytes);
        bulk_g2s(my_ring + s * kRowFloats, m_old + in_off(i0 + s), in_bytes, &my_bar[s]);
      }
  }
  // the chunk's potentials are staged while the first message rows are already in flight
  for (int t = threadIdx.x; t < (i1 - i0) * TJ * 4; t += blockDim.x) {
    const int r = t / (TJ * 4), c = t - r * (TJ * 4);
    lp_s[t] = (c < nj * 4) ? clip_lp(lp[g.first_pot + 4 * (int64_t(i0 + r) * g.J + j0) + c]) : 0.f;
  }
  __syncthreads();
  if (!active) return;

  const float* SL = S + (int64_t(bt) * a.Vs) * 32 + lane;
  float* PL = part + (int64_t(bt) * part_rows) * 32 + lane;
  const RunArgs2 c2 = make_args2(a);
  f32x2 Sc[TJ], ac[TJ];  // (state 0, state 1) pairs
#pragma unroll
  for (int jj = 0; jj < TJ; ++jj) {
    const int64_t vs = g.col_vs[min(j0 + jj, g.J - 1)];
    Sc[jj] = pk2(SL[vs * 32], SL[(vs + 1) * 32]);
    ac[jj] = 0ull;
  }
  float dmax = 0.f;
  int64_t rvs = g.row_vs[i0];
  float Sr0 = SL[rvs * 32], Sr1 = SL[(rvs + 1) * 32];
  for (int r = 0; r < nrows; ++r) {
    const int i = i0 + r;
    const int stage = r % kStages;
    float* buf = my_ring + stage * kRowFloats + lane;
    // row sums of the NEXT row: issue the loads before waiting on this row's data
    float nSr0 = 0.f, nSr1 = 0.f;
    if (r + 1 < nrows) {
      rvs = g.row_vs[i + 1];
      nSr0 = SL[rvs * 32];
      nSr1 = SL[(rvs + 1) * 32];
    }
    mbar_wait(&my_bar[stage], (r / kStages) & 1);
    const float* lrow = lp_s + r * TJ * 4;
    f32x2 ar = 0ull;
    const f32x2 Sr = pk2(Sr0, Sr1);
#pragma unroll
    for (int jj = 0; jj < TJ; ++jj) {
      if (jj < nj) {
        const float4 lq = *reinterpret_cast<const float4*>(lrow + 4 * jj);
        float xa, xb;
        if (kInFull) {  // normalised input: max(m0, m1) == 0, the difference is exact
          xa = buf[(4 * jj + 1) * 32] - buf[(4 * jj) * 32];
          xb = buf[(4 * jj + 3) * 32] - buf[(4 * jj + 2) * 32];
        } else {
          xa = buf[(2 * jj) * 32];
          xb = buf[(2 * jj + 1) * 32];
        }
        float xan, xbn;
        f32x2 na, nb;
        dmax = fmaxf(dmax, pw2_update_bin<kSumProduct, kDelta>(xa, xb, Sr, Sc[jj], pk2(lq.x, lq.y), pk2(lq.z, lq.w),
                                                               c2, xan, xbn, na, nb));
        // compressed in place: rows 2jj, 2jj+1 of the stage were read already (<= 4jj)
        buf[(2 * jj) * 32] = xan;
        buf[(2 * jj + 1) * 32] = xbn;
        ar = add2(ar, na);
        ac[jj] = add2(ac[jj], nb);
      }
    }
    float ar0, ar1;
    upk2(ar, ar0, ar1);
    const int64_t pr = (int64_t(g.row_part[i]) + 2 * js) * 32;
    PL[pr] = ar0;
    PL[pr + 32] = ar1;
    // the row is final in shared memory: hand it to the async proxy and store it
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(c_new + out_off(i), my_ring + stage * kRowFloats, out_bytes);
      bulk_commit();
      // refill the stage the PREVIOUS row used once its store has drained
      const int nr = r + kStages - 1;
      if (nr < nrows) {
        bulk_wait_read<1>();
        const int ns = nr % kStages;
        mbar_expect_tx(&my_bar[ns], in_bytes);
        bulk_g2s(my_ring + ns * kRowFloats, m_old + in_off(i0 + nr), in_bytes, &my_bar[ns]);
      }
    }
    Sr0 = nSr0;
    Sr1 = nSr1;
  }
#pragma unroll
  for (int jj = 0; jj < TJ; ++jj) {
    if (jj < nj) {
      const int64_t pc = (int64_t(g.col_part[j0 + jj]) + 2 * rc) * 32;
      float ac0, ac1;
      upk2(ac[jj], ac0, ac1);
      PL[pc] = ac0;
      PL[pc + 32] = ac1;
    }
  }
  if (kDelta && b < batch) publish_delta(a.deltas, int64_t(b) * a.delta_stride + a.delta_off, dmax);
  if (lane == 0) bulk_wait_read<0>();  // shared memory must outlive the pending stores
}

// ---------------------------------------------------------------------------
// K1-fused: S_v = ev_v + (messages of the edges that no fused block covers, in
// ascending message index) + (partial sums written by the fused blocks, in
// ascending partial row).  One thread per (var-state, sample).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_var_reduce(BatchMap mp, int64_t num_var_states, int64_t Es, int64_t part_rows,
             const int32_t* __restrict__ vs_var, const int32_t* __restrict__ var_first_state,
             const int32_t* __restrict__ rest_ptr, const int32_t* __restrict__ rest_edge_msg,
             const int32_t* __restrict__ part_first, const int32_t* __restrict__ part_count, View ev,
             const float* __restrict__ m, const float* __restrict__ part, float* __restrict__ S) {
  UnitLoop L = unit_loop(mp, num_var_states);
  if (!L.b_ok) return;
  const LaneView evL = lane_view(ev, mp, L.b);
  const float* mL = m + lane_off(mp, Es, L.b);
  const float* PL = part + lane_off(mp, part_rows, L.b);
  float* SL = S + lane_off(mp, num_var_states, L.b);
  const int sh = mp.bx_log;
  for (int64_t v = L.u; v < L.u_end; v += L.step) {
    const int var = vs_var[v];
    const int64_t st = v - var_first_state[var];
    float acc = evL.at(v);
    for (int64_t k = rest_ptr[var]; k < rest_ptr[var + 1]; ++k)
      acc += mL[(rest_edge_msg[k] + st) << sh];
    // partial rows of (var, state st): first + 2*k + st (fused blocks hold binary variables)
    const int64_t p0 = part_first[var] + st;
    const int cnt = part_count[var];
    int k = 0;
    for (; k + 4 <= cnt; k += 4) {
      const float a0 = PL[(p0 + 2 * k) << sh];
      const float a1 = PL[(p0 + 2 * (k + 1)) << sh];
      const float a2 = PL[(p0 + 2 * (k + 2)) << sh];
      const float a3 = PL[(p0 + 2 * (k + 3)) << sh];
      acc += a0; acc += a1; acc += a2; acc += a3;
    }
    for (; k < cnt; ++k) acc += PL[(p0 + 2 * k) << sh];
    SL[v << sh] = acc;
  }
}

// Device-side description of one enum block (see pgx_enum_block in pgx.h).
// cfg_es[k*arity + a]: edge-state offset (within the factor's message span) that
// configuration k assigns to variable a.  t_ptr/t_k: for every edge-state offset
// the ascending list of configurations containing it (the transpose of cfg_es,
// which the reference never builds; it scatter-maxes over the R expanded rows).
struct EnumBlockDev {
  int64_t num_factors;
  int64_t first_edge, first_msg, first_pot;
  int32_t arity, num_configs, ns;  // ns = edge-states per factor
  const int32_t* cfg_es;
  const int32_t* t_ptr;
  const int32_t* t_k;
  const int32_t* edge_off;  // [arity + 1]
  // per-factor offsets when the block merges several descriptor blocks (else null and the
  // factors are the arithmetic progression first_* + f * stride)
  const int32_t* fac_edge;
  const int32_t* fac_msg;
  const int32_t* fac_pot;
  // arity 2 and the configurations are sorted by the first variable's state: the configs of
  // state a of variable 0 are the contiguous range [t_ptr[a], t_ptr[a + 1]) of k
  int32_t sorted0;
  __device__ __forceinline__ int64_t msg_base(int64_t f) const { return fac_msg ? fac_msg[f] : first_msg + f * ns; }
  __device__ __forceinline__ int64_t edge_base(int64_t f) const { return fac_edge ? fac_edge[f] : first_edge + f * arity; }
  __device__ __forceinline__ int64_t pot_base(int64_t f) const { return fac_pot ? fac_pot[f] : first_pot + f * num_configs; }
};

// ---------------------------------------------------------------------------
// K2b: EnumFactor update, small factors (ns <= 64): one thread per (factor,
// sample); q staged in a per-thread array, edge-state by edge-state walk of
// the transposed configuration lists.  Exact ascending-config order for both
// the max and the sum.
// ---------------------------------------------------------------------------
template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_enum_small(BatchMap mp, EnumBlockDev blk, const int32_t* __restrict__ edge_vs, View lp,
             const float* __restrict__ S, const float* __restrict__ m_old,
             float* __restrict__ m_new, RunArgs a) {
  UnitLoop L = unit_loop(mp, blk.num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int sh = mp.bx_log;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float* mo = m_old + moff;
  float* mn = m_new + moff;
  const float* SL = S + lane_off(mp, a.Vs, L.b);
  const LaneView lpL = lane_view(lp, mp, L.b);
  const float T = a.T;
  float q[kSmallMaxNS];
  float nv[kSmallMaxNS];
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t mbase = blk.msg_base(f);
    const int64_t ebase = blk.edge_base(f);
    const int64_t pbase = blk.pot_base(f);
    for (int e = 0; e < blk.arity; ++e) {
      const int64_t vs = edge_vs[ebase + e];
      for (int s = blk.edge_off[e]; s < blk.edge_off[e + 1]; ++s)
        q[s] = SL[(vs + s - blk.edge_off[e]) << sh] - mo[(mbase + s) << sh];
    }
    for (int s = 0; s < blk.ns; ++s) {
      const int j0 = blk.t_ptr[s], j1 = blk.t_ptr[s + 1];
      float M = -INFINITY;
      for (int j = j0; j < j1; ++j) {
        const int k = blk.t_k[j];
        float sk = 0.f;
        for (int e = 0; e < blk.arity; ++e) sk += q[blk.cfg_es[k * blk.arity + e]];
        sk += clip_lp(lpL.at(pbase + k));
        M = fmaxf(M, sk);
      }
      float val = M;
      if (kSumProduct) {
        float sum = 0.f;
        for (int j = j0; j < j1; ++j) {
          const int k = blk.t_k[j];
          float sk = 0.f;
          for (int e = 0; e < blk.arity; ++e) sk += q[blk.cfg_es[k * blk.arity + e]];
          sk += clip_lp(lpL.at(pbase + k));
          sum += expf((sk - M) / T);
        }
        val = T * logf(sum) + M;
      }
      nv[s] = damp(mo[(mbase + s) << sh], val - q[s], a.d, a.one_minus_d);
    }
    for (int e = 0; e < blk.arity; ++e) {
      const int s0 = blk.edge_off[e], s1 = blk.edge_off[e + 1];
      float mx = -INFINITY;
      for (int s = s0; s < s1; ++s) mx = fmaxf(mx, nv[s]);
      for (int s = s0; s < s1; ++s) {
        const float out = fmaxf(nv[s] - mx, kMsgNegInf);
        const int64_t idx = (mbase + s) << sh;
        dmax = fmaxf(dmax, fabsf(out - mo[idx]));
        mn[idx] = out;
      }
    }
  }
  publish_delta(a.deltas, int64_t(L.b) * a.delta_stride + a.delta_off, dmax);
}

// ---------------------------------------------------------------------------
// K2c: EnumFactor update, large factors (RCN: 2 x 625 states, up to 375 769
// configurations): one CTA per (factor, sample).  q and the damped values live
// in shared memory; threads own edge-states and walk their configuration lists
// (exact order, no atomics); per-edge max by block reduction.
// Dynamic smem: 2 * ns floats + 32 floats.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float block_max(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float r = -INFINITY;
  for (int i = 0; i < nw; ++i) r = fmaxf(r, red[i]);
  return r;
}

template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_enum_big(BatchMap mp, EnumBlockDev blk, const int32_t* __restrict__ edge_vs, View lp,
           const float* __restrict__ S, const float* __restrict__ m_old,
           float* __restrict__ m_new, RunArgs a) {
  extern __shared__ float smem[];
  float* q = smem;
  float* nv = smem + blk.ns;
  float* red = smem + 2 * blk.ns;
  const int sh = mp.bx_log;
  const float T = a.T;
  const int64_t total = blk.num_factors * mp.batch;
  for (int64_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
    const int64_t f = unit / mp.batch;
    const int b = int(unit - f * mp.batch);
    const int64_t moff = lane_off(mp, a.Es, b);
    const float* mo = m_old + moff;
    float* mn = m_new + moff;
    const float* SL = S + lane_off(mp, a.Vs, b);
    const LaneView lpL = lane_view(lp, mp, b);
    const int64_t mbase = blk.msg_base(f);
    const int64_t ebase = blk.edge_base(f);
    const int64_t pbase = blk.pot_base(f);
    __syncthreads();  // previous unit done with q / nv
    for (int e = 0; e < blk.arity; ++e) {
      const int64_t vs = edge_vs[ebase + e];
      const int s0 = blk.edge_off[e], s1 = blk.edge_off[e + 1];
      for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x)
        q[s] = SL[(vs + s - s0) << sh] - mo[(mbase + s) << sh];
    }
    __syncthreads();
    for (int s = threadIdx.x; s < blk.ns; s += blockDim.x) {
      const int j0 = blk.t_ptr[s], j1 = blk.t_ptr[s + 1];
      float M = -INFINITY;
      for (int j = j0; j < j1; ++j) {
        const int k = blk.t_k[j];
        float sk = 0.f;
        for (int e = 0; e < blk.arity; ++e) sk += q[blk.cfg_es[k * blk.arity + e]];
        sk += clip_lp(lpL.at(pbase + k));
        M = fmaxf(M, sk);
      }
      float val = M;
      if (kSumProduct) {
        float sum = 0.f;
        for (int j = j0; j < j1; ++j) {
          const int k = blk.t_k[j];
          float sk = 0.f;
          for (int e = 0; e < blk.arity; ++e) sk += q[blk.cfg_es[k * blk.arity + e]];
          sk += clip_lp(lpL.at(pbase + k));
          sum += expf((sk - M) / T);
        }
        val = T * logf(sum) + M;
      }
      nv[s] = damp(mo[(mbase + s) << sh], val - q[s], a.d, a.one_minus_d);
    }
    float dmax = 0.f;
    for (int e = 0; e < blk.arity; ++e) {
      const int s0 = blk.edge_off[e], s1 = blk.edge_off[e + 1];
      __syncthreads();
      float mx = -INFINITY;
      for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x) mx = fmaxf(mx, nv[s]);
      mx = block_max(mx, red);
      for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
        const float out = fmaxf(nv[s] - mx, kMsgNegInf);
        const int64_t idx = (mbase + s) << sh;
        dmax = fmaxf(dmax, fabsf(out - mo[idx]));
        mn[idx] = out;
      }
    }
    if (a.deltas != nullptr) {
      dmax = block_max(dmax, red);
      if (threadIdx.x == 0) publish_delta(a.deltas, int64_t(b) * a.delta_stride + a.delta_off, dmax);
    }
  }
}

// ---------------------------------------------------------------------------
// K2c-max: max-product update of large PAIRWISE factors whose configuration table is
// sorted by the first variable's state (RCN lateral factors, examples/rcn.ipynb cell
// 24-26).  Configuration-major: every valid configuration is visited ONCE per
// iteration (the reference visits each twice, through 5 expanded R-sized arrays):
// a warp owns a state a of variable 0, its lanes stride the contiguous config range of
// a (coalesced reads of the table and of the potentials), s_k = (q_a + q_b) + lp_k,
// the max over k for a by warp shuffle, for the partner states b by an ordered-int
// atomicMax in shared memory.  max is order-independent, so the result is bit-identical
// to the edge-state-major kernel and to the oracle.
// Dynamic smem: 2 * ns floats + 32 floats.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_float_shared(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void __launch_bounds__(kThreads)
k_enum_big_maxprod(BatchMap mp, EnumBlockDev blk, const int32_t* __restrict__ edge_vs, View lp,
                   const float* __restrict__ S, const float* __restrict__ m_old,
                   float* __restrict__ m_new, RunArgs a) {
  extern __shared__ float smem[];
  float* q = smem;
  float* M = smem + blk.ns;
  float* red = smem + 2 * blk.ns;
  const int sh = mp.bx_log;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int n0 = blk.edge_off[1];  // states of variable 0
  const int64_t total = blk.num_factors * mp.batch;
  for (int64_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
    const int64_t f = unit / mp.batch;
    const int b = int(unit - f * mp.batch);
    const int64_t moff = lane_off(mp, a.Es, b);
    const float* mo = m_old + moff;
    float* mn = m_new + moff;
    const float* SL = S + lane_off(mp, a.Vs, b);
    const LaneView lpL = lane_view(lp, mp, b);
    const int64_t mbase = blk.msg_base(f), ebase = blk.edge_base(f), pbase = blk.pot_base(f);
    __syncthreads();  // previous unit done with q / M
    for (int e = 0; e < 2; ++e) {
      const int64_t vs = edge_vs[ebase + e];
      const int s0 = blk.edge_off[e], s1 = blk.edge_off[e + 1];
      for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
        q[s] = SL[(vs + s - s0) << sh] - mo[(mbase + s) << sh];
        M[s] = -INFINITY;
      }
    }
    __syncthreads();
    for (int s = warp; s < n0; s += nwarp) {
      const int k0 = blk.t_ptr[s], k1 = blk.t_ptr[s + 1];
      const float qa = q[s];
      float best = -INFINITY;
      for (int k = k0 + lane; k < k1; k += 32) {
        const int es_b = blk.cfg_es[2 * k + 1];
        const float sk = (qa + q[es_b]) + clip_lp(lpL.at(pbase + k));
        best = fmaxf(best, sk);
        atomic_max_float_shared(&M[es_b], sk);
      }
      for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
      if (lane == 0) M[s] = best;
    }
    __syncthreads();
    float dmax = 0.f;
    for (int e = 0; e < 2; ++e) {
      const int s0 = blk.edge_off[e], s1 = blk.edge_off[e + 1];
      float mx = -INFINITY;
      for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
        // f = M - q, damped; M is reused to hold the damped value
        const float nvs = damp(mo[(mbase + s) << sh], M[s] - q[s], a.d, a.one_minus_d);
        M[s] = nvs;
        mx = fmaxf(mx, nvs);
      }
      mx = block_max(mx, red);
      for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
        const float out = fmaxf(M[s] - mx, kMsgNegInf);
        const int64_t idx = (mbase + s) << sh;
        dmax = fmaxf(dmax, fabsf(out - mo[idx]));
        mn[idx] = out;
      }
    }
    if (a.deltas != nullptr) {
      dmax = block_max(dmax, red);
      if (threadIdx.x == 0) publish_delta(a.deltas, int64_t(b) * a.delta_stride + a.delta_off, dmax);
    }
  }
}

// ---------------------------------------------------------------------------
// K2c-max2: the same update without shared-memory atomics (ATOMS on spread addresses costs
// ~2 cycles per LANE on this part, which made the kernel above atomics-bound), for ALL such
// groups of the graph in ONE launch.
//   * LANE-PER-STATE: lane l of lane-group g owns state a = 32 g + l of the first variable and
//     walks a's configuration list; its maximum over the list (the a-side message) is a plain
//     running maximum in a register - no warp reduction, no list logic in the kernel;
//   * the plan arranges the walk in ROUNDS (one configuration per lane) such that the partner
//     states b of a round fall into pairwise distinct shared-memory banks (a lane takes any
//     of its remaining configurations whose bank is free, else idles that round: distinct b
//     AND conflict-free accesses), so every warp keeps a PRIVATE copy Mw[warp][b] of the partner-side maxima and
//     updates it with a plain read-max-write (__syncwarp between rounds); the copies are
//     max-reduced once per factor.  max is order-independent: bit-identical to the other
//     kernels and to the oracle;
//   * one 4-byte schedule entry (k | b << 20, coalesced, L2-resident, shared by all factors of
//     the group) and the 4-byte potential (HBM; a lane streams its own list, so a fetched
//     sector serves its next 8 rounds out of L1) per configuration; ~20 instructions per
//     32 configurations;
//   * loads run one trip (kBigTrip rounds) ahead of their use in registers;
//   * work units (factor, sample) of all groups are sorted by configuration count
//     (descending) and handed out through an atomic counter: the launch ends balanced.
// Dynamic smem: (2 ns + nwarps * (n1 + 32) + 32) floats of the largest group.
// ---------------------------------------------------------------------------
constexpr int kBigWarps = kThreads / 32;
struct BigMaxGroup {
  EnumBlockDev blk;
  const uint32_t* rounds;    // [num_rounds][32] k | partner state << 20, 0xffffffff = idle
  const int32_t* round_ptr;  // [num_groups + 1]
  int32_t num_groups;        // lane-groups = ceil(states of variable 0 / 32)
  // permuted-potential path: the run starts by copying every factor's (clipped) potentials into
  // round order, lpR[perm_base + f * 32 * num_rounds + 32 * round + lane] (-inf at idle
  // entries), so that the hot loop's two loads per configuration - the potential and the
  // 2-byte partner state - are both coalesced and the loop needs no select at all
  const uint32_t* rounds_b;  // [num_rounds / 2][32] partner states of rounds 2p, 2p + 1 (16 bits each);
                             // idle: n1 + j, a dummy slot in a bank no lane of the round uses.  Every lane-group has an
                             // even number of rounds.
  int64_t perm_base;
  int32_t num_rounds;
};

// One-off per run: potentials -> round order (see BigMaxGroup).
template <bool kFlatLp>
__global__ void __launch_bounds__(kThreads)
k_bigmax_permute(BatchMap mp, const BigMaxGroup* __restrict__ groups, const int2* __restrict__ units,
                 int64_t num_units, View lp, float* __restrict__ lpR) {
  for (int64_t u = blockIdx.y; u < num_units; u += gridDim.y) {
    const int2 uf = units[u];
    const BigMaxGroup& G = groups[uf.x];
    const int64_t n = int64_t(G.num_rounds) * 32;
    const LaneView lpL = lane_view(lp, mp, 0);
    const float* __restrict__ src = lpL.q + (G.blk.pot_base(uf.y) << lpL.sh);
    float* __restrict__ dst = lpR + G.perm_base + int64_t(uf.y) * n;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
      const uint32_t e = G.rounds[i];
      dst[i] = e == 0xffffffffu ? -INFINITY : clip_lp(kFlatLp ? src[e & 0xfffffu] : src[size_t(e & 0xfffffu) << lpL.sh]);
    }
  }
}

constexpr int kBigTrip = 8;

// kFlatLp: potentials addressed without a sample-tile shift (shared or batch-major);
// kPerm: potentials come from the round-ordered copy lpR (potentials shared by the batch)
template <bool kFlatLp, bool kPerm>
__global__ void __launch_bounds__(kThreads, 3)
k_enum_big_maxprod_all(BatchMap mp, const BigMaxGroup* __restrict__ groups, const int2* __restrict__ units,
                       int64_t num_units, unsigned int* __restrict__ counter,
                       const int32_t* __restrict__ edge_vs, View lp, const float* __restrict__ lpR,
                       const float* __restrict__ S,
                       const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  extern __shared__ float smem[];
  __shared__ unsigned int s_unit;
  const int sh = mp.bx_log;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t total = num_units * mp.batch;
  for (;;) {
    __syncthreads();  // previous unit done with shared memory (and with s_unit)
    if (threadIdx.x == 0) s_unit = atomicAdd(counter, 1u);
    __syncthreads();
    const int64_t unit = s_unit;
    if (unit >= total) break;
    const int64_t u = unit / mp.batch;
    const int b = int(unit - u * mp.batch);
    const int2 uf = units[u];
    const BigMaxGroup& G = groups[uf.x];
    const EnumBlockDev& blk = G.blk;
    const int64_t f = uf.y;
    const int ns = blk.ns, n0 = blk.edge_off[1], n1 = ns - n0;
    float* q = smem;                       // [ns]
    float* M = q + ns;                     // [ns] maxima, then damped values
    float* Mw = M + ns;                    // [kBigWarps][n1 + 32] per-warp partner-side maxima
    float* red = Mw + kBigWarps * (n1 + 32);
    const int64_t moff = lane_off(mp, a.Es, b);
    const float* mo = m_old + moff;
    float* mn = m_new + moff;
    const float* SL = S + lane_off(mp, a.Vs, b);
    const LaneView lpL = lane_view(lp, mp, b);
    const int64_t mbase = blk.msg_base(f), ebase = blk.edge_base(f), pbase = blk.pot_base(f);
    for (int e = 0; e < 2; ++e) {
      const int64_t vs = edge_vs[ebase + e];
      const int s0 = blk.edge_off[e], s1 = blk.edge_off[e + 1];
      for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
        q[s] = SL[(vs + s - s0) << sh] - mo[(mbase + s) << sh];
        M[s] = -INFINITY;
      }
    }
    for (int i = threadIdx.x; i < kBigWarps * (n1 + 32); i += blockDim.x) Mw[i] = -INFINITY;
    __syncthreads();

    {  // ---- configurations: this warp's lane-groups, round by round ---------------------------
      const uint32_t* __restrict__ rounds = G.rounds;
      const float* __restrict__ lpu = lpL.q + (pbase << lpL.sh);
      const int lsh = lpL.sh;
      const uint32_t qb_s = smem_u32(q) + 4u * n0;
      const uint32_t mw_s = smem_u32(Mw + warp * (n1 + 32));
      const uint32_t dummy_s = mw_s + 4u * (n1 + lane);  // idle lanes read-max-write their own slot
      auto lds_f = [](uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; };
      auto sts_f = [](uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); };
      const float* __restrict__ lpr = lpR + G.perm_base + f * (int64_t(G.num_rounds) * 32) + lane;
      const uint32_t* __restrict__ rbl = G.rounds_b + lane;
      // one trip = kPermTrip rounds; the next trip's loads (2 KiB of potentials per warp) are in
      // flight while this one is consumed: with 24 warps per SM that is ~50 KiB per SM in
      // flight, what HBM latency x bandwidth asks for
      constexpr int kPermTrip = 16;
      for (int grp = warp; kPerm && grp < G.num_groups; grp += kBigWarps) {
        const int a_own = grp * 32 + lane;
        const float qa = a_own < n0 ? q[a_own] : 0.f;
        const int r_end = G.round_ptr[grp + 1];
        const uint32_t idle2 = uint32_t(n1 + lane) * 0x10001u;
        uint32_t en_n[kPermTrip / 2];
        float rl_n[kPermTrip];
        auto request = [&](int r0) {
#pragma unroll
          for (int p = 0; p < kPermTrip / 2; ++p) {
            const bool in = r0 + 2 * p < r_end;  // rounds come in pairs
            en_n[p] = in ? __ldg(rbl + (size_t(r0 + 2 * p) << 4)) : idle2;
            rl_n[2 * p] = in ? __ldcs(lpr + (size_t(r0 + 2 * p) << 5)) : -INFINITY;
            rl_n[2 * p + 1] = in ? __ldcs(lpr + (size_t(r0 + 2 * p + 1) << 5)) : -INFINITY;
          }
        };
        float best = -INFINITY;
        int r0 = G.round_ptr[grp];
        if (r0 < r_end) request(r0);
        for (; r0 < r_end; r0 += kPermTrip) {
          uint32_t en[kPermTrip / 2];
          float rl[kPermTrip];
#pragma unroll
          for (int p = 0; p < kPermTrip / 2; ++p) en[p] = en_n[p];
#pragma unroll
          for (int p = 0; p < kPermTrip; ++p) rl[p] = rl_n[p];
          if (r0 + kPermTrip < r_end) request(r0 + kPermTrip);
          // idle entries: potential -inf, partner "state" n1 + j = a dummy slot
          // (the q read lands in M[], finite or -inf: the sum stays -inf).
          // All q reads of the trip first (read-only: they overlap), then the read-max-write chain.
          uint32_t b_s[kPermTrip];
          float sk[kPermTrip];
#pragma unroll
          for (int p = 0; p < kPermTrip; ++p) {
            b_s[p] = ((p & 1) ? (en[p >> 1] >> 16) : (en[p >> 1] & 0xffffu)) << 2;
            sk[p] = lds_f(qb_s + b_s[p]);
          }
#pragma unroll
          for (int p = 0; p < kPermTrip; ++p) {
            sk[p] = (qa + sk[p]) + rl[p];
            best = fmaxf(best, sk[p]);
          }
#pragma unroll
          for (int p = 0; p < kPermTrip; ++p) {
            const uint32_t slot = mw_s + b_s[p];
            sts_f(slot, fmaxf(lds_f(slot), sk[p]));
            asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");
          }
        }
        if (a_own < n0) M[a_own] = best;
      }
      for (int grp = warp; !kPerm && grp < G.num_groups; grp += kBigWarps) {
        const int a_own = grp * 32 + lane;
        const float qa = a_own < n0 ? q[a_own] : 0.f;
        const int r_end = G.round_ptr[grp + 1];
        uint32_t en_n[kBigTrip];
        float rl_n[kBigTrip];
        auto request = [&](int r0) {
#pragma unroll
          for (int p = 0; p < kBigTrip; ++p)
            en_n[p] = r0 + p < r_end ? __ldg(rounds + (size_t(r0 + p) << 5) + lane) : 0xffffffffu;
#pragma unroll
          for (int p = 0; p < kBigTrip; ++p) {
            // unconditional loads: idle lanes read configuration 0 of the factor
            const uint32_t k = en_n[p] == 0xffffffffu ? 0u : (en_n[p] & 0xfffffu);
            rl_n[p] = kFlatLp ? __ldg(lpu + k) : __ldg(lpu + (size_t(k) << lsh));
          }
        };
        float best = -INFINITY;
        int r0 = G.round_ptr[grp];
        if (r0 < r_end) request(r0);
        for (; r0 < r_end; r0 += kBigTrip) {
          uint32_t en[kBigTrip];
          float rl[kBigTrip];
#pragma unroll
          for (int p = 0; p < kBigTrip; ++p) { en[p] = en_n[p]; rl[p] = rl_n[p]; }
          if (r0 + kBigTrip < r_end) request(r0 + kBigTrip);  // next trip in flight during this one
#pragma unroll
          for (int p = 0; p < kBigTrip; ++p) {
            const bool on = en[p] != 0xffffffffu;            // absent rounds of the last trip are idle entries
            const uint32_t b_s = on ? (en[p] >> 20) << 2 : 0u;  // idle lanes read partner state 0
            float sk = (qa + lds_f(qb_s + b_s)) + clip_lp(rl[p]);
            sk = on ? sk : -INFINITY;
            best = fmaxf(best, sk);
            const uint32_t slot = on ? mw_s + b_s : dummy_s;
            sts_f(slot, fmaxf(lds_f(slot), sk));
            __syncwarp();
          }
        }
        if (a_own < n0) M[a_own] = best;
      }
    }
    __syncthreads();
    for (int s = threadIdx.x; s < n1; s += blockDim.x) {
      float v = Mw[s];
      for (int w = 1; w < kBigWarps; ++w) v = fmaxf(v, Mw[w * (n1 + 32) + s]);
      M[n0 + s] = v;
    }
    __syncthreads();
    float dmax = 0.f;
    for (int e = 0; e < 2; ++e) {
      const int s0 = blk.edge_off[e], s1 = blk.edge_off[e + 1];
      float mx = -INFINITY;
      for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
        const float nvs = damp(mo[(mbase + s) << sh], M[s] - q[s], a.d, a.one_minus_d);
        M[s] = nvs;
        mx = fmaxf(mx, nvs);
      }
      mx = block_max(mx, red);
      for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
        const float out = fmaxf(M[s] - mx, kMsgNegInf);
        const int64_t idx = (mbase + s) << sh;
        dmax = fmaxf(dmax, fabsf(out - mo[idx]));
        mn[idx] = out;
      }
    }
    if (a.deltas != nullptr) {
      dmax = block_max(dmax, red);
      if (threadIdx.x == 0) publish_delta(a.deltas, int64_t(b) * a.delta_stride + a.delta_off, dmax);
    }
  }
}

// ---------------------------------------------------------------------------
// Writes the two states of a binary edge whose factor->variable message is
// (0, x) or (x, 0): damping + normalisation + clip + delta.
//   lo = message index of the edge's state 0; mo / mn are lane pointers.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float write_binary_edge(const float* __restrict__ mo,
                                                   float* __restrict__ mn, int64_t lo, int sh,
                                                   float f0, float f1, float d, float one_minus_d) {
  const int64_t i0 = lo << sh, i1 = (lo + 1) << sh;
  const float a0 = mo[i0], a1 = mo[i1];
  float n0 = damp(a0, f0, d, one_minus_d), n1 = damp(a1, f1, d, one_minus_d);
  const float mx = fmaxf(n0, n1);
  n0 = fmaxf(n0 - mx, kMsgNegInf);
  n1 = fmaxf(n1 - mx, kMsgNegInf);
  mn[i0] = n0;
  mn[i1] = n1;
  return fmaxf(fabsf(n0 - a0), fabsf(n1 - a1));
}

// Device-side logical / pool wiring.  parent_ptr[f]..parent_ptr[f+1] indexes the
// parents of factor f; *_msg are global message indices of the wiring's "p_i" /
// "c" state, *_vs the var-state index of that same state.
struct LogicalDev {
  int64_t num_factors;
  const int32_t* parent_ptr;
  const int32_t* parents_msg;
  const int32_t* parents_vs;
  const int32_t* children_msg;
  const int32_t* children_vs;
  int32_t off;  // +1 OR / Pool, -1 AND
  int32_t uniform;  // > 0: every factor has exactly this many parents (parent_ptr[f] = f * uniform)
};

// ---------------------------------------------------------------------------
// K4: OR / AND update, closed form from per-factor sums and the two largest
// parent differences (pgmax/factor/logical.py:561-779).
// One thread per (factor, sample).  Factors with <= kRegParents parents (the AND
// factors of the deconvolution graphs have 2) keep the parents' variable->factor
// messages in registers: every message is read once.  Wider factors (ORs with up
// to 180 parents) make two passes over the parents, loading kChunk parents' worth
// of independent gathers at a time.
// ---------------------------------------------------------------------------
constexpr int kRegParents = 4;
constexpr int kParentChunk = 8;

// Arithmetic shared by both paths, exactly App. A.3.
struct LogicalAcc {
  float Sb = 0.f, acc = 0.f, d1 = -INFINITY, d2 = -INFINITY;
  int64_t istar = 0;
  template <bool kSumProduct>
  __device__ __forceinline__ void add(int64_t i, float a_i, float b_i, float T) {
    const float dl = a_i - b_i;
    Sb += b_i;
    acc += kSumProduct ? logaddexp_t(a_i, b_i, T) : fmaxf(b_i, a_i);
    if (dl >= d1) { d2 = d1; d1 = dl; istar = i; }  // first arg-max = LARGEST tied index
    else if (dl > d2) d2 = dl;
  }
  template <bool kSumProduct>
  __device__ __forceinline__ float child_relevant(float T) const {
    if (kSumProduct) {
      float CR = logminusexp_t(acc, Sb, T, 1e-4f);
      if (T < kTempStabThre) CR = fmaxf(CR, logaddexp_t(Sb + d1, Sb + d2, T));
      return CR;
    }
    return acc + fminf(0.f, d1);
  }
  // message difference (relevant - other) to parent i
  template <bool kSumProduct>
  __device__ __forceinline__ float parent_out(int64_t i, float a_i, float b_i, float ca, float cb,
                                              float T, bool single) const {
    float PR, PO;
    if (kSumProduct) {
      const float l_i = logaddexp_t(a_i, b_i, T);
      const float Lw = acc - l_i, Sw = Sb - b_i;
      PR = ca + Lw;
      const float o1 = cb + Sw, o2 = ca + Lw, o3 = ca + Sw;
      PO = logminusexp_t(logaddexp_t(o1, o2, T), o3, T, 1e-4f);
      if (T < kTempStabThre) {
        const float bound = (i == istar) ? (Sw + d2) : (Sw + d1);
        PO = fmaxf(PO, logaddexp_t(o1, ca + bound, T));
      }
    } else {
      const float mu = fmaxf(b_i, a_i);
      PR = (acc + ca) - mu;
      const float o1 = (cb + Sb) - b_i;
      const float o2 = PR + ((i == istar) ? fminf(0.f, d2) : fminf(0.f, d1));
      PO = fmaxf(o1, o2);
    }
    if (single) { PR = ca; PO = cb; }  // logical.py:739-757
    return PR - PO;
  }
};

// Groups whose factors all have the same number n <= kRegParents of parents (the AND
// factors of the deconvolution graphs: n = 2).  One thread per (factor, sample), TWO factors
// per iteration: the wiring of both is loaded first, then all their gathers (2 * 2(n + 1)
// message / var-sum pairs in flight), then the closed form of App. A.3 for each.
constexpr int kLogicalUnits = 2;

template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_logical_uniform(BatchMap mp, LogicalDev w, const float* __restrict__ S,
                  const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  UnitLoop L = unit_loop(mp, w.num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int sh = mp.bx_log;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float* mo = m_old + moff;
  float* mn = m_new + moff;
  const float* SL = S + lane_off(mp, a.Vs, L.b);
  const int off = w.off, n = w.uniform;
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  auto write_edge = [&](int64_t pm, float x) {
    const int64_t lo = (off > 0) ? pm : pm - 1;
    dmax = fmaxf(dmax, write_binary_edge(mo, mn, lo, sh, off > 0 ? 0.f : x, off > 0 ? x : 0.f, d, one_minus_d));
  };
  for (int64_t f0 = L.u; f0 < L.u_end; f0 += kLogicalUnits * L.step) {
    int32_t c[kLogicalUnits], cvs[kLogicalUnits], pm[kLogicalUnits][kRegParents], pv[kLogicalUnits][kRegParents];
#pragma unroll
    for (int u = 0; u < kLogicalUnits; ++u) {
      const int64_t f = f0 + u * L.step;
      if (f < L.u_end) {
        c[u] = w.children_msg[f];
        cvs[u] = w.children_vs[f];
#pragma unroll
        for (int j = 0; j < kRegParents; ++j)
          if (j < n) { pm[u][j] = w.parents_msg[f * n + j]; pv[u][j] = w.parents_vs[f * n + j]; }
      }
    }
    float ca[kLogicalUnits], cb[kLogicalUnits], av[kLogicalUnits][kRegParents], bv[kLogicalUnits][kRegParents];
#pragma unroll
    for (int u = 0; u < kLogicalUnits; ++u) {
      if (f0 + u * L.step < L.u_end) {
        ca[u] = SL[int64_t(cvs[u] + off) << sh] - mo[int64_t(c[u] + off) << sh];
        cb[u] = SL[int64_t(cvs[u]) << sh] - mo[int64_t(c[u]) << sh];
#pragma unroll
        for (int j = 0; j < kRegParents; ++j)
          if (j < n) {
            av[u][j] = SL[int64_t(pv[u][j] + off) << sh] - mo[int64_t(pm[u][j] + off) << sh];
            bv[u][j] = SL[int64_t(pv[u][j]) << sh] - mo[int64_t(pm[u][j]) << sh];
          }
      }
    }
#pragma unroll
    for (int u = 0; u < kLogicalUnits; ++u) {
      const int64_t f = f0 + u * L.step;
      if (f < L.u_end) {
        const int64_t p0 = f * n;
        LogicalAcc A;
        A.istar = p0;
#pragma unroll
        for (int j = 0; j < kRegParents; ++j)
          if (j < n) A.add<kSumProduct>(p0 + j, av[u][j], bv[u][j], T);
#pragma unroll
        for (int j = 0; j < kRegParents; ++j)
          if (j < n)
            write_edge(pm[u][j], A.parent_out<kSumProduct>(p0 + j, av[u][j], bv[u][j], ca[u], cb[u], T, n == 1));
        write_edge(c[u], A.child_relevant<kSumProduct>(T) - A.Sb);
      }
    }
  }
  publish_delta(a.deltas, int64_t(L.b) * a.delta_stride + a.delta_off, dmax);
}

// kNarrow: every factor of the group has <= kRegParents parents (the wide path is compiled out,
// which halves the register count and doubles the resident warps).
template <bool kSumProduct, bool kNarrow>
__global__ void __launch_bounds__(kThreads)
k_logical(BatchMap mp, LogicalDev w, const float* __restrict__ S,
          const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  UnitLoop L = unit_loop(mp, w.num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int sh = mp.bx_log;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float* mo = m_old + moff;
  float* mn = m_new + moff;
  const float* SL = S + lane_off(mp, a.Vs, L.b);
  const int off = w.off;
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  // variable->factor message of the wiring's state (at msg index pm / var-state pv) and of
  // the "relevant" state (+off)
  auto q_rel = [&](int64_t pm, int64_t pv) { return SL[(pv + off) << sh] - mo[(pm + off) << sh]; };
  auto q_oth = [&](int64_t pm, int64_t pv) { return SL[pv << sh] - mo[pm << sh]; };
  auto write_edge = [&](int64_t pm, float x) {  // x = message of the "+off" state; the other state gets 0
    const int64_t lo = (off > 0) ? pm : pm - 1;
    dmax = fmaxf(dmax, write_binary_edge(mo, mn, lo, sh, off > 0 ? 0.f : x, off > 0 ? x : 0.f, d, one_minus_d));
  };
  // wiring of one factor; fetched one factor ahead so that its (dependent) index loads
  // overlap the current factor's gathers
  struct Wiring {
    int64_t p0, p1;
    int32_t c, cvs;
    int32_t pm[kRegParents], pv[kRegParents];
  };
  auto load_wiring = [&](int64_t f, Wiring& x) {
    if (w.uniform > 0) { x.p0 = f * w.uniform; x.p1 = x.p0 + w.uniform; }
    else { x.p0 = w.parent_ptr[f]; x.p1 = w.parent_ptr[f + 1]; }
    x.c = w.children_msg[f];
    x.cvs = w.children_vs[f];
    if (x.p1 - x.p0 <= kRegParents) {
#pragma unroll
      for (int j = 0; j < kRegParents; ++j)
        if (x.p0 + j < x.p1) { x.pm[j] = w.parents_msg[x.p0 + j]; x.pv[j] = w.parents_vs[x.p0 + j]; }
    }
  };
  Wiring cur_w, next_w;
  if (L.u < L.u_end) load_wiring(L.u, cur_w);
  for (int64_t f = L.u; f < L.u_end; f += L.step, cur_w = next_w) {
    if (f + L.step < L.u_end) load_wiring(f + L.step, next_w);
    const int64_t p0 = cur_w.p0, p1 = cur_w.p1;
    const int64_t c = cur_w.c, cvs = cur_w.cvs;
    const float ca = q_rel(c, cvs), cb = q_oth(c, cvs);
    const bool single = (p1 - p0) == 1;
    LogicalAcc A;
    A.istar = p0;
    if (kNarrow || p1 - p0 <= kRegParents) {
      float av[kRegParents], bv[kRegParents];
#pragma unroll
      for (int j = 0; j < kRegParents; ++j) {
        if (p0 + j < p1) {
          av[j] = q_rel(cur_w.pm[j], cur_w.pv[j]);
          bv[j] = q_oth(cur_w.pm[j], cur_w.pv[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < kRegParents; ++j)
        if (p0 + j < p1) A.add<kSumProduct>(p0 + j, av[j], bv[j], T);
#pragma unroll
      for (int j = 0; j < kRegParents; ++j)
        if (p0 + j < p1)
          write_edge(cur_w.pm[j], A.parent_out<kSumProduct>(p0 + j, av[j], bv[j], ca, cb, T, single));
    } else if (!kNarrow) {
      // Pass 1: sums in ascending parent order, first / second max of the differences.
      int64_t i = p0;
      for (; i + kParentChunk <= p1; i += kParentChunk) {
        float av[kParentChunk], bv[kParentChunk];
#pragma unroll
        for (int j = 0; j < kParentChunk; ++j) {
          const int64_t pm = w.parents_msg[i + j], pv = w.parents_vs[i + j];
          av[j] = q_rel(pm, pv);
          bv[j] = q_oth(pm, pv);
        }
#pragma unroll
        for (int j = 0; j < kParentChunk; ++j) A.add<kSumProduct>(i + j, av[j], bv[j], T);
      }
      for (; i < p1; ++i) {
        const int64_t pm = w.parents_msg[i], pv = w.parents_vs[i];
        A.add<kSumProduct>(i, q_rel(pm, pv), q_oth(pm, pv), T);
      }
      // Pass 2: outgoing messages to the parents.
      i = p0;
      for (; i + kParentChunk <= p1; i += kParentChunk) {
        int64_t pm[kParentChunk];
        float av[kParentChunk], bv[kParentChunk];
#pragma unroll
        for (int j = 0; j < kParentChunk; ++j) {
          pm[j] = w.parents_msg[i + j];
          const int64_t pv = w.parents_vs[i + j];
          av[j] = q_rel(pm[j], pv);
          bv[j] = q_oth(pm[j], pv);
        }
#pragma unroll
        for (int j = 0; j < kParentChunk; ++j)
          write_edge(pm[j], A.parent_out<kSumProduct>(i + j, av[j], bv[j], ca, cb, T, single));
      }
      for (; i < p1; ++i) {
        const int64_t pm = w.parents_msg[i], pv = w.parents_vs[i];
        write_edge(pm, A.parent_out<kSumProduct>(i, q_rel(pm, pv), q_oth(pm, pv), ca, cb, T, single));
      }
    }
    write_edge(c, A.child_relevant<kSumProduct>(T) - A.Sb);
  }
  publish_delta(a.deltas, int64_t(L.b) * a.delta_stride + a.delta_off, dmax);
}

// ---------------------------------------------------------------------------
// K4-pull: OR / AND update for full sample tiles (TW = 32: a warp = the 32 samples of ONE
// factor, every index warp-uniform) that needs the variable-sum array only for HIGH-degree
// variables.  For a variable with one or two incident edges the kernel re-derives
// S_v = ev_v + (incident messages in ascending message index) itself - the same additions
// in the same order as k_var_sums, so results are bit-identical - which removes the S
// write + gather and the message re-read of k_var_sums for those variables (in the
// deconvolution graphs: every SW and X variable, 2/3 of all var-states); k_var_sums then
// only runs over the listed high-degree var-states (S and W).
// Wiring per edge (EdgeW, one 16-byte load): msg = message index of the state the
// reference's wiring points at (state 0 for OR, state 1 for AND), vs = var-state of that
// state, other = message index (same state) of the variable's only other edge, -1 if the
// variable has no other edge, -2 if its sum is to be read from S.
// All loads of a factor are issued before the first use (no data-dependent branch between
// them); addresses are 32-bit offsets from per-lane bases.
// ---------------------------------------------------------------------------
struct EdgeW {
  int32_t msg, vs, other, pad;
};

struct LogicalPullDev {
  int64_t num_factors;
  const int32_t* parent_ptr;  // [F + 1] (null when uniform)
  const EdgeW* parents;       // [P]
  const EdgeW* children;      // [F]
  int32_t off;                // +1 OR, -1 AND
  int32_t uniform;            // > 0: every factor has this many parents
};

// What stays live per edge between the loads and their use: 8 registers.
struct EdgeIn {
  float m_p, m_r;  // old messages: pointed state, relevant (+off) state
  float a_p, a_r;  // S (kind 0) or evidence
  float o_p, o_r;  // the other edge's messages (kinds 2, 3)
  int32_t msg;
  int32_t kind;    // 0: sums from S; 1: no other edge; 2: other edge first; 3: own edge first
};

// kBin: the message arrays are in binary-difference storage (one float x = n1 - n0 per edge,
// every edge of the graph has two states, edge e holds message rows 2e, 2e + 1; see
// bin_expand): row of an edge = msg >> 1, and (pointed, relevant) = (state 0, state 1) for
// off = +1, (state 1, state 0) for off = -1.
template <bool kBin>
__device__ __forceinline__ size_t msg_rows(const RunArgs& a) { return kBin ? size_t(a.Es) >> 1 : size_t(a.Es); }

// Sum-product closed forms can return an infinite difference (an empty logminusexp); the
// update then leaves BOTH states of the edge at the clip value -1e32 (inf - inf = NaN, and
// fmaxf(NaN, -1e32) = -1e32), the one normalised pair whose maximum is not 0.  The stored
// difference encodes it as NaN (kFloor variants; max-product never produces it).
// load_msg only LOADS (kBin: the raw difference goes to m_p); expand_msg, called on the
// consumer side (edge_q), turns it into the two states - no arithmetic sits between the loads
// of a batch of edges, so they all stay in flight together.
template <bool kBin, bool kFloor>
__device__ __forceinline__ void expand_msg(int off, float& m_p, float& m_r) {
  if (!kBin) return;
  const float x = m_p;
  const float xs = off > 0 ? x : -x;  // relevant - pointed
  m_p = fminf(-xs, 0.f);
  m_r = fminf(xs, 0.f);
  if (kFloor && x != x) m_p = m_r = kMsgNegInf;
}

template <bool kBin, bool kFloor>
__device__ __forceinline__ void load_msg(const float* __restrict__ mo, int32_t msg, int off, float& m_p, float& m_r) {
  if (kBin) {
    m_p = mo[(uint32_t(msg) >> 1) << 5];
    m_r = 0.f;
  } else {
    m_p = mo[uint32_t(msg) << 5];
    m_r = mo[uint32_t(msg + off) << 5];
  }
}

// Issues the (up to) 6 loads of an edge (4 in binary-difference storage); no data-dependent branch.
template <bool kBin, bool kFloor>
__device__ __forceinline__ EdgeIn load_edge(const EdgeW& e, int off, const float* __restrict__ mo,
                                            const float* __restrict__ evq, int esh,
                                            const float* __restrict__ SL) {
  EdgeIn r;
  r.msg = e.msg;
  r.kind = e.other == -2 ? 0 : (e.other == -1 ? 1 : (e.other < e.msg ? 2 : 3));
  load_msg<kBin, kFloor>(mo, e.msg, off, r.m_p, r.m_r);
  const bool from_s = e.other == -2;
  r.a_p = *(from_s ? SL + (uint32_t(e.vs) << 5) : evq + (uint32_t(e.vs) << esh));
  r.a_r = *(from_s ? SL + (uint32_t(e.vs + off) << 5) : evq + (uint32_t(e.vs + off) << esh));
  r.o_p = 0.f;
  r.o_r = 0.f;
  if (e.other >= 0) load_msg<kBin, kFloor>(mo, e.other, off, r.o_p, r.o_r);
  return r;
}
// variable -> factor messages (pointed state, relevant state): S - m with S accumulated from
// the evidence in ascending message index
template <bool kBin, bool kFloor>
__device__ __forceinline__ void edge_q(EdgeIn& r, int off, float& q_p, float& q_r) {
  expand_msg<kBin, kFloor>(off, r.m_p, r.m_r);
  if (r.kind >= 2) expand_msg<kBin, kFloor>(off, r.o_p, r.o_r);
  float s_p = r.a_p, s_r = r.a_r;
  if (r.kind != 0) {
    const bool other_first = r.kind == 2;
    s_p += other_first ? r.o_p : r.m_p;
    s_r += other_first ? r.o_r : r.m_r;
    if (r.kind >= 2) {
      s_p += other_first ? r.m_p : r.o_p;
      s_r += other_first ? r.m_r : r.o_r;
    }
  }
  q_p = s_p - r.m_p;
  q_r = s_r - r.m_r;
}

// new message (x at the relevant state, 0 at the pointed state): damping, normalisation,
// clip, store; returns max|new - old| when kDelta
template <bool kDelta, bool kBin, bool kFloor>
__device__ __forceinline__ float store_edge(float* __restrict__ mn, int off, const EdgeIn& r, float x, float d,
                                            float one_minus_d) {
  float n_p = damp(r.m_p, 0.f, d, one_minus_d), n_r = damp(r.m_r, x, d, one_minus_d);
  const float mx = fmaxf(n_p, n_r);
  n_p = fmaxf(n_p - mx, kMsgNegInf);
  n_r = fmaxf(n_r - mx, kMsgNegInf);
  if (kBin) {  // one of n_p, n_r is the exact zero: the difference loses nothing
    float xd = off > 0 ? n_r - n_p : n_p - n_r;
    if (kFloor && fmaxf(n_p, n_r) < 0.f) xd = __int_as_float(0x7fc00000);  // both states at the floor
    mn[(uint32_t(r.msg) >> 1) << 5] = xd;
  } else {
    mn[uint32_t(r.msg) << 5] = n_p;
    mn[uint32_t(r.msg + off) << 5] = n_r;
  }
  return kDelta ? fmaxf(fabsf(n_p - r.m_p), fabsf(n_r - r.m_r)) : 0.f;
}

// Factors with <= NP parents (AND factors: NP = 2), everything in registers, U factors per
// warp iteration (their loads are all in flight together).  kUniform: every factor has
// exactly NP parents.
template <bool kSumProduct, bool kDelta, int NP, int U, bool kUniform, bool kBin>
__global__ void __launch_bounds__(kThreads)
k_logical_pull_small(int batch, LogicalPullDev w, View ev, const float* __restrict__ S,
                     const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y * 32 + lane;
  if (b >= batch) return;
  const size_t tile = blockIdx.y;
  const float* mo = m_old + tile * msg_rows<kBin>(a) * 32 + lane;
  float* mn = m_new + tile * msg_rows<kBin>(a) * 32 + lane;
  const float* SL = S + tile * size_t(a.Vs) * 32 + lane;
  const float* evq = ev.kind == 1 ? ev.p + tile * size_t(ev.n_rows) * 32 + lane : ev.p;
  const int esh = ev.kind == 1 ? 5 : 0;
  const int off = w.off;
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  const int64_t gwarp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float dmax = 0.f;
  for (int64_t f0 = gwarp; f0 < w.num_factors; f0 += U * nwarps) {
    EdgeIn ce[U], pe[U][NP];
    int np[U];
    int64_t p0[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t f = f0 + u * nwarps;
      np[u] = 0;
      p0[u] = 0;
      if (f < w.num_factors) {
        if (kUniform) { p0[u] = f * NP; np[u] = NP; }
        else { p0[u] = w.parent_ptr[f]; np[u] = int(w.parent_ptr[f + 1] - p0[u]); }
        ce[u] = load_edge<kBin, kSumProduct>(w.children[f], off, mo, evq, esh, SL);
#pragma unroll
        for (int j = 0; j < NP; ++j)
          if (kUniform || j < np[u]) pe[u][j] = load_edge<kBin, kSumProduct>(w.parents[p0[u] + j], off, mo, evq, esh, SL);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (f0 + u * nwarps < w.num_factors) {
        float c_p, c_r, q_p[NP], q_r[NP];
        edge_q<kBin, kSumProduct>(ce[u], off, c_p, c_r);
        LogicalAcc A;
        A.istar = p0[u];
#pragma unroll
        for (int j = 0; j < NP; ++j)
          if (kUniform || j < np[u]) {
            edge_q<kBin, kSumProduct>(pe[u][j], off, q_p[j], q_r[j]);
            A.add<kSumProduct>(p0[u] + j, q_r[j], q_p[j], T);
          }
#pragma unroll
        for (int j = 0; j < NP; ++j)
          if (kUniform || j < np[u]) {
            const float x = A.parent_out<kSumProduct>(p0[u] + j, q_r[j], q_p[j], c_r, c_p, T, np[u] == 1);
            dmax = fmaxf(dmax, store_edge<kDelta, kBin, kSumProduct>(mn, off, pe[u][j], x, d, one_minus_d));
          }
        dmax = fmaxf(dmax, store_edge<kDelta, kBin, kSumProduct>(mn, off, ce[u], A.child_relevant<kSumProduct>(T) - A.Sb, d, one_minus_d));
      }
    }
  }
  if (kDelta) publish_delta(a.deltas, int64_t(b) * a.delta_stride + a.delta_off, dmax);
}

// Any number of parents (OR factors with up to hundreds): two passes over the parents.  The
// wiring of 32 parents is fetched with ONE coalesced load (lane j holds parent i + j) and
// handed out by shuffles; the parents' loads are issued kParentChunk at a time.  All lanes
// stay alive for the shuffles; lanes beyond the batch read a valid sample and store nothing.
template <bool kSumProduct, bool kDelta, bool kBin>
__global__ void __launch_bounds__(128)
k_logical_pull_wide(int batch, LogicalPullDev w, View ev, const float* __restrict__ S,
                    const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  const int lane = threadIdx.x & 31;
  const int b_raw = blockIdx.y * 32 + lane;
  const bool live = b_raw < batch;
  const int ll = live ? lane : 0;  // dead lanes shadow sample 0 of the tile (always valid)
  const size_t tile = blockIdx.y;
  const float* mo = m_old + tile * msg_rows<kBin>(a) * 32 + ll;
  float* mn = m_new + tile * msg_rows<kBin>(a) * 32 + ll;
  const float* SL = S + tile * size_t(a.Vs) * 32 + ll;
  const float* evq = ev.kind == 1 ? ev.p + tile * size_t(ev.n_rows) * 32 + ll : ev.p;
  const int esh = ev.kind == 1 ? 5 : 0;
  const int off = w.off;
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  const int64_t gwarp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float dmax = 0.f;
  for (int64_t f = gwarp; f < w.num_factors; f += nwarps) {
    int64_t p0, p1;
    if (w.uniform > 0) { p0 = f * w.uniform; p1 = p0 + w.uniform; }
    else { p0 = w.parent_ptr[f]; p1 = w.parent_ptr[f + 1]; }
    EdgeIn ce = load_edge<kBin, kSumProduct>(w.children[f], off, mo, evq, esh, SL);
    LogicalAcc A;
    A.istar = p0;
    const bool single = (p1 - p0) == 1;
    float c_p = 0.f, c_r = 0.f;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1) edge_q<kBin, kSumProduct>(ce, off, c_p, c_r);
      EdgeW mine = (p0 + lane < p1) ? w.parents[p0 + lane] : EdgeW{0, 0, -1, 0};
      for (int64_t i32 = p0; i32 < p1; i32 += 32) {
        const EdgeW held = mine;
        if (i32 + 32 + lane < p1) mine = w.parents[i32 + 32 + lane];  // next 32, in flight during this block
        const int n32 = int(min(int64_t(32), p1 - i32));
#pragma unroll 1
        for (int c0 = 0; c0 < n32; c0 += kParentChunk) {
          EdgeIn r[kParentChunk];
#pragma unroll
          for (int j = 0; j < kParentChunk; ++j) {
            EdgeW e;
            e.msg = __shfl_sync(0xffffffffu, held.msg, (c0 + j) & 31);
            e.vs = __shfl_sync(0xffffffffu, held.vs, (c0 + j) & 31);
            e.other = __shfl_sync(0xffffffffu, held.other, (c0 + j) & 31);
            if (c0 + j < n32) r[j] = load_edge<kBin, kSumProduct>(e, off, mo, evq, esh, SL);
          }
#pragma unroll
          for (int j = 0; j < kParentChunk; ++j)
            if (c0 + j < n32) {
              float q_p, q_r;
              edge_q<kBin, kSumProduct>(r[j], off, q_p, q_r);
              const int64_t i = i32 + c0 + j;
              if (pass == 0) {
                A.add<kSumProduct>(i, q_r, q_p, T);
              } else {
                const float x = A.parent_out<kSumProduct>(i, q_r, q_p, c_r, c_p, T, single);
                if (live) dmax = fmaxf(dmax, store_edge<kDelta, kBin, kSumProduct>(mn, off, r[j], x, d, one_minus_d));
              }
            }
        }
      }
    }
    if (live)
      dmax = fmaxf(dmax, store_edge<kDelta, kBin, kSumProduct>(mn, off, ce, A.child_relevant<kSumProduct>(T) - A.Sb, d, one_minus_d));
  }
  if (kDelta && live) publish_delta(a.deltas, int64_t(b_raw) * a.delta_stride + a.delta_off, dmax);
}

// The wide update split in two launches, so that only the part that must be serial is:
//   pass 1 (k_logical_wide_reduce): one warp per (factor, sample tile) walks the parents once,
//     accumulating the sums in ascending parent order and the two largest differences, writes
//     the child's message and the factor's aggregates [F][8][32 samples] (tile-blocked);
//   pass 2 (k_logical_wide_emit): one warp per (PARENT, sample tile) - fully parallel,
//     bandwidth-bound - re-derives its own variable -> factor message and emits the message to
//     the parent from the aggregates.
// Same arithmetic as k_logical_pull_wide (LogicalAcc), hence bit-identical.
constexpr int kAggRows = 8;  // acc, Sb, d1, d2, istar - p0 (int bits), c_p, c_r, unused

template <bool kSumProduct, bool kDelta, bool kBin>
__global__ void __launch_bounds__(32, 22)  // one warp per CTA: a serial chain holds only its own warp
k_logical_wide_reduce(int batch, LogicalPullDev w, View ev, const float* __restrict__ S,
                      const float* __restrict__ m_old, float* __restrict__ m_new, float* __restrict__ agg,
                      RunArgs a) {
  const int lane = threadIdx.x & 31;
  const int b_raw = blockIdx.y * 32 + lane;
  const bool live = b_raw < batch;
  const int ll = live ? lane : 0;
  const size_t tile = blockIdx.y;
  const float* mo = m_old + tile * msg_rows<kBin>(a) * 32 + ll;
  float* mn = m_new + tile * msg_rows<kBin>(a) * 32 + ll;
  const float* SL = S + tile * size_t(a.Vs) * 32 + ll;
  const float* evq = ev.kind == 1 ? ev.p + tile * size_t(ev.n_rows) * 32 + ll : ev.p;
  const int esh = ev.kind == 1 ? 5 : 0;
  float* aggL = agg + tile * size_t(w.num_factors) * kAggRows * 32 + ll;
  const int off = w.off;
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  const int64_t gwarp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float dmax = 0.f;
  for (int64_t f = gwarp; f < w.num_factors; f += nwarps) {
    int64_t p0, p1;
    if (w.uniform > 0) { p0 = f * w.uniform; p1 = p0 + w.uniform; }
    else { p0 = w.parent_ptr[f]; p1 = w.parent_ptr[f + 1]; }
    EdgeIn ce = load_edge<kBin, kSumProduct>(w.children[f], off, mo, evq, esh, SL);
    LogicalAcc A;
    A.istar = p0;
    EdgeW mine = (p0 + lane < p1) ? w.parents[p0 + lane] : EdgeW{0, 0, -1, 0};
    for (int64_t i32 = p0; i32 < p1; i32 += 32) {
      const EdgeW held = mine;
      if (i32 + 32 + lane < p1) mine = w.parents[i32 + 32 + lane];
      const int n32 = int(min(int64_t(32), p1 - i32));
#pragma unroll 1
      for (int c0 = 0; c0 < n32; c0 += kParentChunk) {
        EdgeIn r[kParentChunk];
#pragma unroll
        for (int j = 0; j < kParentChunk; ++j) {
          EdgeW e;
          e.msg = __shfl_sync(0xffffffffu, held.msg, (c0 + j) & 31);
          e.vs = __shfl_sync(0xffffffffu, held.vs, (c0 + j) & 31);
          e.other = __shfl_sync(0xffffffffu, held.other, (c0 + j) & 31);
          if (c0 + j < n32) r[j] = load_edge<kBin, kSumProduct>(e, off, mo, evq, esh, SL);
        }
#pragma unroll
        for (int j = 0; j < kParentChunk; ++j)
          if (c0 + j < n32) {
            float q_p, q_r;
            edge_q<kBin, kSumProduct>(r[j], off, q_p, q_r);
            A.add<kSumProduct>(i32 + c0 + j, q_r, q_p, T);
          }
      }
    }
    float c_p, c_r;
    edge_q<kBin, kSumProduct>(ce, off, c_p, c_r);
    if (live) {
      float* g = aggL + size_t(f) * kAggRows * 32;
      g[0] = A.acc; g[32] = A.Sb; g[64] = A.d1; g[96] = A.d2;
      g[128] = __int_as_float(int(A.istar - p0)); g[160] = c_p; g[192] = c_r;
      dmax = fmaxf(dmax, store_edge<kDelta, kBin, kSumProduct>(mn, off, ce, A.child_relevant<kSumProduct>(T) - A.Sb, d, one_minus_d));
    }
  }
  if (kDelta && live) publish_delta(a.deltas, int64_t(b_raw) * a.delta_stride + a.delta_off, dmax);
}

// parent_factor[i] = factor of parent i (ascending).  kEmitUnits parents per warp iteration.
constexpr int kEmitUnits = 2;

template <bool kSumProduct, bool kDelta, bool kBin>
__global__ void __launch_bounds__(kThreads)
k_logical_wide_emit(int batch, LogicalPullDev w, const int32_t* __restrict__ parent_factor, int64_t num_parents,
                    View ev, const float* __restrict__ S, const float* __restrict__ m_old,
                    float* __restrict__ m_new, const float* __restrict__ agg, RunArgs a) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y * 32 + lane;
  if (b >= batch) return;
  const size_t tile = blockIdx.y;
  const float* mo = m_old + tile * msg_rows<kBin>(a) * 32 + lane;
  float* mn = m_new + tile * msg_rows<kBin>(a) * 32 + lane;
  const float* SL = S + tile * size_t(a.Vs) * 32 + lane;
  const float* evq = ev.kind == 1 ? ev.p + tile * size_t(ev.n_rows) * 32 + lane : ev.p;
  const int esh = ev.kind == 1 ? 5 : 0;
  const float* aggL = agg + tile * size_t(w.num_factors) * kAggRows * 32 + lane;
  const int off = w.off;
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  const int64_t gwarp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float dmax = 0.f;
  for (int64_t i0 = gwarp; i0 < num_parents; i0 += kEmitUnits * nwarps) {
    EdgeIn r[kEmitUnits];
    float g[kEmitUnits][7];
    int64_t p0[kEmitUnits], p1[kEmitUnits];
#pragma unroll
    for (int u = 0; u < kEmitUnits; ++u) {
      const int64_t i = i0 + u * nwarps;
      if (i < num_parents) {
        const int f = parent_factor[i];
        if (w.uniform > 0) { p0[u] = int64_t(f) * w.uniform; p1[u] = p0[u] + w.uniform; }
        else { p0[u] = w.parent_ptr[f]; p1[u] = w.parent_ptr[f + 1]; }
        r[u] = load_edge<kBin, kSumProduct>(w.parents[i], off, mo, evq, esh, SL);
        const float* gp = aggL + size_t(f) * kAggRows * 32;
#pragma unroll
        for (int k = 0; k < 7; ++k) g[u][k] = gp[k * 32];
      }
    }
#pragma unroll
    for (int u = 0; u < kEmitUnits; ++u) {
      const int64_t i = i0 + u * nwarps;
      if (i < num_parents) {
        LogicalAcc A;
        A.acc = g[u][0]; A.Sb = g[u][1]; A.d1 = g[u][2]; A.d2 = g[u][3];
        A.istar = p0[u] + __float_as_int(g[u][4]);
        float q_p, q_r;
        edge_q<kBin, kSumProduct>(r[u], off, q_p, q_r);
        const float x = A.parent_out<kSumProduct>(i, q_r, q_p, g[u][6], g[u][5], T, p1[u] - p0[u] == 1);
        dmax = fmaxf(dmax, store_edge<kDelta, kBin, kSumProduct>(mn, off, r[u], x, d, one_minus_d));
      }
    }
  }
  if (kDelta) publish_delta(a.deltas, int64_t(b) * a.delta_stride + a.delta_off, dmax);
}

// ---------------------------------------------------------------------------
// K5: Pool update (pgmax/factor/pool.py:328-474).
// ---------------------------------------------------------------------------
template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_pool(BatchMap mp, LogicalDev w, const float* __restrict__ S, const float* __restrict__ m_old,
       float* __restrict__ m_new, RunArgs a) {
  UnitLoop L = unit_loop(mp, w.num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int sh = mp.bx_log;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float* mo = m_old + moff;
  float* mn = m_new + moff;
  const float* SL = S + lane_off(mp, a.Vs, L.b);
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  // difference (state 1 - state 0) of the variable->factor message of a binary edge
  auto diff = [&](int64_t pm, int64_t pv) {
    return (SL[(pv + 1) << sh] - mo[(pm + 1) << sh]) - (SL[pv << sh] - mo[pm << sh]);
  };
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t p0 = w.parent_ptr[f], p1 = w.parent_ptr[f + 1];
    const int64_t c = w.children_msg[f], cvs = w.children_vs[f];
    const float D = diff(c, cvs);
    float d1 = -INFINITY, d2 = -INFINITY;
    int64_t istar = p0;
    for (int64_t i = p0; i < p1; ++i) {
      const float dl = diff(w.parents_msg[i], w.parents_vs[i]);
      if (dl >= d1) { d2 = d1; d1 = dl; istar = i; }
      else if (dl > d2) d2 = dl;
    }
    const bool single = (p1 - p0) == 1;
    float out_ind = d1, G = 0.f, out_star = 0.f;
    if (kSumProduct) {
      // logsumexp over the choices with the precomputed max, and over the set where
      // the arg-max choice is replaced by -D (own max), both in ascending order.
      float sum = 0.f, mx2 = -INFINITY;
      for (int64_t i = p0; i < p1; ++i) {
        const float dl = diff(w.parents_msg[i], w.parents_vs[i]);
        sum += expf((dl - d1) / T);
        mx2 = fmaxf(mx2, (i == istar) ? -D : dl);
      }
      out_ind = T * logf(sum) + d1;
      G = logaddexp_t(out_ind, -D, T);
      float sum2 = 0.f;
      for (int64_t i = p0; i < p1; ++i) {
        const float dl = diff(w.parents_msg[i], w.parents_vs[i]);
        sum2 += expf((((i == istar) ? -D : dl) - mx2) / T);
      }
      out_star = -(T * logf(sum2) + mx2);
    }
    for (int64_t i = p0; i < p1; ++i) {
      const int64_t pm = w.parents_msg[i];
      float x;
      if (kSumProduct) {
        const float dl = diff(pm, w.parents_vs[i]);
        x = (i == istar) ? out_star : -logminusexp_t(G, dl, T, 1e-30f);
      } else {
        x = fminf(D, -((i == istar) ? d2 : d1));
      }
      if (single) x = D;  // pool.py:430-450
      dmax = fmaxf(dmax, write_binary_edge(mo, mn, pm, sh, 0.f, x, d, one_minus_d));
    }
    dmax = fmaxf(dmax, write_binary_edge(mo, mn, c, sh, 0.f, out_ind, d, one_minus_d));
  }
  publish_delta(a.deltas, int64_t(L.b) * a.delta_stride + a.delta_off, dmax);
}

// ---------------------------------------------------------------------------
// K6: fused beliefs + MAP decode + marginals + tie count
// (pgmax/infer/inferer.py:218-222,259-264; pgmax/infer/bp.py:283-288).
// One thread per (variable, sample).  Outputs are in the ABI's batch-major
// layout.  beliefs / marginals / map / ties may each be null.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_decode(BatchMap mp, int64_t num_vars, int64_t num_var_states,
         const int32_t* __restrict__ var_first_state, const int32_t* __restrict__ var_ptr,
         const int32_t* __restrict__ var_edge_msg, View ev, View m,
         float* __restrict__ beliefs, int32_t* __restrict__ map_out,
         float* __restrict__ marginals, int32_t* __restrict__ ties) {
  UnitLoop L = unit_loop(mp, num_vars);
  if (!L.b_ok) return;
  const int b = L.b;
  const LaneView evL = lane_view(ev, mp, b), mL = lane_view(m, mp, b);
  int ntie = 0;
  for (int64_t var = L.u; var < L.u_end; var += L.step) {
    const int64_t v0 = var_first_state[var], v1 = var_first_state[var + 1];
    const int64_t k0 = var_ptr[var], k1 = var_ptr[var + 1];
    float best = -INFINITY, second = -INFINITY;
    int arg = 0;
    for (int64_t v = v0; v < v1; ++v) {
      float acc = evL.at(v);
      for (int64_t k = k0; k < k1; ++k) acc += mL.at(var_edge_msg[k] + (v - v0));
      if (beliefs) beliefs[int64_t(b) * num_var_states + v] = acc;
      if (acc > best) { second = best; best = acc; arg = int(v - v0); }
      else if (acc > second) second = acc;
    }
    if (map_out) map_out[int64_t(b) * num_vars + var] = arg;
    if (v1 - v0 >= 2 && best == second) ++ntie;
    if (marginals) {
      // exp(x - logsumexp(x)), logsumexp = max + log sum exp(x - max)
      float sum = 0.f;
      for (int64_t v = v0; v < v1; ++v) {
        float acc = evL.at(v);
        for (int64_t k = k0; k < k1; ++k) acc += mL.at(var_edge_msg[k] + (v - v0));
        sum += expf(acc - best);
      }
      const float lse = best + logf(sum);
      for (int64_t v = v0; v < v1; ++v) {
        float acc = evL.at(v);
        for (int64_t k = k0; k < k1; ++k) acc += mL.at(var_edge_msg[k] + (v - v0));
        marginals[int64_t(b) * num_var_states + v] = expf(acc - lse);
      }
    }
  }
  if (ties != nullptr && ntie > 0) atomicAdd(ties + b, ntie);
}

// ---------------------------------------------------------------------------
// K7: energy of a decoding (pgmax/infer/energy.py:53-148 and the per-type compute_energy:
// factor/enum.py:276-323, logical.py:295-358, pool.py:184-239).  One thread per (unit, sample),
// unit = variable or factor; every launch reduces its units to kEnergyChunks partial sums per
// sample in a fixed order (no float atomics), k_energy_sum adds the partials serially.
//   variable v:        -evidence[v, state(v)]
//   EnumFactor:        -log_potential of THE configuration the decoding selects (unclipped, as
//                      the reference), +inf if no valid configuration matches
//   OR / AND factor:   +inf unless [all parents in the pointed state] == [child in the pointed state]
//   Pool factor:       +inf unless #choices in state 1 == [indicator in state 1]
// map: [batch][num_vars] int32 (ABI order of the variables), or one shared row.
// ---------------------------------------------------------------------------
constexpr int kEnergyChunks = 32;

struct EnergyArgs {
  const int32_t* map;
  int64_t map_stride;       // num_vars, or 0 for a shared decoding
  const float* ev;
  int64_t ev_stride;        // V_s, or 0
  const float* lp;
  int64_t lp_stride;        // C, or 0
  const int32_t* var_first_state;
  const int32_t* vs_var;
  const int32_t* edge_vs;
};

__device__ __forceinline__ void energy_block_reduce(float acc, float* __restrict__ out) {
  __shared__ float red[kThreads];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (int(threadIdx.x) < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

// state the decoding assigns to the variable that owns var-state `vs`, and the state `vs` is
__device__ __forceinline__ void decoded_at(const EnergyArgs& e, const int32_t* mapb, int64_t vs, int& decoded, int& own) {
  const int var = e.vs_var[vs];
  decoded = mapb[var];
  own = int(vs - e.var_first_state[var]);
}

__global__ void __launch_bounds__(kThreads)
k_energy_vars(EnergyArgs e, int64_t num_vars, float* __restrict__ partial, int slots) {
  const int b = blockIdx.y;
  const int32_t* mapb = e.map + int64_t(b) * e.map_stride;
  const float* evb = e.ev + int64_t(b) * e.ev_stride;
  float acc = 0.f;
  for (int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; v < num_vars; v += int64_t(gridDim.x) * blockDim.x) {
    const int64_t v0 = e.var_first_state[v];
    if (e.var_first_state[v + 1] > v0) acc -= evb[v0 + mapb[v]];
  }
  energy_block_reduce(acc, partial + (int64_t(b) * slots) * kEnergyChunks + blockIdx.x);
}

__global__ void __launch_bounds__(kThreads)
k_energy_enum(EnergyArgs e, EnumBlockDev blk, float* __restrict__ partial, int slots, int slot) {
  const int b = blockIdx.y;
  const int32_t* mapb = e.map + int64_t(b) * e.map_stride;
  const float* lpb = e.lp + int64_t(b) * e.lp_stride;
  float acc = 0.f;
  for (int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; f < blk.num_factors; f += int64_t(gridDim.x) * blockDim.x) {
    const int64_t ebase = blk.edge_base(f);
    // configurations that contain the decoded state of the first variable, ascending k
    int dec, own;
    decoded_at(e, mapb, e.edge_vs[ebase], dec, own);
    float en = INFINITY;
    const int n0 = blk.edge_off[1] - blk.edge_off[0];
    if (dec >= 0 && dec < n0) {
      const int j_end = blk.t_ptr[dec + 1];
      for (int j = blk.t_ptr[dec]; j < j_end; ++j) {
        const int k = blk.t_k[j];
        bool match = true;
        for (int a = 1; a < blk.arity && match; ++a) {
          int dec_a;
          decoded_at(e, mapb, e.edge_vs[ebase + a], dec_a, own);
          match = blk.cfg_es[int64_t(k) * blk.arity + a] == blk.edge_off[a] + dec_a;
        }
        if (match) { en = -lpb[blk.pot_base(f) + k]; break; }
      }
    }
    acc += en;
  }
  energy_block_reduce(acc, partial + (int64_t(b) * slots + slot) * kEnergyChunks + blockIdx.x);
}

template <bool kPool>
__global__ void __launch_bounds__(kThreads)
k_energy_logical(EnergyArgs e, LogicalDev w, float* __restrict__ partial, int slots, int slot) {
  const int b = blockIdx.y;
  const int32_t* mapb = e.map + int64_t(b) * e.map_stride;
  float acc = 0.f;
  for (int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; f < w.num_factors; f += int64_t(gridDim.x) * blockDim.x) {
    const int64_t p0 = w.uniform > 0 ? f * w.uniform : w.parent_ptr[f];
    const int64_t p1 = w.uniform > 0 ? p0 + w.uniform : w.parent_ptr[f + 1];
    int dec, own;
    decoded_at(e, mapb, w.children_vs[f], dec, own);
    int lhs;
    if (kPool) {  // children_vs / parents_vs point at state 0: count the choices in state 1
      const int child = dec == own + 1;
      lhs = 0;
      for (int64_t i = p0; i < p1; ++i) {
        decoded_at(e, mapb, w.parents_vs[i], dec, own);
        lhs += dec == own + 1;
      }
      acc += lhs == child ? 0.f : INFINITY;
    } else {
      const int child = dec == own;
      lhs = 1;
      for (int64_t i = p0; i < p1; ++i) {
        decoded_at(e, mapb, w.parents_vs[i], dec, own);
        lhs &= dec == own;
      }
      acc += lhs == child ? 0.f : INFINITY;
    }
  }
  energy_block_reduce(acc, partial + (int64_t(b) * slots + slot) * kEnergyChunks + blockIdx.x);
}

__global__ void k_energy_sum(const float* __restrict__ partial, int n, int64_t batch, float* __restrict__ out) {
  const int64_t b = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  float acc = 0.f;
  for (int i = 0; i < n; ++i) acc += partial[b * n + i];
  out[b] = acc;
}

}  // namespace pgx
