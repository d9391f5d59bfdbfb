// pgx kernels - layout of the batch axis, views, update_utils, layout conversions.  Part of pgx_kernels.cuh (included in this order).
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
constexpr int kTailMaxSamples = 8;     // a batch tail of at most this many samples runs beside the full tiles (pgx.cu)

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

// One sample, wide edges (RCN: 625 states): a warp per edge, lanes stride the edge's contiguous
// states (coalesced; the thread-per-edge kernel above walks them with a stride of one edge
// between neighbouring threads).  max is order-independent: same values.
__global__ void __launch_bounds__(kThreads)
k_normalize_edges_warp(int64_t num_edges, const int32_t* __restrict__ edge_msg_start, float* __restrict__ m) {
  const int lane = threadIdx.x & 31;
  const int64_t gwarp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t e = gwarp; e < num_edges; e += nwarps) {
    const int64_t s0 = edge_msg_start[e], s1 = edge_msg_start[e + 1];
    float mx = -INFINITY;
    for (int64_t s = s0 + lane; s < s1; s += 32) mx = fmaxf(mx, m[s]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    for (int64_t s = s0 + lane; s < s1; s += 32) m[s] = fmaxf(m[s] - mx, kMsgNegInf);
  }
}

}  // namespace pgx
