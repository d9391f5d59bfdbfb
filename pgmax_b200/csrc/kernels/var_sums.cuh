// pgx kernels - K1: variable sums.  Part of pgx_kernels.cuh (included in this order).
#pragma once

#include "common.cuh"

namespace pgx {

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

// K1-list for VERY high-degree variables (the W variables of the deconvolution model: 529 edges
// each) on binary-difference storage: a CTA per (variable, sample tile).  A single warp walking such
// a list pays the memory latency once per 32 edges, 17 times in a row; here the CTA's warps gather
// kVsBigSuper edges' rows into shared memory at once (every load of the super-chunk in flight) and
// warp 0 then adds them from shared memory in ascending message index - the same additions in the
// same order, hence bit-identical to k_var_sums_list_bin.
constexpr int kVsBigWarps = 8;
constexpr int kVsBigSuper = kVsBigWarps * 32;  // edges per super-chunk
constexpr int kVsBigDegree = 96;               // variables with at least this many edges take this kernel

__global__ void __launch_bounds__(kVsBigWarps * 32)
k_var_sums_big_bin(int batch, int nbt, int64_t E, int64_t Vs, const int2* __restrict__ vs_csr,
                   const int32_t* __restrict__ var_edge_msg, const int32_t* __restrict__ list, View ev,
                   const float* __restrict__ c, float* __restrict__ S) {
  __shared__ float xs[kVsBigSuper][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile_i = int(blockIdx.x % unsigned(nbt));
  const int64_t i = blockIdx.x / unsigned(nbt);
  const bool live = tile_i * 32 + lane < batch;
  const int ll = live ? lane : 0;
  const size_t tile = tile_i;
  const float* cL = c + tile * size_t(E) * 32 + ll;
  float* SL = S + tile * size_t(Vs) * 32 + ll;
  const float* evq = ev.kind == 1 ? ev.p + tile * size_t(ev.n_rows) * 32 + ll : ev.p;
  const int esh = ev.kind == 1 ? 5 : 0;
  const int v = list[2 * i];  // var-state of state 0; state 1 is v + 1
  const int2 r = vs_csr[v];
  const int k1 = r.x + (r.y >> kVsStateBits);
  float acc0 = 0.f, acc1 = 0.f;
  if (warp == 0) { acc0 = evq[uint32_t(v) << esh]; acc1 = evq[uint32_t(v + 1) << esh]; }
  for (int k0 = r.x; k0 < k1; k0 += kVsBigSuper) {
    // warp w gathers edges k0 + 32 w .. + 31: one coalesced index load, 32 row loads in flight
    const int kb = k0 + warp * 32;
    const int mine = kb + lane < k1 ? var_edge_msg[kb + lane] : 0;
    float x[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int idx = __shfl_sync(0xffffffffu, mine, j);
      x[j] = (kb + j < k1) ? cL[(uint32_t(idx) >> 1) << 5] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) xs[warp * 32 + j][lane] = x[j];
    __syncthreads();
    if (warp == 0) {
      const int n = min(kVsBigSuper, k1 - k0);
      for (int j = 0; j < n; ++j) {
        const float xv = xs[j][lane];
        const bool fl = xv != xv;  // both states at the floor (load_msg)
        acc0 += fl ? kMsgNegInf : fminf(-xv, 0.f);
        acc1 += fl ? kMsgNegInf : fminf(xv, 0.f);
      }
    }
    __syncthreads();
  }
  if (warp == 0 && live) {
    SL[uint32_t(v) << 5] = acc0;
    SL[uint32_t(v + 1) << 5] = acc1;
  }
}

// K1-bin: the sums of ALL variables of a graph whose variables are all binary, messages in
// binary-difference storage (k_enum_pw2_bin): thread per (variable, sample), both states' sums in
// one walk, every stored difference read once.  The same additions in the same order as
// k_var_sums on the expanded messages: bit-identical.
__global__ void __launch_bounds__(kThreads)
k_var_sums_bin(BatchMap mp, int64_t num_vars, int64_t E, int64_t Vs, const int2* __restrict__ vs_csr,
               const int32_t* __restrict__ var_edge_msg, View ev, const float* __restrict__ c,
               float* __restrict__ S) {
  UnitLoop L = unit_loop(mp, num_vars);
  if (!L.b_ok) return;
  const LaneView evL = lane_view(ev, mp, L.b);
  const float* cL = c + lane_off(mp, E, L.b);
  float* SL = S + lane_off(mp, Vs, L.b);
  auto add = [](float x, float& acc0, float& acc1) {
    const bool fl = x != x;  // both states at the floor
    acc0 += fl ? kMsgNegInf : fminf(-x, 0.f);
    acc1 += fl ? kMsgNegInf : fminf(x, 0.f);
  };
  for (int64_t u0 = L.u; u0 < L.u_end; u0 += kVsUnits * L.step) {
    int2 row[kVsUnits];  // (begin, end) of the variable's incident edges
    bool low = true;
#pragma unroll
    for (int u = 0; u < kVsUnits; ++u) {
      const int64_t v = 2 * (u0 + u * L.step);
      const int2 r = u0 + u * L.step < L.u_end ? vs_csr[v] : make_int2(0, 0);
      row[u] = make_int2(r.x, r.x + (r.y >> kVsStateBits));
      low = low && (row[u].y - row[u].x <= kVsLowDeg);
    }
    if (low) {
      float a0[kVsUnits], a1[kVsUnits], x[kVsUnits][kVsLowDeg];
#pragma unroll
      for (int u = 0; u < kVsUnits; ++u) {
        const int64_t v = 2 * (u0 + u * L.step);
        const bool on = u0 + u * L.step < L.u_end;
        a0[u] = on ? evL.at(v) : 0.f;
        a1[u] = on ? evL.at(v + 1) : 0.f;
#pragma unroll
        for (int j = 0; j < kVsLowDeg; ++j)
          x[u][j] = (row[u].x + j < row[u].y) ? cL[int64_t(var_edge_msg[row[u].x + j] >> 1) << 5] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kVsUnits; ++u) {
#pragma unroll
        for (int j = 0; j < kVsLowDeg; ++j)
          if (row[u].x + j < row[u].y) add(x[u][j], a0[u], a1[u]);
        const int64_t v = 2 * (u0 + u * L.step);
        if (u0 + u * L.step < L.u_end) { SL[v << 5] = a0[u]; SL[(v + 1) << 5] = a1[u]; }
      }
      continue;
    }
#pragma unroll 1
    for (int u = 0; u < kVsUnits; ++u) {
      if (u0 + u * L.step >= L.u_end) break;
      const int64_t v = 2 * (u0 + u * L.step);
      float acc0 = evL.at(v), acc1 = evL.at(v + 1);
      int64_t k = row[u].x;
      const int64_t k1 = row[u].y;
      for (; k + 16 <= k1; k += 16) {  // independent loads 16 at a time, added in ascending order
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = cL[int64_t(var_edge_msg[k + j] >> 1) << 5];
#pragma unroll
        for (int j = 0; j < 16; ++j) add(x[j], acc0, acc1);
      }
      for (; k < k1; ++k) add(cL[int64_t(var_edge_msg[k] >> 1) << 5], acc0, acc1);
      SL[v << 5] = acc0;
      SL[(v + 1) << 5] = acc1;
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

}  // namespace pgx
