// pgx kernels - K2a-lattice: index-free 2-D lattice kernels.  Part of pgx_kernels.cuh (included in this order).
#pragma once

#include "pairwise.cuh"

namespace pgx {

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

}  // namespace pgx
