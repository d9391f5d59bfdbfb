// pgx kernels - K2a-lattice on binary-difference storage, with row segments and halo inputs
// (single-GPU Ising 8192^2 and the row strips of BASELINE.json configs[4]).
// Part of pgx_kernels.cuh (included in this order).
#pragma once

#include "dense_grid.cuh"

namespace pgx {

// ---------------------------------------------------------------------------
// The lattice of lattice.cuh (variable (l, j) owns the vertical factor to (l + 1, j) and the
// horizontal factor to (l, j + 1) mod N) with its messages held in binary-difference storage
// (dense_grid.cuh: one float x = n1 - n0 per two-state edge, lossless after normalisation):
// ONE float4 per cell,
//     c[l * N + j] = (V.a, V.b, H.a, H.b)      a = the owning variable's edge, b = the other end,
// instead of eight floats.  Per iteration HBM sees 16 B of messages in and 16 B out per cell,
// 32 B of potentials and 8 B of evidence: 72 B per cell = 9 B per edge-state (the full layout of
// k_lattice_stream moves 13, SURVEY 8(d)'s algorithmic count is 17).  Values are bit-identical
// to the full layout: every sum adds the expanded pair (min(-x, 0), min(x, 0)), one of which is
// the exact zero, in the same order (evidence first, then ascending message index, wrap-around
// neighbours last), and the update is pw2_update_bin (same operations and roundings as
// pw2_update followed by n1 - n0).
//
// Same persistent, warp-specialised structure as k_lattice_stream (one CTA per SM, a producer
// warp moving tiles with TMA bulk copies on mbarriers, 16 consumer warps updating in place in
// shared memory), plus what the row strips need:
//   * row segments: a launch updates the owner rows [seg_begin[s], seg_end[s]), s = 0, 1, only
//     (a strip's interior rows in one launch while the halo exchange is in flight, its first
//     and last row in a second launch once the halo has landed);
//   * up_add: [2 N] values added to the evidence of row 0 (torus = 0): the messages into the
//     strip's first row from the vertical factors of the strip above;
//   * torus = 0: a ghost row R below the last owner row, whose evidence (ghost_ev, or row R of
//     the evidence array) is the partial variable sum of the strip below's first row.
// ---------------------------------------------------------------------------
struct LatticeBinArgs {
  int32_t R, N;        // owner rows, columns
  int32_t torus;       // 1: rows wrap; 0: ghost row R below the last owner row (receives only)
  int32_t seg_begin[2], seg_end[2];  // owner-row ranges this launch updates (empty: begin >= end)
  const float* up_add; // [2 N] or null
  const float* ghost_ev;  // [2 N] evidence of the ghost row (torus = 0), or null: row R of the evidence array
};

template <int TR_, int TC_, int STAGES_>
struct LbCfg {
  static constexpr int TR = TR_, TC = TC_, kStages = STAGES_;
  static constexpr int MR = TR + 2, MC = TC + 2;
  static constexpr int kMsgF4 = MR * MC;        // float4 per message stage (one per cell)
  static constexpr int kLpF4 = TR * TC * 2;     // float4 per potential stage (two factors per cell)
  static constexpr int kSumF2 = (TR + 1) * (TC + 1);
  static constexpr size_t smem_bytes() {
    return size_t(kStages) * (kMsgF4 + kLpF4) * sizeof(float4) + size_t(2) * kSumF2 * sizeof(float2) +
           2 * kStages * sizeof(uint64_t);
  }
};
constexpr int kLbConsumers = 512;              // 16 warps
constexpr int kLbThreads = kLbConsumers + 32;  // + producer warp

__device__ __forceinline__ f32x2 bin_pair(float x) { return pk2(fminf(-x, 0.f), fminf(x, 0.f)); }

// Tile i of this CTA -> (first owner row, rows of the tile that are updated, first column).
struct LbTile {
  int l0, rows, j0;
};
template <class Cfg>
__device__ __forceinline__ LbTile lb_tile(const LatticeBinArgs& g, int64_t tile, int tiles_x, int ty0) {
  const int ty = int(tile / tiles_x), tx = int(tile - int64_t(ty) * tiles_x);
  LbTile t;
  const int s = ty < ty0 ? 0 : 1;
  const int tyl = s == 0 ? ty : ty - ty0;
  t.l0 = g.seg_begin[s] + tyl * Cfg::TR;
  t.rows = min(Cfg::TR, g.seg_end[s] - t.l0);
  t.j0 = tx * Cfg::TC;
  return t;
}

template <bool kSumProduct, bool kDelta, class Cfg>
__global__ void __launch_bounds__(kLbThreads, 1)
k_lattice_bin(LatticeBinArgs g, const float* __restrict__ ev, const float* __restrict__ lp,
              const float4* __restrict__ c_old, float4* __restrict__ c_new, RunArgs a) {
  constexpr int TR = Cfg::TR, TC = Cfg::TC, MC = Cfg::MC, MR = Cfg::MR, kStages = Cfg::kStages;
  extern __shared__ __align__(128) unsigned char lb_raw[];
  float4* msg_s = reinterpret_cast<float4*>(lb_raw);                         // [stage][MR][MC]
  float4* lp_s = msg_s + kStages * Cfg::kMsgF4;                              // [stage][TR][TC][2]
  float2* sum_s = reinterpret_cast<float2*>(lp_s + kStages * Cfg::kLpF4);    // [2][TR + 1][TC + 1]
  uint64_t* full = reinterpret_cast<uint64_t*>(sum_s + 2 * Cfg::kSumF2);     // [stage] tile landed
  uint64_t* done = full + kStages;                                           // [stage] tile updated
  const int R = g.R, N = g.N;
  const bool torus = g.torus != 0;
  const int tiles_x = (N + TC - 1) / TC;
  const int ty0 = max(0, (g.seg_end[0] - g.seg_begin[0] + TR - 1) / TR);
  const int ty1 = max(0, (g.seg_end[1] - g.seg_begin[1] + TR - 1) / TR);
  const int64_t num_tiles = int64_t(tiles_x) * (ty0 + ty1);
  const int64_t my_tiles = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&done[s], kLbConsumers);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const float4* lp4 = reinterpret_cast<const float4*>(lp);

  if (threadIdx.x >= kLbConsumers) {
    // ------------------------------- producer warp ---------------------------------------
    const int lane = threadIdx.x & 31;
    auto load_tile = [&](int64_t i) {
      const LbTile t = lb_tile<Cfg>(g, blockIdx.x + i * gridDim.x, tiles_x, ty0);
      const int l0 = t.l0, j0 = t.j0;
      const int stage = int(i % kStages);
      float4* ms = msg_s + stage * Cfg::kMsgF4;
      float4* ls = lp_s + stage * Cfg::kLpF4;
      const int ncols = min(TC + 1, N - j0);        // cells from column j0 on (incl. the right halo if inside)
      const bool wrap_right = j0 + TC >= N;         // right halo of the last valid column is column 0
      const int lcols = min(TC, N - j0);
      // rows: sm row rr <-> lattice row l0 - 1 + rr; only rows up to the tile's last updated row + 1
      int l = l0 - 1 + lane;
      bool row_ok = lane < MR && lane <= t.rows + 1;
      if (torus) { row_ok = row_ok && l <= R; l = l < 0 ? R - 1 : (l == R ? 0 : l); }
      else row_ok = row_ok && l >= 0 && l < R;      // the ghost row owns no factor
      const bool lp_ok = lane < t.rows;
      const uint32_t row_bytes = uint32_t(16 + ncols * 16 + (wrap_right ? 16 : 0));
      const uint32_t my_bytes = (row_ok ? row_bytes : 0u) + (lp_ok ? uint32_t(lcols) * 32u : 0u);
      uint32_t total = my_bytes;
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
      if (lane == 0) mbar_expect_tx(&full[stage], total);
      __syncwarp();
      if (row_ok) {
        const float4* src = c_old + int64_t(l) * N;
        float4* dst = ms + lane * MC;
        bulk_g2s(dst, src + (j0 == 0 ? N - 1 : j0 - 1), 16, &full[stage]);
        bulk_g2s(dst + 1, src + j0, uint32_t(ncols) * 16u, &full[stage]);
        if (wrap_right) bulk_g2s(dst + 1 + ncols, src, 16, &full[stage]);
      }
      if (lp_ok)
        bulk_g2s(ls + lane * TC * 2, lp4 + (int64_t(l0 + lane) * N + j0) * 2, uint32_t(lcols) * 32u, &full[stage]);
    };
    for (int s = 0; s < kStages - 1; ++s)
      if (s < my_tiles) load_tile(s);
    for (int64_t i = 0; i < my_tiles; ++i) {
      // the stage of tile i - 1 has been drained below: refill it with tile i + kStages - 1
      if (i + kStages - 1 < my_tiles) load_tile(i + kStages - 1);
      const int stage = int(i % kStages);
      mbar_wait(&done[stage], uint32_t(i / kStages) & 1u);
      const LbTile t = lb_tile<Cfg>(g, blockIdx.x + i * gridDim.x, tiles_x, ty0);
      const int lcols = min(TC, N - t.j0);
      if (lane < t.rows) {
        bulk_s2g(c_new + int64_t(t.l0 + lane) * N + t.j0, msg_s + stage * Cfg::kMsgF4 + (lane + 1) * MC + 1,
                 uint32_t(lcols) * 16u);
        bulk_commit();
      }
      bulk_wait_read<0>();  // this stage's shared memory may be overwritten from here on
      __syncwarp();
    }
    return;
  }

  // --------------------------------- consumer warps -----------------------------------------
  const float2* ev2 = reinterpret_cast<const float2*>(ev);
  const float2* up2 = reinterpret_cast<const float2*>(g.up_add);
  const float2* ghost2 = reinterpret_cast<const float2*>(g.ghost_ev);
  const RunArgs2 c2 = make_args2(a);
  float dmax = 0.f;
  for (int64_t i = 0; i < my_tiles; ++i) {
    const LbTile t = lb_tile<Cfg>(g, blockIdx.x + i * gridDim.x, tiles_x, ty0);
    const int l0 = t.l0, j0 = t.j0;
    const int stage = int(i % kStages);
    float4* sm = msg_s + stage * Cfg::kMsgF4;
    const float4* lq = lp_s + stage * Cfg::kLpF4;
    float2* Ss = sum_s + (i & 1) * Cfg::kSumF2;
    // evidence of this thread's variables: requested before the wait
    constexpr int kVars = (Cfg::kSumF2 + kLbConsumers - 1) / kLbConsumers;
    float2 e[kVars];
#pragma unroll
    for (int k = 0; k < kVars; ++k) {
      const int tt = threadIdx.x + k * kLbConsumers;
      const int rr = tt / (TC + 1), cc = tt - rr * (TC + 1);
      int l = l0 + rr, j = j0 + cc;
      e[k] = make_float2(0.f, 0.f);
      if (tt < Cfg::kSumF2 && rr <= t.rows && l <= R && j <= N) {
        if (j == N) j = 0;
        if (torus && l == R) l = 0;
        e[k] = (ghost2 != nullptr && l == R) ? __ldg(ghost2 + j) : __ldg(ev2 + (int64_t(l) * N + j));
        if (up2 != nullptr && l == 0 && !torus) {  // halo: messages from the strip above
          const float2 u = __ldg(up2 + j);
          e[k].x += u.x;
          e[k].y += u.y;
        }
      }
    }
    mbar_wait(&full[stage], uint32_t(i / kStages) & 1u);
    // ---- variable sums (same order as k_lattice) -------------------------------------------
#pragma unroll
    for (int k = 0; k < kVars; ++k) {
      const int tt = threadIdx.x + k * kLbConsumers;
      const int rr = tt / (TC + 1), cc = tt - rr * (TC + 1);
      int l = l0 + rr, j = j0 + cc;
      if (tt >= Cfg::kSumF2 || rr > t.rows || l > R || j > N) continue;
      if (j == N) j = 0;
      if (torus && l == R) l = 0;
      const bool has_own = l < R;
      const bool has_up = torus || l > 0;
      const bool up_wrap = torus && l == 0;
      const bool left_wrap = j == 0;
      // cell of column j0 + cc sits at sm column cc + 1, except the wrapped right halo
      const int col = (j0 + cc == N) ? (N - j0) + 1 : cc + 1;
      const float4 own = sm[(rr + 1) * MC + col];
      const float up = sm[rr * MC + col].y;                // V factor of the row above: its b edge
      const float left = sm[(rr + 1) * MC + col - 1].w;    // H factor of the left neighbour: its b edge
      f32x2 s = pk2(e[k].x, e[k].y);
      if (has_up && !up_wrap) s = add2(s, bin_pair(up));
      if (has_own && !left_wrap) s = add2(s, bin_pair(left));
      if (has_own) { s = add2(s, bin_pair(own.x)); s = add2(s, bin_pair(own.z)); }
      if (has_own && left_wrap) s = add2(s, bin_pair(left));
      if (up_wrap) s = add2(s, bin_pair(up));
      float s0, s1;
      upk2(s, s0, s1);
      Ss[tt] = make_float2(s0, s1);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kLbConsumers) : "memory");
    // ---- the two factors of every cell, in place ---------------------------------------------
    constexpr int kCells = TR * TC / kLbConsumers;
    static_assert(TR * TC % kLbConsumers == 0, "tile cells must divide over the consumer threads");
#pragma unroll
    for (int k = 0; k < kCells; ++k) {
      const int cell = threadIdx.x + k * kLbConsumers;
      const int rr = cell / TC, cc = cell - rr * TC;
      if (rr < t.rows && j0 + cc < N) {
        const float4 lv = lq[cell * 2], lh = lq[cell * 2 + 1];
        float4* slot = sm + (rr + 1) * MC + cc + 1;
        const float4 x = *slot;
        const float2 sa = Ss[rr * (TC + 1) + cc];
        const float2 sv = Ss[(rr + 1) * (TC + 1) + cc];
        const float2 sh = Ss[rr * (TC + 1) + cc + 1];
        const f32x2 Sa = pk2(sa.x, sa.y);
        float4 o;
        f32x2 na, nb;
        dmax = fmaxf(dmax, pw2_update_bin<kSumProduct, kDelta>(
                               x.x, x.y, Sa, pk2(sv.x, sv.y), pk2(clip_lp(lv.x), clip_lp(lv.y)),
                               pk2(clip_lp(lv.z), clip_lp(lv.w)), c2, o.x, o.y, na, nb));
        dmax = fmaxf(dmax, pw2_update_bin<kSumProduct, kDelta>(
                               x.z, x.w, Sa, pk2(sh.x, sh.y), pk2(clip_lp(lh.x), clip_lp(lh.y)),
                               pk2(clip_lp(lh.z), clip_lp(lh.w)), c2, o.z, o.w, na, nb));
        *slot = o;
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

// Full layout (8 floats per cell, any values) -> normalised binary differences (float4 per
// cell): normalize_and_clip_msgs (bp.py:249-259) and the compression in one pass; for
// already-normalised input the normalisation is the identity.
__global__ void __launch_bounds__(kThreads)
k_lattice_compress(const float4* __restrict__ m, float4* __restrict__ c, int64_t cells) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < cells; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 v = m[2 * i], h = m[2 * i + 1];
    auto diff = [](float m0, float m1) {
      const float mx = fmaxf(m0, m1);
      return fmaxf(m1 - mx, kMsgNegInf) - fmaxf(m0 - mx, kMsgNegInf);
    };
    c[i] = make_float4(diff(v.x, v.y), diff(v.z, v.w), diff(h.x, h.y), diff(h.z, h.w));
  }
}

__global__ void __launch_bounds__(kThreads)
k_lattice_expand(const float4* __restrict__ c, float4* __restrict__ m, int64_t cells) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < cells; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 x = c[i];
    m[2 * i] = make_float4(fminf(-x.x, 0.f), fminf(x.x, 0.f), fminf(-x.y, 0.f), fminf(x.y, 0.f));
    m[2 * i + 1] = make_float4(fminf(-x.z, 0.f), fminf(x.z, 0.f), fminf(-x.w, 0.f), fminf(x.w, 0.f));
  }
}

// ---------------------------------------------------------------------------
// Row strips (pgx_strip_*): what a rank sends each iteration, from the compressed messages.
//   down[2 j .. 2 j + 1] = the message of the last row's vertical factor into the variable
//                          below (owned by the next rank), both states;
//   up[2 j .. 2 j + 1]   = ev + (messages into the first row's variable j from THIS rank's
//                          factors, ascending message index): the ghost evidence of the
//                          previous rank, whose own vertical message completes the sum.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_strip_pack(int32_t R, int32_t N, const float* __restrict__ ev, const float4* __restrict__ c,
             float2* __restrict__ down, float2* __restrict__ up) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const float xb = c[int64_t(R - 1) * N + j].y;
  down[j] = make_float2(fminf(-xb, 0.f), fminf(xb, 0.f));
  const float4 own = c[j];
  const float left = c[j == 0 ? N - 1 : j - 1].w;
  const float2 e = reinterpret_cast<const float2*>(ev)[j];
  f32x2 s = pk2(e.x, e.y);
  if (j != 0) s = add2(s, bin_pair(left));
  s = add2(s, bin_pair(own.x));
  s = add2(s, bin_pair(own.z));
  if (j == 0) s = add2(s, bin_pair(left));
  float s0, s1;
  upk2(s, s0, s1);
  up[j] = make_float2(s0, s1);
}

// Beliefs of the owned variables of a strip (or of the whole torus when torus = 1) from the
// compressed messages: the variable sums of k_lattice_bin, written out.
__global__ void __launch_bounds__(kThreads)
k_lattice_bin_beliefs(LatticeBinArgs g, const float* __restrict__ ev, const float4* __restrict__ c,
                      float2* __restrict__ out) {
  const int64_t cells = int64_t(g.R) * g.N;
  const float2* ev2 = reinterpret_cast<const float2*>(ev);
  const float2* up2 = reinterpret_cast<const float2*>(g.up_add);
  const bool torus = g.torus != 0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < cells; i += int64_t(gridDim.x) * blockDim.x) {
    const int l = int(i / g.N), j = int(i - int64_t(l) * g.N);
    float2 e = ev2[i];
    if (up2 != nullptr && l == 0 && !torus) { e.x += up2[j].x; e.y += up2[j].y; }
    const bool has_up = torus || l > 0;
    const bool up_wrap = torus && l == 0;
    const bool left_wrap = j == 0;
    const float4 own = c[i];
    const float up = has_up ? c[int64_t(l == 0 ? g.R - 1 : l - 1) * g.N + j].y : 0.f;
    const float left = c[int64_t(l) * g.N + (j == 0 ? g.N - 1 : j - 1)].w;
    f32x2 s = pk2(e.x, e.y);
    if (has_up && !up_wrap) s = add2(s, bin_pair(up));
    if (!left_wrap) s = add2(s, bin_pair(left));
    s = add2(s, bin_pair(own.x));
    s = add2(s, bin_pair(own.z));
    if (left_wrap) s = add2(s, bin_pair(left));
    if (up_wrap) s = add2(s, bin_pair(up));
    float s0, s1;
    upk2(s, s0, s1);
    out[i] = make_float2(s0, s1);
  }
}

}  // namespace pgx
