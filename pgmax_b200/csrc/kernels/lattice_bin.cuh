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
//   * up_add: [2 N] messages into the strip's first row from the vertical factors of the strip
//     above, added to row 0's sums right after the evidence (or last, on the strip that holds the
//     torus' row 0: up_last) - the position the single graph's ascending message order gives them;
//   * torus = 0: a ghost row R below the last owner row; its evidence and the messages of the next
//     strip's own factors arrive unsummed (ghost_terms) and are added in the single graph's order,
//     so that N strips are BIT-IDENTICAL to one graph (ghost_ev / row R of the evidence array: the
//     plain partial-sum form, kept for callers that pre-sum).
// ---------------------------------------------------------------------------
struct LatticeBinArgs {
  int32_t R, N;        // owner rows, columns
  int32_t torus;       // 1: rows wrap; 0: ghost row R below the last owner row (receives only)
  int32_t seg_begin[2], seg_end[2];  // owner-row ranges this launch updates (empty: begin >= end)
  const float* up_add; // [2 N] or null
  const float* ghost_ev;  // [2 N] evidence of the ghost row (torus = 0), or null: row R of the evidence array
  // Row strips, exact single-graph summation order (pgx_strip_*): the ghost row's terms arrive
  // UNSUMMED, ghost_terms[8 j .. 8 j + 4] = (ev0, ev1, left H.b, own V.a, own H.a) of the next
  // strip's first row (compressed messages), so that the ghost variable's sum is formed exactly as
  // the single graph forms it (evidence, the message from above, left, own V, own H; the left
  // neighbour last in column 0); and on the strip that holds the torus' row 0 the message from
  // above - the largest message index of the variable - is added LAST (up_last).
  const float* ghost_terms;  // [8 N] or null
  int32_t up_last;
  int32_t ghost_up_last;  // the ghost row is the torus' row 0 (last strip): there too the message from above comes last
};

template <int TR_, int TC_, int STAGES_, int CONSUMERS_ = 512, int CTAS_ = 1>
struct LbCfg {
  static constexpr int TR = TR_, TC = TC_, kStages = STAGES_;
  static constexpr int kCtasPerSm = CTAS_;             // resident CTAs per SM the kernel is built for
  static constexpr int kConsumers = CONSUMERS_;        // consumer threads (+ one producer warp)
  static constexpr int kThreads = kConsumers + 32;
  static constexpr int kRpp = kConsumers / TC;         // tile rows the consumers cover side by side
  static constexpr int MR = TR + 2, MC = TC + 2;
  static constexpr int kMsgF4 = MR * MC;        // float4 per message stage (one per cell)
  static constexpr int kLpF4 = TR * TC * 2;     // float4 per potential stage (two factors per cell)
  static constexpr int kSumF2 = (TR + 1) * (TC + 1);
  static_assert(kConsumers % TC == 0 && TC % 32 == 0, "a warp must not straddle tile rows");
  static constexpr size_t smem_bytes() {
    return size_t(kStages) * (kMsgF4 + kLpF4) * sizeof(float4) + size_t(2) * kSumF2 * sizeof(float2) +
           2 * kStages * sizeof(uint64_t);
  }
};

__device__ __forceinline__ f32x2 bin_pair(float x) { return pk2(fminf(-x, 0.f), fminf(x, 0.f)); }

// Tile -> (first owner row, rows of the tile that are updated, first column).
struct LbTile {
  int l0, rows, j0;
};
template <class Cfg>
__device__ __forceinline__ LbTile lb_tile(const LatticeBinArgs& g, unsigned tile, unsigned tiles_x, int ty0) {
  const unsigned ty = tile / tiles_x, tx = tile - ty * tiles_x;
  LbTile t;
  const bool first = int(ty) < ty0;  // (selects, not array indexing: the struct stays in constant memory)
  const int tyl = first ? int(ty) : int(ty) - ty0;
  t.l0 = (first ? g.seg_begin[0] : g.seg_begin[1]) + tyl * Cfg::TR;
  t.rows = min(Cfg::TR, (first ? g.seg_end[0] : g.seg_end[1]) - t.l0);
  t.j0 = int(tx) * Cfg::TC;
  return t;
}

__device__ __forceinline__ void mbar_wait_u32(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}

// kClip: clip the potentials to +-1e6 on the fly (bp.py:85-87); the host passes false when it has
// checked that every potential of the run is already inside the range (the usual case).
template <bool kSumProduct, bool kDelta, class Cfg, bool kClip = true>
__global__ void __launch_bounds__(Cfg::kThreads, Cfg::kCtasPerSm)
k_lattice_bin(LatticeBinArgs g, const float* __restrict__ ev, const float* __restrict__ lp,
              const float4* __restrict__ c_old, float4* __restrict__ c_new, RunArgs a) {
  constexpr int TR = Cfg::TR, TC = Cfg::TC, MC = Cfg::MC, MR = Cfg::MR, kStages = Cfg::kStages;
  constexpr int kCons = Cfg::kConsumers, kRpp = Cfg::kRpp;
  extern __shared__ __align__(128) unsigned char lb_raw[];
  float4* msg_s = reinterpret_cast<float4*>(lb_raw);                         // [stage][MR][MC]
  float4* lp_s = msg_s + kStages * Cfg::kMsgF4;                              // [stage][TR][TC][2]
  float2* sum_s = reinterpret_cast<float2*>(lp_s + kStages * Cfg::kLpF4);    // [2][TR + 1][TC + 1]
  uint64_t* full = reinterpret_cast<uint64_t*>(sum_s + 2 * Cfg::kSumF2);     // [stage] tile landed
  uint64_t* done = full + kStages;                                           // [stage] tile updated
  const int R = g.R, N = g.N;
  const bool torus = g.torus != 0;
  const unsigned tiles_x = unsigned((N + TC - 1) / TC);
  const int ty0 = max(0, (g.seg_end[0] - g.seg_begin[0] + TR - 1) / TR);
  const int ty1 = max(0, (g.seg_end[1] - g.seg_begin[1] + TR - 1) / TR);
  const unsigned num_tiles = tiles_x * unsigned(ty0 + ty1);
  const int my_tiles = blockIdx.x < num_tiles ? int((num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&done[s], kCons);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const float4* lp4 = reinterpret_cast<const float4*>(lp);

  if (threadIdx.x >= kCons) {
    // ------------------------------- producer warp ---------------------------------------
    const int lane = threadIdx.x & 31;
    auto load_tile = [&](int i) {
      const LbTile t = lb_tile<Cfg>(g, blockIdx.x + unsigned(i) * gridDim.x, tiles_x, ty0);
      const int l0 = t.l0, j0 = t.j0;
      const int stage = i % kStages;
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
    for (int i = 0; i < my_tiles; ++i) {
      // the stage of tile i - 1 has been drained below: refill it with tile i + kStages - 1
      if (i + kStages - 1 < my_tiles) load_tile(i + kStages - 1);
      const int stage = i % kStages;
      mbar_wait(&done[stage], uint32_t(i / kStages) & 1u);
      const LbTile t = lb_tile<Cfg>(g, blockIdx.x + unsigned(i) * gridDim.x, tiles_x, ty0);
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
  // Thread -> (tile row r0 + k * kRpp, tile column cc): the column is fixed, a warp never
  // straddles rows, so every row condition below is warp-uniform.  Column TC of the sums (the
  // right halo) is an extra variable for the threads tid <= TR (row tid).
  const float2* ev2 = reinterpret_cast<const float2*>(ev);
  const float2* up2 = reinterpret_cast<const float2*>(g.up_add);
  const float2* ghost2 = reinterpret_cast<const float2*>(g.ghost_ev);
  const RunArgs2 c2 = make_args2(a);
  const int tid = threadIdx.x;
  const int cc = tid % TC, r0 = tid / TC;
  constexpr int kSumPass = (TR + 1 + kRpp - 1) / kRpp;
  constexpr int kFacPass = (TR + kRpp - 1) / kRpp;
  float dmax = 0.f;

  // Tiles of this CTA: tile index blockIdx.x + i * gridDim.x, coordinates advanced incrementally.
  struct Cursor {
    unsigned ty, tx;
  };
  const unsigned step_y = gridDim.x / tiles_x, step_x = gridDim.x - step_y * tiles_x;
  auto tile_at = [&](const Cursor& c) {
    LbTile t;
    const bool first = int(c.ty) < ty0;  // (selects, not array indexing: the struct stays in constant memory)
    const int tyl = first ? int(c.ty) : int(c.ty) - ty0;
    t.l0 = (first ? g.seg_begin[0] : g.seg_begin[1]) + tyl * TR;
    t.rows = min(TR, (first ? g.seg_end[0] : g.seg_end[1]) - t.l0);
    t.j0 = int(c.tx) * TC;
    return t;
  };
  auto advance = [&](Cursor& c) {
    c.ty += step_y;
    c.tx += step_x;
    if (c.tx >= tiles_x) { c.tx -= tiles_x; ++c.ty; }
  };
  // An interior tile touches no lattice border: every variable of its (TR + 1) x (TC + 1) sums is
  // an owner variable with all four neighbours in place (no wrap, no ghost, no halo addend).
  auto is_interior = [&](const LbTile& t) {
    return t.rows == TR && t.l0 >= 1 && t.l0 + TR <= R - 1 && t.j0 >= 1 && t.j0 + TC <= N - 1;
  };

  // evidence (+ halo addend) of variable (l0 + rr, j0 + c) of tile t; zeros if it is not needed
  auto ev_at = [&](const LbTile& t, int rr, int c) {
    int l = t.l0 + rr, j = t.j0 + c;
    float2 e = make_float2(0.f, 0.f);
    if (rr <= t.rows && l <= R && j <= N) {
      if (j == N) j = 0;
      if (torus && l == R) l = 0;
      if (g.ghost_terms != nullptr && l == R) e = __ldg(reinterpret_cast<const float2*>(g.ghost_terms + 8 * int64_t(j)));
      else e = (ghost2 != nullptr && l == R) ? __ldg(ghost2 + j) : __ldg(ev2 + (int64_t(l) * N + j));
      if (up2 != nullptr && l == 0 && !torus && !g.up_last) {  // halo: messages from the strip above
        const float2 u = __ldg(up2 + j);
        e.x += u.x;
        e.y += u.y;
      }
    }
    return e;
  };
  // S of variable (l0 + rr, j0 + c) from the staged messages: evidence first, then the incident
  // messages in ascending message index, wrap-around neighbours last (the order of k_lattice)
  auto var_sum = [&](const LbTile& t, const float4* sm, int rr, int c, float2 e) {
    int l = t.l0 + rr;
    const int j = t.j0 + c;
    if (torus && l == R) l = 0;
    const bool has_own = l < R;
    const bool has_up = torus || l > 0;
    const bool up_wrap = torus && l == 0;
    const bool left_wrap = j == 0 || j == N;
    // cell of column j0 + c sits at sm column c + 1, except the wrapped right halo
    const int col = (j == N) ? (N - t.j0) + 1 : c + 1;
    const float4 own = sm[(rr + 1) * MC + col];
    const float up = sm[rr * MC + col].y;                // V factor of the row above: its b edge
    const float left = sm[(rr + 1) * MC + col - 1].w;    // H factor of the left neighbour: its b edge
    const bool ghost = !has_own && g.ghost_terms != nullptr;
    f32x2 s = pk2(e.x, e.y);
    if (has_up && !up_wrap && !(ghost && g.ghost_up_last)) s = add2(s, bin_pair(up));
    if (has_own && !left_wrap) s = add2(s, bin_pair(left));
    if (has_own) { s = add2(s, bin_pair(own.x)); s = add2(s, bin_pair(own.z)); }
    if (has_own && left_wrap) s = add2(s, bin_pair(left));
    if (up_wrap) s = add2(s, bin_pair(up));
    if (ghost) {
      // ghost variable = first row of the next strip: its own strip's terms after the message from above
      const int jj = j == N ? 0 : j;
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(g.ghost_terms + 8 * int64_t(jj)));  // ev0, ev1, left, own V
      const float own_h = __ldg(g.ghost_terms + 8 * int64_t(jj) + 4);
      if (!left_wrap) s = add2(s, bin_pair(t4.z));
      s = add2(s, bin_pair(t4.w));
      s = add2(s, bin_pair(own_h));
      if (left_wrap) s = add2(s, bin_pair(t4.z));
      if (g.ghost_up_last) s = add2(s, bin_pair(up));
    }
    if (g.up_last && up2 != nullptr && l == 0 && !torus) {  // the torus' row 0: the message from row n - 1 comes last
      const float2 u = __ldg(up2 + (j == N ? 0 : j));
      s = add2(s, pk2(u.x, u.y));
    }
    float s0, s1;
    upk2(s, s0, s1);
    return make_float2(s0, s1);
  };
  // the same for an interior tile: up, left, own V, own H - the ascending message order
  auto var_sum_interior = [&](const float4* sm, int rr, int c, float2 e) {
    const float4 own = sm[(rr + 1) * MC + c + 1];
    const float up = sm[rr * MC + c + 1].y;
    const float left = sm[(rr + 1) * MC + c].w;
    f32x2 s = add2(pk2(e.x, e.y), bin_pair(up));
    s = add2(s, bin_pair(left));
    s = add2(s, bin_pair(own.x));
    s = add2(s, bin_pair(own.z));
    float s0, s1;
    upk2(s, s0, s1);
    return make_float2(s0, s1);
  };
  auto load_evidence = [&](const LbTile& t, float2 (&e)[kSumPass], float2& e_x) {
    if (is_interior(t)) {
      const float2* base = ev2 + (int64_t(t.l0) * N + t.j0);
#pragma unroll
      for (int k = 0; k < kSumPass; ++k)
        if (r0 + k * kRpp <= TR) e[k] = __ldg(base + ((r0 + k * kRpp) * N + cc));
      if (tid <= TR) e_x = __ldg(base + (tid * N + TC));
    } else {
#pragma unroll
      for (int k = 0; k < kSumPass; ++k) e[k] = ev_at(t, r0 + k * kRpp, cc);
      if (tid <= TR) e_x = ev_at(t, tid, TC);
    }
  };

  const uint32_t full_u32 = smem_u32(full), done_u32 = smem_u32(done);
  float2 e[kSumPass], e_x = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < kSumPass; ++k) e[k] = make_float2(0.f, 0.f);
  Cursor cur{blockIdx.x / tiles_x, blockIdx.x % tiles_x};
  LbTile t = tile_at(cur);
  if (my_tiles > 0) load_evidence(t, e, e_x);
  for (int i = 0; i < my_tiles; ++i) {
    const int stage = i % kStages;
    float4* sm = msg_s + stage * Cfg::kMsgF4;
    const float4* lq = lp_s + stage * Cfg::kLpF4;
    float2* Ss = sum_s + (i & 1) * Cfg::kSumF2;
    const bool interior = is_interior(t);
    mbar_wait_u32(full_u32 + stage * 8, uint32_t(i / kStages) & 1u);
    // ---- variable sums ------------------------------------------------------------------------
    if (interior) {
#pragma unroll
      for (int k = 0; k < kSumPass; ++k) {
        const int rr = r0 + k * kRpp;
        if (rr <= TR) Ss[rr * (TC + 1) + cc] = var_sum_interior(sm, rr, cc, e[k]);
      }
      if (tid <= TR) Ss[tid * (TC + 1) + TC] = var_sum_interior(sm, tid, TC, e_x);
    } else {
#pragma unroll
      for (int k = 0; k < kSumPass; ++k) {
        const int rr = r0 + k * kRpp;
        if (rr <= t.rows && t.l0 + rr <= R && t.j0 + cc <= N) Ss[rr * (TC + 1) + cc] = var_sum(t, sm, rr, cc, e[k]);
      }
      if (tid <= TR && tid <= t.rows && t.l0 + tid <= R && t.j0 + TC <= N) Ss[tid * (TC + 1) + TC] = var_sum(t, sm, tid, TC, e_x);
    }
    // the next tile's evidence: in flight during this tile's factor phase
    const LbTile t_now = t;
    if (i + 1 < my_tiles) {
      advance(cur);
      t = tile_at(cur);
      load_evidence(t, e, e_x);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kCons) : "memory");
    // ---- the two factors of every cell, in place ---------------------------------------------
#pragma unroll
    for (int k = 0; k < kFacPass; ++k) {
      const int rr = r0 + k * kRpp;
      if (rr < t_now.rows && t_now.j0 + cc < N) {
        const int cell = rr * TC + cc;
        float4 lv = lq[cell * 2], lh = lq[cell * 2 + 1];
        if (kClip) {
          lv = make_float4(clip_lp(lv.x), clip_lp(lv.y), clip_lp(lv.z), clip_lp(lv.w));
          lh = make_float4(clip_lp(lh.x), clip_lp(lh.y), clip_lp(lh.z), clip_lp(lh.w));
        }
        float4* slot = sm + (rr + 1) * MC + cc + 1;
        const float4 x = *slot;
        const float2 sa = Ss[rr * (TC + 1) + cc];
        const float2 sv = Ss[(rr + 1) * (TC + 1) + cc];
        const float2 sh = Ss[rr * (TC + 1) + cc + 1];
        const f32x2 Sa = pk2(sa.x, sa.y);
        float4 o;
        f32x2 na, nb;
        dmax = fmaxf(dmax, pw2_update_bin<kSumProduct, kDelta>(x.x, x.y, Sa, pk2(sv.x, sv.y), pk2(lv.x, lv.y),
                                                               pk2(lv.z, lv.w), c2, o.x, o.y, na, nb));
        dmax = fmaxf(dmax, pw2_update_bin<kSumProduct, kDelta>(x.z, x.w, Sa, pk2(sh.x, sh.y), pk2(lh.x, lh.y),
                                                               pk2(lh.z, lh.w), c2, o.z, o.w, na, nb));
        *slot = o;
      }
    }
    fence_proxy_async();
    mbar_arrive_u32(done_u32 + stage * 8);
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
//   up[8 j .. 8 j + 4]   = (ev0, ev1, left H.b, own V.a, own H.a) of the first row's variable j,
//                          UNSUMMED: the previous rank adds them to its ghost variable in the
//                          single graph's order (LatticeBinArgs::ghost_terms).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_strip_pack(int32_t R, int32_t N, const float* __restrict__ ev, const float4* __restrict__ c,
             float2* __restrict__ down, float4* __restrict__ up) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const float xb = c[int64_t(R - 1) * N + j].y;
  down[j] = make_float2(fminf(-xb, 0.f), fminf(xb, 0.f));
  const float4 own = c[j];
  const float left = c[j == 0 ? N - 1 : j - 1].w;
  const float2 e = reinterpret_cast<const float2*>(ev)[j];
  up[2 * j] = make_float4(e.x, e.y, left, own.x);
  up[2 * j + 1] = make_float4(own.z, 0.f, 0.f, 0.f);
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
    if (up2 != nullptr && l == 0 && !torus && !g.up_last) { e.x += up2[j].x; e.y += up2[j].y; }
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
    if (g.up_last && up2 != nullptr && l == 0 && !torus) s = add2(s, pk2(up2[j].x, up2[j].y));
    float s0, s1;
    upk2(s, s0, s1);
    out[i] = make_float2(s0, s1);
  }
}

}  // namespace pgx
