// pgx kernels - K2a-fused: dense-grid pairwise blocks (RBM) on binary-difference storage.  Part of pgx_kernels.cuh (included in this order).
#pragma once

#include "lattice.cuh"

namespace pgx {

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
// (kNarrow: 4 warps also with compressed input - batches whose last group of 8 sample tiles would
// be at most half full, e.g. the 128-sample shards of a batch of 1024 split over 8 GPUs: the idle
// warps of an 8-warp CTA would still hold half of the SM's shared memory)
__host__ __device__ constexpr int bip_warps(bool in_full, bool narrow = false) { return (in_full || narrow) ? 4 : 8; }

// dynamic shared memory of k_enum_pw2_bip
__host__ __device__ constexpr size_t bip_smem_bytes(int RI, int TJ, bool in_full, bool narrow = false) {
  return size_t(bip_warps(in_full, narrow)) * kBipStages * (in_full ? 4 : 2) * TJ * 32 * sizeof(float)  // rings
         + size_t(RI) * TJ * 4 * sizeof(float)                                                          // potentials
         + size_t(bip_warps(in_full, narrow)) * kBipStages * sizeof(uint64_t);                          // mbarriers
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
template <bool kSumProduct, int TJ, bool kDelta, bool kInFull, bool kNarrow = false>
__global__ void __launch_bounds__(bip_warps(kInFull, kNarrow) * 32)
k_enum_pw2_bip(int batch, int nbt_groups, BipDev g, const float* __restrict__ lp,
               const float* __restrict__ S, const float* __restrict__ m_old, int64_t old_rows,
               float* __restrict__ c_new, int64_t c_rows, float* __restrict__ part, int64_t part_rows,
               RunArgs a) {
  constexpr int kIn = kInFull ? 4 : 2;           // floats per factor and sample in the input rows
  constexpr int kStages = kBipStages;
  constexpr int kBipWarps = bip_warps(kInFull, kNarrow);
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

}  // namespace pgx
