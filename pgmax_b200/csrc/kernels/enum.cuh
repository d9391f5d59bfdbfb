// pgx kernels - K2b / K2c: general EnumFactors, small and RCN-size.  Part of pgx_kernels.cuh (included in this order).
#pragma once

#include <type_traits>

#include "dense_grid.cuh"

namespace pgx {

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
// kRaw (smooth dual LP-MAP, pgmax/infer/dual_lp.py:119-140): the variable->factor message is
// -m_old, the potentials are NOT clipped, and the update is written as is (normalize=False,
// no damping, no delta); S is unused.
// ---------------------------------------------------------------------------
// kSmem: the two per-thread arrays live in dynamic shared memory, one column per thread
// ([2 ns][blockDim] floats: conflict-free, sized by the block's ns) instead of two 64-float local
// arrays - 512 B of local memory per thread is 1 MB per SM at full occupancy, four times the L1,
// so the dynamically indexed walks went to L2 (the 17 x 3-state "heretic" factors: 1.28 ms per
// iteration before).  Same operations in the same order: bit-identical.
template <bool kSumProduct, bool kRaw = false, bool kSmem = false>
__global__ void __launch_bounds__(kThreads)
k_enum_small(BatchMap mp, EnumBlockDev blk, const int32_t* __restrict__ edge_vs, View lp,
             const float* __restrict__ S, const float* __restrict__ m_old,
             float* __restrict__ m_new, RunArgs a) {
  extern __shared__ float es_cols[];
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
  struct Local { float v[kSmallMaxNS]; };
  struct Column {
    float* p;
    __device__ __forceinline__ float& operator[](int s) const { return p[s * kThreads]; }
  };
  typename std::conditional<kSmem, Column, Local>::type q_store, nv_store;
  if constexpr (kSmem) {
    q_store.p = es_cols + threadIdx.x;
    nv_store.p = es_cols + size_t(blk.ns) * kThreads + threadIdx.x;
  }
  auto q = [&](int s) -> float& { if constexpr (kSmem) return q_store[s]; else return q_store.v[s]; };
  auto nv = [&](int s) -> float& { if constexpr (kSmem) return nv_store[s]; else return nv_store.v[s]; };
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t mbase = blk.msg_base(f);
    const int64_t ebase = blk.edge_base(f);
    const int64_t pbase = blk.pot_base(f);
    for (int e = 0; e < blk.arity; ++e) {
      const int64_t vs = edge_vs[ebase + e];
      for (int s = blk.edge_off[e]; s < blk.edge_off[e + 1]; ++s)
        q(s) = kRaw ? -mo[(mbase + s) << sh] : SL[(vs + s - blk.edge_off[e]) << sh] - mo[(mbase + s) << sh];
    }
    for (int s = 0; s < blk.ns; ++s) {
      const int j0 = blk.t_ptr[s], j1 = blk.t_ptr[s + 1];
      float M = -INFINITY;
      for (int j = j0; j < j1; ++j) {
        const int k = blk.t_k[j];
        float sk = 0.f;
        for (int e = 0; e < blk.arity; ++e) sk += q(blk.cfg_es[k * blk.arity + e]);
        sk += kRaw ? lpL.at(pbase + k) : clip_lp(lpL.at(pbase + k));
        M = fmaxf(M, sk);
      }
      float val = M;
      if (kSumProduct) {
        float sum = 0.f;
        for (int j = j0; j < j1; ++j) {
          const int k = blk.t_k[j];
          float sk = 0.f;
          for (int e = 0; e < blk.arity; ++e) sk += q(blk.cfg_es[k * blk.arity + e]);
          sk += kRaw ? lpL.at(pbase + k) : clip_lp(lpL.at(pbase + k));
          sum += expf((sk - M) / T);
        }
        val = T * logf(sum) + M;
      }
      nv(s) = kRaw ? val - q(s) : damp(mo[(mbase + s) << sh], val - q(s), a.d, a.one_minus_d);
    }
    if constexpr (kRaw) {
      for (int s = 0; s < blk.ns; ++s) mn[(mbase + s) << sh] = nv(s);
    } else {
      for (int e = 0; e < blk.arity; ++e) {
        const int s0 = blk.edge_off[e], s1 = blk.edge_off[e + 1];
        float mx = -INFINITY;
        for (int s = s0; s < s1; ++s) mx = fmaxf(mx, nv(s));
        for (int s = s0; s < s1; ++s) {
          const float out = fmaxf(nv(s) - mx, kMsgNegInf);
          const int64_t idx = (mbase + s) << sh;
          dmax = fmaxf(dmax, fabsf(out - mo[idx]));
          mn[idx] = out;
        }
      }
    }
  }
  if (!kRaw) publish_delta(a.deltas, int64_t(L.b) * a.delta_stride + a.delta_off, dmax);
}

// T log(z) + m with the hardware lg2 (absolute error <= 2^-22 of lg2, i.e. <= 1.7e-7 T on the message;
// z >= 1 whenever the edge-state has a configuration, log(0) = -inf otherwise): the accurate logf was a
// quarter of the instructions of the kernels below once their hot loops were lean.
__device__ __forceinline__ float lse_tail(float z, float m, float c_log) { return __fmaf_rn(c_log, lg2_approx(z), m); }

// ---------------------------------------------------------------------------
// K2b-cm: the same update CONFIGURATION-major.  k_enum_small walks, for every edge-state, the
// list of configurations that contain it: every configuration's score s_k is formed arity times
// per pass, behind three index loads each.  Here a thread walks its factor's configurations once
// per pass - s_k once, then arity read-max-write (pass 1) / read-add-write (pass 2) updates of
// per-thread shared-memory columns M[], Z[] - which visits the terms of every edge-state in the
// same ascending configuration order as the lists do:
//   * max-product: the same maxima (order-independent): bit-identical to k_enum_small;
//   * sum-product: the same order of additions; exp((s_k - M) / T) is ex2((s_k - M) * log2(e) / T)
//     (MUFU.EX2, relative error 2^-22, the approximation the pairwise kernels already use,
//     pairwise.cuh) instead of expf of a division: within the sum-product tolerance (tested), and
//     a fifth of the instructions.
// Dynamic smem: (2 or 3) * ns * blockDim floats.
// ---------------------------------------------------------------------------
// kArity: 2 / 3 = compile-time arity (the walks over a configuration's variables unroll; with a
// run-time arity those loops were two thirds of the instructions), 0 = any.
template <bool kSumProduct, int kArity>
__global__ void __launch_bounds__(kThreads)
k_enum_small_cm(BatchMap mp, EnumBlockDev blk, const int32_t* __restrict__ edge_vs, View lp,
                const float* __restrict__ S, const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  extern __shared__ float es_cols[];
  UnitLoop L = unit_loop(mp, blk.num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int sh = mp.bx_log, ns = blk.ns, arity = kArity > 0 ? kArity : blk.arity, C = blk.num_configs;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float* mo = m_old + moff;
  float* mn = m_new + moff;
  const float* SL = S + lane_off(mp, a.Vs, L.b);
  const LaneView lpL = lane_view(lp, mp, L.b);
  const float T = a.T, c = 1.4426950408889634f / T;
  float* q = es_cols + threadIdx.x;          // [ns] columns, stride kThreads
  float* M = q + size_t(ns) * kThreads;      // maxima, then the damped values
  float* Z = M + size_t(ns) * kThreads;      // sums (sum-product only)
  const int32_t* __restrict__ cfg = blk.cfg_es;
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t mbase = blk.msg_base(f), ebase = blk.edge_base(f), pbase = blk.pot_base(f);
    for (int e = 0; e < arity; ++e) {
      const int64_t vs = edge_vs[ebase + e];
      for (int s = blk.edge_off[e]; s < blk.edge_off[e + 1]; ++s) {
        q[s * kThreads] = SL[(vs + s - blk.edge_off[e]) << sh] - mo[(mbase + s) << sh];
        M[s * kThreads] = -INFINITY;
        if (kSumProduct) Z[s * kThreads] = 0.f;
      }
    }
    auto score = [&](int k) {
      float sk = 0.f;
#pragma unroll
      for (int e = 0; e < arity; ++e) sk += q[cfg[k * arity + e] * kThreads];
      return sk + clip_lp(lpL.at(pbase + k));
    };
    for (int k = 0; k < C; ++k) {
      const float sk = score(k);
#pragma unroll
      for (int e = 0; e < arity; ++e) {
        float* slot = M + cfg[k * arity + e] * kThreads;
        *slot = fmaxf(*slot, sk);
      }
    }
    if (kSumProduct) {
      for (int k = 0; k < C; ++k) {
        const float sk = score(k);
#pragma unroll
        for (int e = 0; e < arity; ++e) {
          const int es = cfg[k * arity + e] * kThreads;
          Z[es] += ex2_approx((sk - M[es]) * c);
        }
      }
    }
    for (int s = 0; s < ns; ++s) {
      const float val = kSumProduct ? lse_tail(Z[s * kThreads], M[s * kThreads], a.c_log) : M[s * kThreads];
      M[s * kThreads] = damp(mo[(mbase + s) << sh], val - q[s * kThreads], a.d, a.one_minus_d);
    }
    for (int e = 0; e < arity; ++e) {
      const int s0 = blk.edge_off[e], s1 = blk.edge_off[e + 1];
      float mx = -INFINITY;
      for (int s = s0; s < s1; ++s) mx = fmaxf(mx, M[s * kThreads]);
      for (int s = s0; s < s1; ++s) {
        const float out = fmaxf(M[s * kThreads] - mx, kMsgNegInf);
        const int64_t idx = (mbase + s) << sh;
        dmax = fmaxf(dmax, fabsf(out - mo[idx]));
        mn[idx] = out;
      }
    }
  }
  publish_delta(a.deltas, int64_t(L.b) * a.delta_stride + a.delta_off, dmax);
}

// ---------------------------------------------------------------------------
// K2b-dense: pairwise factors whose table holds ALL n0 x n1 configurations in row-major order
// (PairwiseFactorGroup over multi-state variables, pgmax/fgroup/enum.py:201-353: multi-label MRFs,
// the 17 x 3-state "heretic" model).  The configuration-major walk of k_enum_small_cm as two nested
// loops: no configuration-table loads, and the first variable's state i keeps its score term, its
// maximum and its sum in REGISTERS over the inner loop - only the second variable's columns are
// read-modify-written in shared memory.  Every edge-state still receives its terms in ascending
// configuration order (k = i n1 + j): the same operations in the same order as k_enum_small_cm,
// bit-identical to it (tested).
// ---------------------------------------------------------------------------
// kStageLp (potentials shared by the batch, full sample tiles: a warp = 32 samples of ONE factor): the
// factor's n0 n1 potentials are clipped and staged in shared memory by two coalesced loads per warp
// instead of one warp-uniform global load per configuration and pass behind the loop's dependences
// (long-scoreboard stalls were the top stall reason).
constexpr int kDenseMaxConfigs = 256;  // n0 + n1 <= 32
// Shared-memory columns per thread: q [ns] (the first variable's slots are overwritten by its damped
// values as they complete), and the second variable's maxima and sums [n1] each - the first
// variable's maximum and sum never leave registers: per row i the inner loop runs twice back to back
// (maximum, then sum), only the second variable needs the two global passes.
__host__ __device__ constexpr size_t pair_dense_cols(int ns, int n1, bool sum) { return size_t(ns) + (sum ? 2 : 1) * size_t(n1); }

template <bool kSumProduct, bool kStageLp>
__global__ void __launch_bounds__(kThreads)
k_enum_pair_dense(BatchMap mp, EnumBlockDev blk, const int32_t* __restrict__ edge_vs, View lp,
                  const float* __restrict__ S, const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  extern __shared__ float es_cols[];
  UnitLoop L = unit_loop(mp, blk.num_factors);
  const unsigned live = __ballot_sync(0xffffffffu, L.b_ok);  // (lanes beyond the batch leave: the warp syncs below name the rest)
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int sh = mp.bx_log, ns = blk.ns, n0 = blk.edge_off[1], n1 = ns - n0;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float* mo = m_old + moff;
  float* mn = m_new + moff;
  const float* SL = S + lane_off(mp, a.Vs, L.b);
  const LaneView lpL = lane_view(lp, mp, L.b);
  const float T = a.T, c = 1.4426950408889634f / T;
  float* q = es_cols + threadIdx.x;           // [ns] columns, stride kThreads; slots [0, n0) end up holding damped values
  float* qb = q + size_t(n0) * kThreads;      // the second variable's part of q
  float* Mb = q + size_t(ns) * kThreads;      // [n1] maxima, then the damped values
  float* Zb = Mb + size_t(n1) * kThreads;     // [n1] sums (sum-product only)
  // per-warp staging area behind the columns
  float* lps = es_cols + pair_dense_cols(ns, n1, kSumProduct) * kThreads + (threadIdx.x >> 5) * kDenseMaxConfigs;
  const int lane = threadIdx.x & 31;
  const int rank = __popc(live & ((1u << lane) - 1u)), nlive = __popc(live);
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t mbase = blk.msg_base(f), ebase = blk.edge_base(f), pbase = blk.pot_base(f);
    if (kStageLp) {
      __syncwarp(live);  // the previous factor's readers are done
      for (int t = rank; t < n0 * n1; t += nlive) lps[t] = clip_lp(lpL.q[pbase + t]);
      __syncwarp(live);
    }
    for (int e = 0; e < 2; ++e) {
      const int64_t vs = edge_vs[ebase + e];
      for (int s = blk.edge_off[e]; s < blk.edge_off[e + 1]; ++s)
        q[s * kThreads] = SL[(vs + s - blk.edge_off[e]) << sh] - mo[(mbase + s) << sh];
    }
    for (int j = 0; j < n1; ++j) {
      Mb[j * kThreads] = -INFINITY;
      if (kSumProduct) Zb[j * kThreads] = 0.f;
    }
    auto score = [&](float qi, int i, int j) {
      return (qi + qb[j * kThreads]) + (kStageLp ? lps[i * n1 + j] : clip_lp(lpL.at(pbase + int64_t(i) * n1 + j)));
    };
    // pass 1: the second variable's maxima (max-product: the first variable's too)
    for (int i = 0; i < n0; ++i) {
      const float qi = 0.f + q[i * kThreads];
      float mi = -INFINITY;
#pragma unroll 4
      for (int j = 0; j < n1; ++j) {
        const float sk = score(qi, i, j);
        mi = fmaxf(mi, sk);
        Mb[j * kThreads] = fmaxf(Mb[j * kThreads], sk);
      }
      if (!kSumProduct) q[i * kThreads] = damp(mo[(mbase + i) << sh], mi - q[i * kThreads], a.d, a.one_minus_d);
    }
    // pass 2 (sum-product): per row the first variable's maximum, then both sums
    if (kSumProduct) {
      for (int i = 0; i < n0; ++i) {
        const float qi = 0.f + q[i * kThreads];
        float mi = -INFINITY;
#pragma unroll 4
        for (int j = 0; j < n1; ++j) mi = fmaxf(mi, score(qi, i, j));
        float zi = 0.f;
#pragma unroll 4
        for (int j = 0; j < n1; ++j) {
          const float sk = score(qi, i, j);
          zi += ex2_approx((sk - mi) * c);
          Zb[j * kThreads] += ex2_approx((sk - Mb[j * kThreads]) * c);
        }
        const float val = lse_tail(zi, mi, a.c_log);
        q[i * kThreads] = damp(mo[(mbase + i) << sh], val - q[i * kThreads], a.d, a.one_minus_d);
      }
    }
    for (int j = 0; j < n1; ++j) {
      const float val = kSumProduct ? lse_tail(Zb[j * kThreads], Mb[j * kThreads], a.c_log) : Mb[j * kThreads];
      Mb[j * kThreads] = damp(mo[(mbase + n0 + j) << sh], val - qb[j * kThreads], a.d, a.one_minus_d);
    }
    for (int e = 0; e < 2; ++e) {
      const int s0 = blk.edge_off[e], s1 = blk.edge_off[e + 1];
      const float* D = e == 0 ? q : Mb - size_t(n0) * kThreads;  // damped value of edge-state s: D[s]
      float mx = -INFINITY;
      for (int s = s0; s < s1; ++s) mx = fmaxf(mx, D[s * kThreads]);
      for (int s = s0; s < s1; ++s) {
        const float out = fmaxf(D[s * kThreads] - mx, kMsgNegInf);
        const int64_t idx = (mbase + s) << sh;
        dmax = fmaxf(dmax, fabsf(out - mo[idx]));
        mn[idx] = out;
      }
    }
  }
  publish_delta(a.deltas, int64_t(L.b) * a.delta_stride + a.delta_off, dmax);
}

// ---------------------------------------------------------------------------
// K2b-few: complete pairwise tables in which ONE variable has only kFew = 2 ... 4 states (a
// many-state hidden variable against a binary / ternary pixel: the 17 x 3 "heretic" factors).  The
// few-state side's terms, maxima and sums are kFew-element REGISTER arrays (the inner loop over it
// is fully unrolled), the many-state side's are scalars of the outer loop: no shared-memory
// read-modify-write in the hot loops at all; shared memory holds one column per thread (the
// many-state side's q, overwritten by its damped values) and the warp's staged potentials.
// kFewSecond: the few-state variable is the factor's second one (else its first).  Either way both
// indices ascend, so every edge-state receives its terms in ascending configuration order: the same
// operations in the same order as k_enum_pair_dense / k_enum_small_cm - bit-identical (tested).
// Potentials shared by the batch, full sample tiles (a warp = 32 samples of one factor).
// ---------------------------------------------------------------------------
template <bool kSumProduct, int kFew, bool kFewSecond>
__global__ void __launch_bounds__(kThreads)
k_enum_pair_few(BatchMap mp, EnumBlockDev blk, const int32_t* __restrict__ edge_vs, View lp,
                const float* __restrict__ S, const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  extern __shared__ float es_cols[];
  UnitLoop L = unit_loop(mp, blk.num_factors);
  const unsigned live = __ballot_sync(0xffffffffu, L.b_ok);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int ns = blk.ns, nm = ns - kFew;  // states of the many-state variable
  // edge-state offsets of the two sides inside the factor's message span
  const int few0 = kFewSecond ? nm : 0, many0 = kFewSecond ? 0 : kFew;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float* mo = m_old + moff;
  float* mn = m_new + moff;
  const float* SL = S + lane_off(mp, a.Vs, L.b);
  const float T = a.T, c = 1.4426950408889634f / T;
  float* D = es_cols + threadIdx.x;  // [nm] column, stride kThreads: q of the many-state side, then its damped values
  float* lps = es_cols + size_t(nm) * kThreads + (threadIdx.x >> 5) * kDenseMaxConfigs;
  const int lane = threadIdx.x & 31;
  const int rank = __popc(live & ((1u << lane) - 1u)), nlive = __popc(live);
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t mbase = blk.msg_base(f), ebase = blk.edge_base(f), pbase = blk.pot_base(f);
    __syncwarp(live);  // the previous factor's readers are done
    for (int t = rank; t < nm * kFew; t += nlive) lps[t] = clip_lp(lp.p[pbase + t]);
    __syncwarp(live);
    const int64_t vs_few = edge_vs[ebase + (kFewSecond ? 1 : 0)], vs_many = edge_vs[ebase + (kFewSecond ? 0 : 1)];
    float qf[kFew], Mf[kFew], Zf[kFew];
#pragma unroll
    for (int j = 0; j < kFew; ++j) {
      qf[j] = SL[(vs_few + j) << 5] - mo[(mbase + few0 + j) << 5];
      Mf[j] = -INFINITY;
      Zf[j] = 0.f;
    }
    for (int i = 0; i < nm; ++i) D[i * kThreads] = SL[(vs_many + i) << 5] - mo[(mbase + many0 + i) << 5];
    // s_k = ((0 + q of the first variable's state) + q of the second's) + potential, k = first * n1 + second
    auto score = [&](float qi, int i, int j) {
      return kFewSecond ? ((0.f + qi) + qf[j]) + lps[i * kFew + j] : ((0.f + qf[j]) + qi) + lps[j * nm + i];
    };
    // pass 1: maxima of the few-state side (max-product: the many-state side's too, finished per row)
    for (int i = 0; i < nm; ++i) {
      const float qi = D[i * kThreads];
      float mi = -INFINITY;
#pragma unroll
      for (int j = 0; j < kFew; ++j) {
        const float sk = score(qi, i, j);
        mi = fmaxf(mi, sk);
        Mf[j] = fmaxf(Mf[j], sk);
      }
      if (!kSumProduct) D[i * kThreads] = damp(mo[(mbase + many0 + i) << 5], mi - qi, a.d, a.one_minus_d);
    }
    if (kSumProduct) {
      for (int i = 0; i < nm; ++i) {
        const float qi = D[i * kThreads];
        float sk[kFew], mi = -INFINITY, zi = 0.f;
#pragma unroll
        for (int j = 0; j < kFew; ++j) { sk[j] = score(qi, i, j); mi = fmaxf(mi, sk[j]); }
#pragma unroll
        for (int j = 0; j < kFew; ++j) {
          zi += ex2_approx((sk[j] - mi) * c);
          Zf[j] += ex2_approx((sk[j] - Mf[j]) * c);
        }
        D[i * kThreads] = damp(mo[(mbase + many0 + i) << 5], (lse_tail(zi, mi, a.c_log)) - qi, a.d, a.one_minus_d);
      }
    }
    {  // the few-state edge: damp, normalise, write
      float nf[kFew], mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < kFew; ++j) {
        const float val = kSumProduct ? lse_tail(Zf[j], Mf[j], a.c_log) : Mf[j];
        nf[j] = damp(mo[(mbase + few0 + j) << 5], val - qf[j], a.d, a.one_minus_d);
        mx = fmaxf(mx, nf[j]);
      }
#pragma unroll
      for (int j = 0; j < kFew; ++j) {
        const float out = fmaxf(nf[j] - mx, kMsgNegInf);
        const int64_t idx = (mbase + few0 + j) << 5;
        dmax = fmaxf(dmax, fabsf(out - mo[idx]));
        mn[idx] = out;
      }
    }
    {  // the many-state edge
      float mx = -INFINITY;
      for (int i = 0; i < nm; ++i) mx = fmaxf(mx, D[i * kThreads]);
      for (int i = 0; i < nm; ++i) {
        const float out = fmaxf(D[i * kThreads] - mx, kMsgNegInf);
        const int64_t idx = (mbase + many0 + i) << 5;
        dmax = fmaxf(dmax, fabsf(out - mo[idx]));
        mn[idx] = out;
      }
    }
  }
  publish_delta(a.deltas, int64_t(L.b) * a.delta_stride + a.delta_off, dmax);
}

// ---------------------------------------------------------------------------
// K2b-unary: EnumFactors over ONE variable whose configurations are all its states in order (the
// bias factors of an RBM, benchmark/rbm_lib.py:141-158).  Every edge-state is in exactly one
// configuration, so the general update collapses to  f_s = ((0 + q_s) + lp_s) - q_s  for max- AND
// sum-product (the logsumexp of one term is T log(exp(0)) + M = M exactly): the same operations as
// k_enum_small performs for such a factor, without its per-thread arrays and list walks -
// bit-identical (tested), a quarter of its time on the RBM's 1 284 unary factors.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_enum_unary(BatchMap mp, EnumBlockDev blk, const int32_t* __restrict__ edge_vs, View lp,
             const float* __restrict__ S, const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  UnitLoop L = unit_loop(mp, blk.num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int sh = mp.bx_log, ns = blk.ns;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float* mo = m_old + moff;
  float* mn = m_new + moff;
  const float* SL = S + lane_off(mp, a.Vs, L.b);
  const LaneView lpL = lane_view(lp, mp, L.b);
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t mbase = blk.msg_base(f), pbase = blk.pot_base(f);
    const int64_t vs = edge_vs[blk.edge_base(f)];
    auto damped = [&](int s) {
      const float m = mo[(mbase + s) << sh];
      const float q = SL[(vs + s) << sh] - m;
      const float M = (0.f + q) + clip_lp(lpL.at(pbase + s));
      return damp(m, M - q, a.d, a.one_minus_d);
    };
    float mx = -INFINITY;
    for (int s = 0; s < ns; ++s) mx = fmaxf(mx, damped(s));
    for (int s = 0; s < ns; ++s) {
      const float out = fmaxf(damped(s) - mx, kMsgNegInf);
      const int64_t idx = (mbase + s) << sh;
      dmax = fmaxf(dmax, fabsf(out - mo[idx]));
      mn[idx] = out;
    }
  }
  publish_delta(a.deltas, int64_t(L.b) * a.delta_stride + a.delta_off, dmax);
}

// K2b-unary on binary-difference storage (two-state variable; see k_enum_pw2_bin): row = edge index.
__global__ void __launch_bounds__(kThreads)
k_enum_unary_bin(BatchMap mp, EnumBlockDev blk, int64_t E, const int32_t* __restrict__ edge_vs, View lp,
                 const float* __restrict__ S, const float* __restrict__ c_old, float* __restrict__ c_new, RunArgs a) {
  UnitLoop L = unit_loop(mp, blk.num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int64_t coff = lane_off(mp, E, L.b);
  const float* co = c_old + coff;
  float* cn = c_new + coff;
  const float* SL = S + lane_off(mp, a.Vs, L.b);
  const LaneView lpL = lane_view(lp, mp, L.b);
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t row = blk.msg_base(f) >> 1, pbase = blk.pot_base(f);
    const int64_t vs = edge_vs[blk.edge_base(f)];
    float m[2];
    expand_bin(co[row << 5], m[0], m[1]);
    float n[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const float q = SL[(vs + s) << 5] - m[s];
      const float M = (0.f + q) + clip_lp(lpL.at(pbase + s));
      n[s] = damp(m[s], M - q, a.d, a.one_minus_d);
    }
    const float mx = fmaxf(n[0], n[1]);
    n[0] = fmaxf(n[0] - mx, kMsgNegInf);
    n[1] = fmaxf(n[1] - mx, kMsgNegInf);
    dmax = fmaxf(dmax, fmaxf(fabsf(n[0] - m[0]), fabsf(n[1] - m[1])));
    cn[row << 5] = n[1] - n[0];
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

template <bool kSumProduct, bool kRaw = false>  // kRaw: see k_enum_small
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
        q[s] = kRaw ? -mo[(mbase + s) << sh] : SL[(vs + s - s0) << sh] - mo[(mbase + s) << sh];
    }
    __syncthreads();
    for (int s = threadIdx.x; s < blk.ns; s += blockDim.x) {
      const int j0 = blk.t_ptr[s], j1 = blk.t_ptr[s + 1];
      float M = -INFINITY;
      for (int j = j0; j < j1; ++j) {
        const int k = blk.t_k[j];
        float sk = 0.f;
        for (int e = 0; e < blk.arity; ++e) sk += q[blk.cfg_es[k * blk.arity + e]];
        sk += kRaw ? lpL.at(pbase + k) : clip_lp(lpL.at(pbase + k));
        M = fmaxf(M, sk);
      }
      float val = M;
      if (kSumProduct) {
        float sum = 0.f;
        for (int j = j0; j < j1; ++j) {
          const int k = blk.t_k[j];
          float sk = 0.f;
          for (int e = 0; e < blk.arity; ++e) sk += q[blk.cfg_es[k * blk.arity + e]];
          sk += kRaw ? lpL.at(pbase + k) : clip_lp(lpL.at(pbase + k));
          sum += expf((sk - M) / T);
        }
        val = T * logf(sum) + M;
      }
      nv[s] = kRaw ? val - q[s] : damp(mo[(mbase + s) << sh], val - q[s], a.d, a.one_minus_d);
    }
    if (kRaw) {  // every thread wrote nv[s] for the states it now stores
      for (int s = threadIdx.x; s < blk.ns; s += blockDim.x) mn[(mbase + s) << sh] = nv[s];
      continue;
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
// Dynamic smem: (2 ns + 32 + nwarps * (n1 + 32) + 32) floats of the largest group.
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
  // lane l of lane-group g owns state lane_state[32 g + l] of the first variable (n0 = none): the
  // plan groups states with lists of similar length (and distinct shared-memory banks) so that the
  // lanes of a group finish together - fewer idle slots in the round schedule
  const int32_t* lane_state;
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
// CTAs per SM (register cap) and rounds per trip, A/B-measured on the RCN graph (B = 1, kernel
// ms; profiles/r02_z_rcn_ab.txt): (16, 3) 0.201 - (8, 4) 0.185 - (12, 4) 0.179 - (10, 5) 0.173 -
// (12, 5) 0.233 (spills) - (8, 6) 0.187.  Occupancy wins until the trip's registers spill.
#ifndef PGX_BIGMAX_CTAS
#define PGX_BIGMAX_CTAS 5
#endif
template <bool kFlatLp, bool kPerm>
__global__ void __launch_bounds__(kThreads, PGX_BIGMAX_CTAS)
k_enum_big_maxprod_all(BatchMap mp, const BigMaxGroup* __restrict__ groups, const int2* __restrict__ units,
                       int64_t num_units, unsigned int* __restrict__ counter,
                       const int32_t* __restrict__ edge_vs, View lp, const float* __restrict__ lpR,
                       const float* __restrict__ S,
                       const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  extern __shared__ float smem[];
  __shared__ unsigned int s_unit, s_grp;
  const int sh = mp.bx_log;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t total = num_units * mp.batch;
  // lane-groups of a unit are handed to the warps dynamically (20 groups over 8 warps as a static
  // 3 / 2 split leaves the CTA waiting for its slowest warp a fifth of the time)
  auto next_group = [&]() {
    unsigned int gi = 0;
    if (lane == 0) gi = atomicAdd(&s_grp, 1u);
    return int(__shfl_sync(0xffffffffu, gi, 0));
  };
  for (;;) {
    __syncthreads();  // previous unit done with shared memory (and with s_unit)
    if (threadIdx.x == 0) { s_unit = atomicAdd(counter, 1u); s_grp = 0; }
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
    float* M = q + ns + 32;                // [ns] maxima, then damped values ([ns, ns + 32) of q: zeros, what idle schedule entries read as their partner's q)
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
    if (threadIdx.x < 32) q[ns + threadIdx.x] = 0.f;
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
      // one trip = kPermTrip rounds; the next trip's loads (1.25 KiB of potentials per warp) are
      // in flight while this one is consumed: with 40 warps per SM that is ~50 KiB per SM in
      // flight, what HBM latency x bandwidth asks for
#ifndef PGX_PERM_TRIP
#define PGX_PERM_TRIP 10
#endif
      constexpr int kPermTrip = PGX_PERM_TRIP;
      for (int grp = kPerm ? next_group() : G.num_groups; kPerm && grp < G.num_groups; grp = next_group()) {
        const int a_own = G.lane_state[grp * 32 + lane];
        const float qa = a_own < n0 ? q[a_own] : 0.f;
        const int r_end = G.round_ptr[grp + 1];
        const uint32_t idle2 = uint32_t((n1 + lane) << 2) * 0x10001u;  // (table entries are byte offsets: state << 2)
        uint32_t en_n[kPermTrip / 2];
        float rl_n[kPermTrip];
        // full trips issue their loads without a predicate; only a group's last, partial trip
        // pays for the bounds checks
        auto request = [&](int r0) {
          if (r0 + kPermTrip <= r_end) {
#pragma unroll
            for (int p = 0; p < kPermTrip / 2; ++p) {
              en_n[p] = __ldg(rbl + (size_t(r0 + 2 * p) << 4));
              rl_n[2 * p] = __ldcs(lpr + (size_t(r0 + 2 * p) << 5));
              rl_n[2 * p + 1] = __ldcs(lpr + (size_t(r0 + 2 * p + 1) << 5));
            }
          } else {
#pragma unroll
            for (int p = 0; p < kPermTrip / 2; ++p) {
              const bool in = r0 + 2 * p < r_end;  // rounds come in pairs
              en_n[p] = in ? __ldg(rbl + (size_t(r0 + 2 * p) << 4)) : idle2;
              rl_n[2 * p] = in ? __ldcs(lpr + (size_t(r0 + 2 * p) << 5)) : -INFINITY;
              rl_n[2 * p + 1] = in ? __ldcs(lpr + (size_t(r0 + 2 * p + 1) << 5)) : -INFINITY;
            }
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
          // (the q read lands in the zero pad behind q: the sum stays -inf).
          // All q reads of the trip first (read-only: they overlap), then the read-max-write chain.
          uint32_t b_s[kPermTrip];
          float sk[kPermTrip];
#pragma unroll
          for (int p = 0; p < kPermTrip; ++p) {
            b_s[p] = (p & 1) ? (en[p >> 1] >> 16) : (en[p >> 1] & 0xffffu);  // byte offset of the partner state
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
      for (int grp = !kPerm ? next_group() : G.num_groups; !kPerm && grp < G.num_groups; grp = next_group()) {
        const int a_own = G.lane_state[grp * 32 + lane];
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
// K2c-sum: the SUM-product update (T > 0) of the same large sorted pairwise groups on the same
// lane-per-state round schedule and the same round-ordered potentials (k_bigmax_permute): every
// configuration is read ONCE per iteration.  The reference takes logsumexp over each
// edge-state's configurations with the maximum first (pgmax/factor/enum.py:451-475,
// jax.ops.segment_max then segment_sum of exp); visiting a configuration once means the maximum
// is not known when its term is added, so both sides keep a running (maximum m, sum s of
// exp((s_k - m) / T)) pair and rescale s when m grows ("online" logsumexp - the same value up to
// fp32 rounding of the rescaling factors; tested against the oracle at the tolerance of the other
// sum-product paths):
//   * a-side (the lane's own state): registers; one rescale per trip, one ex2 per configuration;
//   * b-side: per-warp private arrays Mw / Sw (conflict-free by the round schedule, as in the
//     max-product kernel): read m, s - update - write back, warp barrier between rounds; the warps'
//     pairs are merged once per factor, in warp order (lane-groups are handed to the warps
//     statically: run-to-run deterministic).
// exp((x - m) / T) = ex2((x - m) * log2(e) / T): MUFU.EX2, two per configuration.
// Dynamic smem: (2 ns + 32 + 2 nwarps (n1 + 32) + 32) floats of the largest group.
// Measured on the RCN graph at T = 1, B = 1 (profiles/r02_z_rcn_sum_ab.txt): 0.270 ms per launch
// (0.44 of the HBM peak on the algorithmic bytes; the instruction mix - two MUFU and five
// shared-memory accesses per configuration - bounds it, not the trip length: 0.270 - 0.290 ms
// over trips of 4 - 12 rounds and 3 - 4 CTAs per SM) against 2.07 ms for the thread-per-edge-state
// kernel k_enum_big, which walks every configuration list twice per side.
// ---------------------------------------------------------------------------
#ifndef PGX_BIGSUM_CTAS
#define PGX_BIGSUM_CTAS 4
#endif
#ifndef PGX_SUM_TRIP
#define PGX_SUM_TRIP 6
#endif
__global__ void __launch_bounds__(kThreads, PGX_BIGSUM_CTAS)
k_enum_big_sumprod_all(BatchMap mp, const BigMaxGroup* __restrict__ groups, const int2* __restrict__ units,
                       int64_t num_units, unsigned int* __restrict__ counter, const int32_t* __restrict__ edge_vs,
                       const float* __restrict__ lpR, const float* __restrict__ S, const float* __restrict__ m_old,
                       float* __restrict__ m_new, RunArgs a) {
  extern __shared__ float smem[];
  __shared__ unsigned int s_unit;
  const int sh = mp.bx_log;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t total = num_units * mp.batch;
  const float T = a.T, c = 1.4426950408889634f / T;
  constexpr float kLow = -3.0e38f;  // "no configuration yet": finite, so that (-inf) - kLow is -inf, not NaN
  constexpr int kTrip = PGX_SUM_TRIP;
  for (;;) {
    __syncthreads();
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
    const int wstride = n1 + 32;
    float* q = smem;                        // [ns]
    float* M = q + ns + 32;                 // [ns] logsumexp values, then damped values ([ns, ns + 32) of q: zeros for idle entries)
    float* Mw = M + ns;                     // [kBigWarps][n1 + 32] per-warp partner-side maxima
    float* Sw = Mw + kBigWarps * wstride;   // [kBigWarps][n1 + 32] ... and sums
    float* red = Sw + kBigWarps * wstride;
    const int64_t moff = lane_off(mp, a.Es, b);
    const float* mo = m_old + moff;
    float* mn = m_new + moff;
    const float* SL = S + lane_off(mp, a.Vs, b);
    const int64_t mbase = blk.msg_base(f), ebase = blk.edge_base(f);
    for (int e = 0; e < 2; ++e) {
      const int64_t vs = edge_vs[ebase + e];
      const int s0 = blk.edge_off[e], s1 = blk.edge_off[e + 1];
      for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
        q[s] = SL[(vs + s - s0) << sh] - mo[(mbase + s) << sh];
        M[s] = 0.f;
      }
    }
    for (int i = threadIdx.x; i < kBigWarps * wstride; i += blockDim.x) { Mw[i] = kLow; Sw[i] = 0.f; }
    if (threadIdx.x < 32) q[ns + threadIdx.x] = 0.f;
    __syncthreads();
    {
      const uint32_t qb_s = smem_u32(q) + 4u * n0;
      const uint32_t mw_s = smem_u32(Mw + warp * wstride);
      const uint32_t sw_off = 4u * kBigWarps * wstride;  // Sw slot = Mw slot + this
      auto lds_f = [](uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; };
      auto sts_f = [](uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); };
      const float* __restrict__ lpr = lpR + G.perm_base + f * (int64_t(G.num_rounds) * 32) + lane;
      const uint32_t* __restrict__ rbl = G.rounds_b + lane;
      // STATIC hand-out of the lane-groups (boustrophedon over the warps; the plan orders the groups
      // by list length): which warp adds which terms into its private (m, s) arrays fixes the fp32
      // rounding of the merged sums, so a dynamic hand-out (as in the max-product kernel, where
      // the order cannot matter) would make runs differ in the last bits
      for (int pass = 0; pass * kBigWarps < G.num_groups; ++pass) {
        const int grp = pass * kBigWarps + ((pass & 1) ? kBigWarps - 1 - warp : warp);
        if (grp >= G.num_groups) continue;
        const int a_own = G.lane_state[grp * 32 + lane];
        const float qa = a_own < n0 ? q[a_own] : 0.f;
        const int r_end = G.round_ptr[grp + 1];
        const uint32_t idle2 = uint32_t((n1 + lane) << 2) * 0x10001u;
        uint32_t en_n[kTrip / 2];
        float rl_n[kTrip];
        auto request = [&](int r0) {
          if (r0 + kTrip <= r_end) {
#pragma unroll
            for (int p = 0; p < kTrip / 2; ++p) {
              en_n[p] = __ldg(rbl + (size_t(r0 + 2 * p) << 4));
              rl_n[2 * p] = __ldcs(lpr + (size_t(r0 + 2 * p) << 5));
              rl_n[2 * p + 1] = __ldcs(lpr + (size_t(r0 + 2 * p + 1) << 5));
            }
          } else {
#pragma unroll
            for (int p = 0; p < kTrip / 2; ++p) {
              const bool in = r0 + 2 * p < r_end;  // rounds come in pairs
              en_n[p] = in ? __ldg(rbl + (size_t(r0 + 2 * p) << 4)) : idle2;
              rl_n[2 * p] = in ? __ldcs(lpr + (size_t(r0 + 2 * p) << 5)) : -INFINITY;
              rl_n[2 * p + 1] = in ? __ldcs(lpr + (size_t(r0 + 2 * p + 1) << 5)) : -INFINITY;
            }
          }
        };
        float m_a = kLow, s_a = 0.f;
        int r0 = G.round_ptr[grp];
        if (r0 < r_end) request(r0);
        for (; r0 < r_end; r0 += kTrip) {
          uint32_t en[kTrip / 2];
          float rl[kTrip];
#pragma unroll
          for (int p = 0; p < kTrip / 2; ++p) en[p] = en_n[p];
#pragma unroll
          for (int p = 0; p < kTrip; ++p) rl[p] = rl_n[p];
          if (r0 + kTrip < r_end) request(r0 + kTrip);
          uint32_t b_s[kTrip];
          float sk[kTrip];
#pragma unroll
          for (int p = 0; p < kTrip; ++p) {
            b_s[p] = (p & 1) ? (en[p >> 1] >> 16) : (en[p >> 1] & 0xffffu);
            sk[p] = lds_f(qb_s + b_s[p]);
          }
          float mt = -INFINITY;
#pragma unroll
          for (int p = 0; p < kTrip; ++p) {
            sk[p] = (qa + sk[p]) + rl[p];  // idle entries: -inf
            mt = fmaxf(mt, sk[p]);
          }
          // a-side: one rescale per trip
          const float hi = fmaxf(m_a, mt);
          s_a *= ex2_approx((m_a - hi) * c);
          m_a = hi;
#pragma unroll
          for (int p = 0; p < kTrip; ++p) s_a += ex2_approx((sk[p] - hi) * c);
          // b-side: read (m, s) - update - write, round by round
#pragma unroll
          for (int p = 0; p < kTrip; ++p) {
            const uint32_t slot = mw_s + b_s[p];
            const float m = lds_f(slot), sm = lds_f(slot + sw_off);
            const float x = sk[p];
            const float top = fmaxf(m, x);
            const float e = ex2_approx((fminf(m, x) - top) * c);
            sts_f(slot, top);
            sts_f(slot + sw_off, x > m ? sm * e + 1.f : sm + e);
            asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");
          }
        }
        if (a_own < n0) M[a_own] = T * logf(s_a) + m_a;  // no configuration: log(0) = -inf
      }
    }
    __syncthreads();
    for (int s = threadIdx.x; s < n1; s += blockDim.x) {
      float m = Mw[s];
      for (int w = 1; w < kBigWarps; ++w) m = fmaxf(m, Mw[w * wstride + s]);
      float sum = 0.f;
      for (int w = 0; w < kBigWarps; ++w) sum += Sw[w * wstride + s] * ex2_approx((Mw[w * wstride + s] - m) * c);
      M[n0 + s] = T * logf(sum) + m;
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

}  // namespace pgx
