// pgx kernels - K2a: pairwise binary EnumFactors (generic, pull, resident).  Part of pgx_kernels.cuh (included in this order).
#pragma once

#include "var_sums.cuh"

namespace pgx {

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
// K2a-bin: the same update with the messages in binary-difference storage (graphs whose every
// edge has two states; full sample tiles): a normalised message is (0, x) or (-x, 0), so ONE float
// x = m1 - m0 per edge is kept between iterations (NaN = both states at the -1e32 floor) - half
// the message bytes of k_var_sums + k_enum_pw2, which are bandwidth-bound.  Expanding is exact and
// the arithmetic on the expanded values is pw2_update's: bit-identical to k_enum_pw2 (tested).
// c_old / c_new: [tile][E][32], row = edge index.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void expand_bin(float x, float& m0, float& m1) {
  const bool fl = x != x;
  m0 = fl ? kMsgNegInf : fminf(-x, 0.f);
  m1 = fl ? kMsgNegInf : fminf(x, 0.f);
}

template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_enum_pw2_bin(BatchMap mp, int64_t num_factors, int64_t first_edge, int64_t first_pot, int64_t E,
               const int32_t* __restrict__ edge_vs, View lp, const float* __restrict__ S,
               const float* __restrict__ c_old, float* __restrict__ c_new, RunArgs a) {
  UnitLoop L = unit_loop(mp, num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int64_t coff = lane_off(mp, E, L.b);
  const float* co = c_old + coff;
  float* cn = c_new + coff;
  const float* SL = S + lane_off(mp, a.Vs, L.b);
  const LaneView lpL = lane_view(lp, mp, L.b);
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t e = first_edge + 2 * f;
    const int64_t vs0 = edge_vs[e], vs1 = edge_vs[e + 1];
    const float x0 = co[e << 5], x1 = co[(e + 1) << 5];
    const float Sv[4] = {SL[vs0 << 5], SL[(vs0 + 1) << 5], SL[vs1 << 5], SL[(vs1 + 1) << 5]};
    const int64_t pb = first_pot + 4 * f;
    const float lpv[4] = {clip_lp(lpL.at(pb)), clip_lp(lpL.at(pb + 1)), clip_lp(lpL.at(pb + 2)),
                          clip_lp(lpL.at(pb + 3))};
    float m[4], n[4];
    expand_bin(x0, m[0], m[1]);
    expand_bin(x1, m[2], m[3]);
    dmax = fmaxf(dmax, pw2_update<kSumProduct>(m, Sv, lpv, a, n));
    cn[e << 5] = n[1] - n[0];
    cn[(e + 1) << 5] = n[3] - n[2];
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

}  // namespace pgx
