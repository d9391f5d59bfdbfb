// pgx kernels - reverse mode through the BP iterations (vector-Jacobian product), sum-product,
// EnumFactors.  Part of pgx_kernels.cuh (included in this order).
#pragma once

#include "sdlp.cuh"

namespace pgx {

// ---------------------------------------------------------------------------
// The reference differentiates THROUGH run_bp with jax.grad (pgmax/infer/bp.py:98 wraps the update
// in jax.checkpoint; examples/grid_mrf.ipynb cells 15-16 take value_and_grad of a loss of the
// marginals with respect to the log potentials).  Behind a C ABI there is no tracer, so the
// library carries the reverse pass itself: pgx_bp_run_vjp re-runs the iterations with the generic
// two-pass kernels keeping every iterate, then walks them backwards.  One iteration, forward
// (App. A.1 / A.2):
//     S_v = ev_v + sum_e m_e          q_e = S_vs(e) - m_e        s_k = sum_{e in k} q_e + lp_k
//     f_e = T log sum_{k contains e} exp(s_k / T) - q_e          u_e = d m_e + (1 - d) f_e
//     m'_e = max(u_e - max_{e' in edge(e)} u_e', -1e32)
// and backward, given G' = dL/dm':
//     gu_e = [e not clipped] G'_e - [e is its edge's arg-max] sum_{e' in edge, not clipped} G'_e'
//     gf_e = (1 - d) gu_e       gs_k = sum_{e in k} gf_e exp((s_k - lse_e) / T)      dL/dlp_k += gs_k
//     gq_e = -gf_e + sum_{k contains e} gs_k       gS_v = sum_{e: vs(e) = v} gq_e   dL/dev_v += gS_v
//     G_e  = d gu_e - gq_e + gS_vs(e)                                                (= dL/dm_e)
// k_enum_vjp does the per-factor part for one block (thread per (factor, sample), factors of at
// most kSmallMaxNS edge-states); the sums over a variable's edges reuse k_var_sums.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_enum_vjp(BatchMap mp, EnumBlockDev blk, const int32_t* __restrict__ edge_vs, View lp, int lp_rows_per_sample,
           const float* __restrict__ S, const float* __restrict__ m_old, const float* __restrict__ g_new,
           float* __restrict__ g_q, float* __restrict__ g_m, float* __restrict__ g_lp, RunArgs a) {
  UnitLoop L = unit_loop(mp, blk.num_factors);
  if (!L.b_ok) return;
  const int sh = mp.bx_log;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float* mo = m_old + moff;
  const float* gn = g_new + moff;
  float* gq_out = g_q + moff;
  float* gm_out = g_m + moff;
  const float* SL = S + lane_off(mp, a.Vs, L.b);
  const LaneView lpL = lane_view(lp, mp, L.b);
  float* glpL = g_lp + lane_off(mp, lp_rows_per_sample, L.b);  // per-sample gradient of the potentials
  const float T = a.T, d = a.d, omd = a.one_minus_d;
  float q[kSmallMaxNS], lse[kSmallMaxNS], gf[kSmallMaxNS], gq[kSmallMaxNS];
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t mbase = blk.msg_base(f), ebase = blk.edge_base(f), pbase = blk.pot_base(f);
    for (int e = 0; e < blk.arity; ++e) {
      const int64_t vs = edge_vs[ebase + e];
      for (int s = blk.edge_off[e]; s < blk.edge_off[e + 1]; ++s)
        q[s] = SL[(vs + s - blk.edge_off[e]) << sh] - mo[(mbase + s) << sh];
    }
    // forward quantities: lse_e (the walk of k_enum_small)
    for (int s = 0; s < blk.ns; ++s) {
      const int j0 = blk.t_ptr[s], j1 = blk.t_ptr[s + 1];
      float M = -INFINITY;
      for (int j = j0; j < j1; ++j) {
        const int k = blk.t_k[j];
        float sk = 0.f;
        for (int e = 0; e < blk.arity; ++e) sk += q[blk.cfg_es[k * blk.arity + e]];
        M = fmaxf(M, sk + clip_lp(lpL.at(pbase + k)));
      }
      float sum = 0.f;
      for (int j = j0; j < j1; ++j) {
        const int k = blk.t_k[j];
        float sk = 0.f;
        for (int e = 0; e < blk.arity; ++e) sk += q[blk.cfg_es[k * blk.arity + e]];
        sum += expf((sk + clip_lp(lpL.at(pbase + k)) - M) / T);
      }
      lse[s] = j1 > j0 ? T * logf(sum) + M : -INFINITY;
    }
    // damping + per-edge normalisation, backwards
    for (int e = 0; e < blk.arity; ++e) {
      const int s0 = blk.edge_off[e], s1 = blk.edge_off[e + 1];
      float mx = -INFINITY;
      int arg = s0;
      for (int s = s0; s < s1; ++s) {
        const float u = damp(mo[(mbase + s) << sh], lse[s] - q[s], d, omd);
        gf[s] = u;  // parked
        if (u > mx) { mx = u; arg = s; }
      }
      float total = 0.f;
      for (int s = s0; s < s1; ++s) {
        const bool live = gf[s] - mx >= kMsgNegInf;  // not clipped
        const float g = live ? gn[(mbase + s) << sh] : 0.f;
        total += g;
        gq[s] = g;  // parked: gu before the arg-max correction
      }
      for (int s = s0; s < s1; ++s) {
        const float gu = gq[s] - (s == arg ? total : 0.f);
        gm_out[(mbase + s) << sh] = d * gu;
        gf[s] = omd * gu;
        gq[s] = -gf[s];
      }
    }
    // configurations: gs_k, the potentials' gradient, and its spread back to the q's
    for (int k = 0; k < blk.num_configs; ++k) {
      float sk = 0.f;
      for (int e = 0; e < blk.arity; ++e) sk += q[blk.cfg_es[k * blk.arity + e]];
      const float raw = lpL.at(pbase + k);
      sk += clip_lp(raw);
      float gs = 0.f;
      for (int e = 0; e < blk.arity; ++e) {
        const int s = blk.cfg_es[k * blk.arity + e];
        if (gf[s] != 0.f && lse[s] > -INFINITY) gs += gf[s] * expf((sk - lse[s]) / T);
      }
      for (int e = 0; e < blk.arity; ++e) gq[blk.cfg_es[k * blk.arity + e]] += gs;
      if (fabsf(raw) <= kLpMaxAbs) glpL[(pbase + k) << sh] += gs;  // (clipped potentials have zero gradient)
    }
    for (int s = 0; s < blk.ns; ++s) gq_out[(mbase + s) << sh] = gq[s];
  }
}

// G_e = g_m_e - g_q_e + gS_vs(e) for every edge-state (thread per (edge, sample)).
__global__ void __launch_bounds__(kThreads)
k_vjp_combine(BatchMap mp, int64_t num_edges, int64_t Es, int64_t Vs, const int32_t* __restrict__ edge_msg_start,
              const int32_t* __restrict__ edge_vs, const float* __restrict__ g_m, const float* __restrict__ g_q,
              const float* __restrict__ gS, float* __restrict__ g_out) {
  UnitLoop L = unit_loop(mp, num_edges);
  if (!L.b_ok) return;
  const int sh = mp.bx_log;
  const int64_t moff = lane_off(mp, Es, L.b);
  const float* gSL = gS + lane_off(mp, Vs, L.b);
  for (int64_t e = L.u; e < L.u_end; e += L.step) {
    const int64_t s0 = edge_msg_start[e], s1 = edge_msg_start[e + 1], vs = edge_vs[e];
    for (int64_t s = s0; s < s1; ++s)
      g_out[moff + (s << sh)] = g_m[moff + (s << sh)] - g_q[moff + (s << sh)] + gSL[(vs + s - s0) << sh];
  }
}

// Backward of the initial normalize_and_clip_msgs (bp.py:92-96) on the un-normalised input m_in:
// g_in_e = [not clipped] G_e - [arg-max] sum_{not clipped} G_e'.  In place on G.
__global__ void __launch_bounds__(kThreads)
k_vjp_normalize(BatchMap mp, int64_t num_edges, int64_t Es, const int32_t* __restrict__ edge_msg_start,
                const float* __restrict__ m_in, float* __restrict__ G) {
  UnitLoop L = unit_loop(mp, num_edges);
  if (!L.b_ok) return;
  const int sh = mp.bx_log;
  const int64_t moff = lane_off(mp, Es, L.b);
  for (int64_t e = L.u; e < L.u_end; e += L.step) {
    const int64_t s0 = edge_msg_start[e], s1 = edge_msg_start[e + 1];
    float mx = -INFINITY;
    int64_t arg = s0;
    for (int64_t s = s0; s < s1; ++s) {
      const float v = m_in[moff + (s << sh)];
      if (v > mx) { mx = v; arg = s; }
    }
    float total = 0.f;
    for (int64_t s = s0; s < s1; ++s) {
      const bool live = m_in[moff + (s << sh)] - mx >= kMsgNegInf;
      if (!live) G[moff + (s << sh)] = 0.f;
      total += G[moff + (s << sh)];
    }
    G[moff + (arg << sh)] -= total;
  }
}

// dst += src over a tile-blocked array of n floats.
__global__ void __launch_bounds__(kThreads)
k_vjp_accumulate(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    dst[i] += src[i];
}

// out[n] = sum over the samples of a tile-blocked [batch][N] array, in ascending sample order
// (gradient of an input that the batch shares).
__global__ void __launch_bounds__(kThreads)
k_vjp_sum_batch(BatchMap mp, int64_t N, const float* __restrict__ src, float* __restrict__ out) {
  for (int64_t n = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; n < N; n += int64_t(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < mp.batch; ++b) acc += src[lane_off(mp, N, b) + (n << mp.bx_log)];
    out[n] = acc;
  }
}

}  // namespace pgx
