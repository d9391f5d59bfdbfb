// pgx kernels - smooth dual LP-MAP (pgmax/infer/dual_lp.py:67-324).  Part of pgx_kernels.cuh (included in this order).
//
// One (sub)gradient step of the dual is
//   S      = evidence + scatter-add of the dual messages           k_var_sums        (dual_lp.py:113-115)
//   P, L_v = softmax / logsumexp (or one-hot arg-max / max) of S   k_sdlp_vars       (:146-160, :196-206)
//   U      = BP updates on vtof = -m, normalize=False, raw lp      k_enum_*<.., true>, k_logical_raw, k_pool_raw (:119-140)
//   per edge: softmax / logsumexp (arg-max / max) of U - m, the gradient P[vs] - softmax, the
//             accelerated step on (m, eta)                          k_sdlp_edges      (:162-181, :208-217, :293-309)
//   objval = sum_v L_v + sum_f max over the factor's edges of L_e   k_sdlp_objval_*   (:183-194, :219-230)
// All arrays are tile-blocked like the BP workspace (common.cuh); sums run in ascending index
// order (the order of a serial scatter-add), the objective is reduced in a fixed tree in fp64.
#pragma once

#include "logical.cuh"

namespace pgx {

// (relevant, other) outgoing messages to parent i: LogicalAcc::parent_out before the subtraction
// (pgmax/factor/logical.py:598-757 with normalize=False, :770-779).
template <bool kSumProduct>
__device__ __forceinline__ void logical_parent_pair(const LogicalAcc& A, int64_t i, float a_i, float b_i, float ca,
                                                    float cb, float T, bool single, float& PR, float& PO) {
  if (kSumProduct) {
    const float l_i = logaddexp_t(a_i, b_i, T);
    const float Lw = A.acc - l_i, Sw = A.Sb - b_i;
    PR = ca + Lw;
    const float o1 = cb + Sw, o2 = ca + Lw, o3 = ca + Sw;
    PO = logminusexp_t(logaddexp_t(o1, o2, T), o3, T, 1e-4f);
    if (T < kTempStabThre) {
      const float bound = (i == A.istar) ? (Sw + A.d2) : (Sw + A.d1);
      PO = fmaxf(PO, logaddexp_t(o1, ca + bound, T));
    }
  } else {
    const float mu = fmaxf(b_i, a_i);
    PR = (A.acc + ca) - mu;
    const float o1 = (cb + A.Sb) - b_i;
    const float o2 = PR + ((i == A.istar) ? fminf(0.f, A.d2) : fminf(0.f, A.d1));
    PO = fmaxf(o1, o2);
  }
  if (single) { PR = ca; PO = cb; }
}

// OR / AND update on vtof = -m, both states written (normalize=False).  Thread per (factor, sample).
template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_logical_raw(BatchMap mp, LogicalDev w, const float* __restrict__ m, float* __restrict__ upd, RunArgs a) {
  UnitLoop L = unit_loop(mp, w.num_factors);
  if (!L.b_ok) return;
  const int sh = mp.bx_log;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float* mo = m + moff;
  float* out = upd + moff;
  const int off = w.off;
  const float T = a.T;
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t p0 = w.parent_ptr[f], p1 = w.parent_ptr[f + 1];
    const int64_t c = w.children_msg[f];
    const float ca = -mo[(c + off) << sh], cb = -mo[c << sh];
    LogicalAcc A;
    for (int64_t i = p0; i < p1; ++i) {
      const int64_t pm = w.parents_msg[i];
      A.add<kSumProduct>(i, -mo[(pm + off) << sh], -mo[pm << sh], T);
    }
    const bool single = (p1 - p0) == 1;
    for (int64_t i = p0; i < p1; ++i) {
      const int64_t pm = w.parents_msg[i];
      float PR, PO;
      logical_parent_pair<kSumProduct>(A, i, -mo[(pm + off) << sh], -mo[pm << sh], ca, cb, T, single, PR, PO);
      out[(pm + off) << sh] = PR;
      out[pm << sh] = PO;
    }
    out[(c + off) << sh] = A.child_relevant<kSumProduct>(T);
    out[c << sh] = A.Sb;
  }
}

// Pool update on vtof = -m, both states written (pgmax/factor/pool.py:328-474, normalize=False :462-474).
template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_pool_raw(BatchMap mp, LogicalDev w, const float* __restrict__ m, float* __restrict__ upd, RunArgs a) {
  UnitLoop L = unit_loop(mp, w.num_factors);
  if (!L.b_ok) return;
  const int sh = mp.bx_log;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float* mo = m + moff;
  float* out = upd + moff;
  const float T = a.T;
  auto diff = [&](int64_t pm) { return (-mo[(pm + 1) << sh]) - (-mo[pm << sh]); };
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t p0 = w.parent_ptr[f], p1 = w.parent_ptr[f + 1];
    const int64_t c = w.children_msg[f];
    const float D = diff(c), ind_one = -mo[(c + 1) << sh];
    float d1 = -INFINITY, d2 = -INFINITY, zeros = 0.f;
    int64_t istar = p0;
    for (int64_t i = p0; i < p1; ++i) {
      const int64_t pm = w.parents_msg[i];
      const float dl = diff(pm);
      zeros += -mo[pm << sh];
      if (dl >= d1) { d2 = d1; d1 = dl; istar = i; }
      else if (dl > d2) d2 = dl;
    }
    const bool single = (p1 - p0) == 1;
    float out_ind = d1, G = 0.f, out_star = 0.f;
    if (kSumProduct) {
      float sum = 0.f, mx2 = -INFINITY;
      for (int64_t i = p0; i < p1; ++i) {
        const float dl = diff(w.parents_msg[i]);
        sum += expf((dl - d1) / T);
        mx2 = fmaxf(mx2, (i == istar) ? -D : dl);
      }
      out_ind = T * logf(sum) + d1;
      G = logaddexp_t(out_ind, -D, T);
      float sum2 = 0.f;
      for (int64_t i = p0; i < p1; ++i) {
        const float dl = diff(w.parents_msg[i]);
        sum2 += expf((((i == istar) ? -D : dl) - mx2) / T);
      }
      out_star = -(T * logf(sum2) + mx2);
    }
    for (int64_t i = p0; i < p1; ++i) {
      const int64_t pm = w.parents_msg[i];
      const float z = -mo[pm << sh];
      float x;
      if (kSumProduct) x = (i == istar) ? out_star : -logminusexp_t(G, diff(pm), T, 1e-30f);
      else x = fminf(D, -((i == istar) ? d2 : d1));
      float one = (zeros + ind_one) - z;
      if (single) { x = D; one = ind_one; }  // pool.py:430-450
      out[(pm + 1) << sh] = one;
      out[pm << sh] = one - x;
    }
    out[(c + 1) << sh] = zeros + out_ind;
    out[c << sh] = zeros;
  }
}

// Per variable: softmax + logsumexp at temperature T of the variable sums
// (update_utils.py:102-131), or for T == 0 the max and a one-hot of the arg-max (ties: the
// LARGEST index, update_utils.py:26-64).  Thread per (variable, sample).
template <bool kSmooth>
__global__ void __launch_bounds__(kThreads)
k_sdlp_vars(BatchMap mp, int64_t num_vars, int64_t Vs, const int32_t* __restrict__ var_first_state,
            const float* __restrict__ S, float* __restrict__ P, float* __restrict__ vval, float T) {
  UnitLoop L = unit_loop(mp, num_vars);
  if (!L.b_ok) return;
  const int sh = mp.bx_log;
  const float* SL = S + lane_off(mp, Vs, L.b);
  float* PL = P + lane_off(mp, Vs, L.b);
  float* vL = vval + lane_off(mp, num_vars, L.b);
  for (int64_t v = L.u; v < L.u_end; v += L.step) {
    const int64_t s0 = var_first_state[v], s1 = var_first_state[v + 1];
    float mx = -INFINITY;
    int64_t arg = -1;
    for (int64_t s = s0; s < s1; ++s) {
      const float x = SL[s << sh];
      if (x >= mx) { mx = x; arg = s; }
    }
    if (kSmooth) {
      float sum = 0.f;
      for (int64_t s = s0; s < s1; ++s) sum += T * expf((SL[s << sh] - mx) / T);
      for (int64_t s = s0; s < s1; ++s) PL[s << sh] = (T * expf((SL[s << sh] - mx) / T)) / sum;
      vL[v << sh] = mx + T * logf(sum / T);
    } else {
      for (int64_t s = s0; s < s1; ++s) PL[s << sh] = (s == arg) ? 1.f : 0.f;
      vL[v << sh] = mx;
    }
  }
}

// Per edge: softmax / logsumexp (arg-max / max) of the outgoing dual messages U - m, the
// (sub)gradient P[var-state] - softmax (dual_lp.py:162-181, 208-217), optionally stored, and
// optionally the accelerated step (dual_lp.py:293-309):
//   eta' = m - step * g;  m' = eta' + momentum * (eta' - eta).   Thread per (edge, sample).
template <bool kSmooth>
__global__ void __launch_bounds__(kThreads)
k_sdlp_edges(BatchMap mp, int64_t num_edges, int64_t Es, int64_t Vs, const int32_t* __restrict__ edge_msg_start,
             const int32_t* __restrict__ edge_vs, const float* __restrict__ U, const float* __restrict__ P,
             float* __restrict__ m, float* __restrict__ eta, float* __restrict__ grad, float* __restrict__ eval,
             float T, int do_step, float step, float momentum) {
  UnitLoop L = unit_loop(mp, num_edges);
  if (!L.b_ok) return;
  const int sh = mp.bx_log;
  const int64_t moff = lane_off(mp, Es, L.b);
  const float* UL = U + moff;
  float* mL = m + moff;
  float* etaL = eta ? eta + moff : nullptr;
  float* gL = grad ? grad + moff : nullptr;
  const float* PL = P + lane_off(mp, Vs, L.b);
  float* eL = eval + lane_off(mp, num_edges, L.b);
  for (int64_t e = L.u; e < L.u_end; e += L.step) {
    const int64_t s0 = edge_msg_start[e], s1 = edge_msg_start[e + 1], vs0 = edge_vs[e];
    float mx = -INFINITY;
    int64_t arg = -1;
    for (int64_t s = s0; s < s1; ++s) {
      const float x = UL[s << sh] - mL[s << sh];
      if (x >= mx) { mx = x; arg = s; }
    }
    float sum = 0.f;
    if (kSmooth) {
      for (int64_t s = s0; s < s1; ++s) sum += T * expf(((UL[s << sh] - mL[s << sh]) - mx) / T);
      eL[e << sh] = mx + T * logf(sum / T);
    } else {
      eL[e << sh] = mx;
    }
    for (int64_t s = s0; s < s1; ++s) {
      const float mv = mL[s << sh];
      float g;
      if (kSmooth) g = PL[(vs0 + s - s0) << sh] - (T * expf(((UL[s << sh] - mv) - mx) / T)) / sum;
      else g = PL[(vs0 + s - s0) << sh] + ((s == arg) ? -1.f : 0.f);
      if (gL) gL[s << sh] = g;
      if (do_step) {
        const float new_eta = mv - step * g;
        mL[s << sh] = new_eta + momentum * (new_eta - etaL[s << sh]);
        etaL[s << sh] = new_eta;
      }
    }
  }
}

// Objective: sum over the variables of vval + sum over the factors of the max over the
// factor's (contiguous) edges of eval.  Stage 1: every thread accumulates its grid-stride
// share in fp64 and stores it at partial[slot][sample]; stage 2: one CTA per sample adds the
// slots in a fixed tree.  Deterministic for a given grid.
__global__ void __launch_bounds__(kThreads)
k_sdlp_objval_partial(BatchMap mp, int64_t num_vars, int64_t num_edges, int64_t num_factors,
                      const int32_t* __restrict__ factor_edge_start, const float* __restrict__ vval,
                      const float* __restrict__ eval, double* __restrict__ partial, int64_t padded) {
  UnitLoop L = unit_loop(mp, num_vars + num_factors);
  if (!L.b_ok) return;
  const int sh = mp.bx_log;
  const float* vL = vval + lane_off(mp, num_vars, L.b);
  const float* eL = eval + lane_off(mp, num_edges, L.b);
  double acc = 0.0;
  const int64_t slot = L.u;
  for (int64_t u = L.u; u < L.u_end; u += L.step) {
    if (u < num_vars) {
      acc += double(vL[u << sh]);
    } else {
      const int64_t f = u - num_vars;
      float mx = -INFINITY;
      for (int64_t e = factor_edge_start[f]; e < factor_edge_start[f + 1]; ++e) mx = fmaxf(mx, eL[e << sh]);
      acc += double(mx);
    }
  }
  partial[slot * padded + L.b] = acc;
}

__global__ void __launch_bounds__(kThreads)
k_sdlp_objval_final(const double* __restrict__ partial, int64_t num_slots, int64_t padded, float* __restrict__ out,
                    int64_t out_stride, int64_t out_off) {
  __shared__ double red[kThreads];
  const int b = blockIdx.x;
  double acc = 0.0;
  for (int64_t s = threadIdx.x; s < num_slots; s += blockDim.x) acc += partial[s * padded + b];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (int(threadIdx.x) < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[int64_t(b) * out_stride + out_off] = float(red[0]);
}

}  // namespace pgx
