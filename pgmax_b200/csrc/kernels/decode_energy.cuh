// pgx kernels - K6 / K7: beliefs + MAP decode, energy of a decoding.  Part of pgx_kernels.cuh (included in this order).
#pragma once

#include "logical.cuh"

namespace pgx {

// ---------------------------------------------------------------------------
// K6: fused beliefs + MAP decode + marginals + tie count
// (pgmax/infer/inferer.py:218-222,259-264; pgmax/infer/bp.py:283-288).
// One thread per (variable, sample).  Outputs are in the ABI's batch-major
// layout.  beliefs / marginals / map / ties may each be null.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_decode(BatchMap mp, int64_t num_vars, int64_t num_var_states,
         const int32_t* __restrict__ var_first_state, const int32_t* __restrict__ var_ptr,
         const int32_t* __restrict__ var_edge_msg, View ev, View m,
         float* __restrict__ beliefs, int32_t* __restrict__ map_out,
         float* __restrict__ marginals, int32_t* __restrict__ ties) {
  UnitLoop L = unit_loop(mp, num_vars);
  if (!L.b_ok) return;
  const int b = L.b;
  const LaneView evL = lane_view(ev, mp, b), mL = lane_view(m, mp, b);
  int ntie = 0;
  for (int64_t var = L.u; var < L.u_end; var += L.step) {
    const int64_t v0 = var_first_state[var], v1 = var_first_state[var + 1];
    const int64_t k0 = var_ptr[var], k1 = var_ptr[var + 1];
    float best = -INFINITY, second = -INFINITY;
    int arg = 0;
    for (int64_t v = v0; v < v1; ++v) {
      float acc = evL.at(v);
      for (int64_t k = k0; k < k1; ++k) acc += mL.at(var_edge_msg[k] + (v - v0));
      if (beliefs) beliefs[int64_t(b) * num_var_states + v] = acc;
      if (acc > best) { second = best; best = acc; arg = int(v - v0); }
      else if (acc > second) second = acc;
    }
    if (map_out) map_out[int64_t(b) * num_vars + var] = arg;
    if (v1 - v0 >= 2 && best == second) ++ntie;
    if (marginals) {
      // exp(x - logsumexp(x)), logsumexp = max + log sum exp(x - max)
      float sum = 0.f;
      for (int64_t v = v0; v < v1; ++v) {
        float acc = evL.at(v);
        for (int64_t k = k0; k < k1; ++k) acc += mL.at(var_edge_msg[k] + (v - v0));
        sum += expf(acc - best);
      }
      const float lse = best + logf(sum);
      for (int64_t v = v0; v < v1; ++v) {
        float acc = evL.at(v);
        for (int64_t k = k0; k < k1; ++k) acc += mL.at(var_edge_msg[k] + (v - v0));
        marginals[int64_t(b) * num_var_states + v] = expf(acc - lse);
      }
    }
  }
  if (ties != nullptr && ntie > 0) atomicAdd(ties + b, ntie);
}

// K6 from the variable sums a run left behind (PGX_RUN_FINAL_SUMS: S = evidence + final messages,
// tile-blocked, i.e. the beliefs): the same arg-max / tie / softmax rules, no message is read.
__global__ void __launch_bounds__(kThreads)
k_decode_sums(BatchMap mp, int64_t num_vars, int64_t num_var_states, const int32_t* __restrict__ var_first_state,
              const float* __restrict__ S, int32_t* __restrict__ map_out, float* __restrict__ marginals,
              int32_t* __restrict__ ties) {
  UnitLoop L = unit_loop(mp, num_vars);
  if (!L.b_ok) return;
  const int b = L.b;
  const float* SL = S + lane_off(mp, num_var_states, b);
  const int sh = mp.bx_log;
  int ntie = 0;
  for (int64_t var = L.u; var < L.u_end; var += L.step) {
    const int64_t v0 = var_first_state[var], v1 = var_first_state[var + 1];
    float best = -INFINITY, second = -INFINITY;
    int arg = 0;
    for (int64_t v = v0; v < v1; ++v) {
      const float acc = SL[v << sh];
      if (acc > best) { second = best; best = acc; arg = int(v - v0); }
      else if (acc > second) second = acc;
    }
    if (map_out) map_out[int64_t(b) * num_vars + var] = arg;
    if (v1 - v0 >= 2 && best == second) ++ntie;
    if (marginals) {
      float sum = 0.f;
      for (int64_t v = v0; v < v1; ++v) sum += expf(SL[v << sh] - best);
      const float lse = best + logf(sum);
      for (int64_t v = v0; v < v1; ++v) marginals[int64_t(b) * num_var_states + v] = expf(SL[v << sh] - lse);
    }
  }
  if (ties != nullptr && ntie > 0) atomicAdd(ties + b, ntie);
}

// ---------------------------------------------------------------------------
// K7: energy of a decoding (pgmax/infer/energy.py:53-148 and the per-type compute_energy:
// factor/enum.py:276-323, logical.py:295-358, pool.py:184-239).  One thread per (unit, sample),
// unit = variable or factor; every launch reduces its units to kEnergyChunks partial sums per
// sample in a fixed order (no float atomics), k_energy_sum adds the partials serially.
//   variable v:        -evidence[v, state(v)]
//   EnumFactor:        -log_potential of THE configuration the decoding selects (unclipped, as
//                      the reference), +inf if no valid configuration matches
//   OR / AND factor:   +inf unless [all parents in the pointed state] == [child in the pointed state]
//   Pool factor:       +inf unless #choices in state 1 == [indicator in state 1]
// map: [batch][num_vars] int32 (ABI order of the variables), or one shared row.
// ---------------------------------------------------------------------------
constexpr int kEnergyChunks = 32;

struct EnergyArgs {
  const int32_t* map;
  int64_t map_stride;       // num_vars, or 0 for a shared decoding
  const float* ev;
  int64_t ev_stride;        // V_s, or 0
  const float* lp;
  int64_t lp_stride;        // C, or 0
  const int32_t* var_first_state;
  const int32_t* vs_var;
  const int32_t* edge_vs;
};

__device__ __forceinline__ void energy_block_reduce(float acc, float* __restrict__ out) {
  __shared__ float red[kThreads];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (int(threadIdx.x) < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

// state the decoding assigns to the variable that owns var-state `vs`, and the state `vs` is
__device__ __forceinline__ void decoded_at(const EnergyArgs& e, const int32_t* mapb, int64_t vs, int& decoded, int& own) {
  const int var = e.vs_var[vs];
  decoded = mapb[var];
  own = int(vs - e.var_first_state[var]);
}

__global__ void __launch_bounds__(kThreads)
k_energy_vars(EnergyArgs e, int64_t num_vars, float* __restrict__ partial, int slots) {
  const int b = blockIdx.y;
  const int32_t* mapb = e.map + int64_t(b) * e.map_stride;
  const float* evb = e.ev + int64_t(b) * e.ev_stride;
  float acc = 0.f;
  for (int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; v < num_vars; v += int64_t(gridDim.x) * blockDim.x) {
    const int64_t v0 = e.var_first_state[v];
    if (e.var_first_state[v + 1] > v0) acc -= evb[v0 + mapb[v]];
  }
  energy_block_reduce(acc, partial + (int64_t(b) * slots) * kEnergyChunks + blockIdx.x);
}

__global__ void __launch_bounds__(kThreads)
k_energy_enum(EnergyArgs e, EnumBlockDev blk, float* __restrict__ partial, int slots, int slot) {
  const int b = blockIdx.y;
  const int32_t* mapb = e.map + int64_t(b) * e.map_stride;
  const float* lpb = e.lp + int64_t(b) * e.lp_stride;
  float acc = 0.f;
  for (int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; f < blk.num_factors; f += int64_t(gridDim.x) * blockDim.x) {
    const int64_t ebase = blk.edge_base(f);
    // configurations that contain the decoded state of the first variable, ascending k
    int dec, own;
    decoded_at(e, mapb, e.edge_vs[ebase], dec, own);
    float en = INFINITY;
    const int n0 = blk.edge_off[1] - blk.edge_off[0];
    if (dec >= 0 && dec < n0) {
      const int j_end = blk.t_ptr[dec + 1];
      for (int j = blk.t_ptr[dec]; j < j_end; ++j) {
        const int k = blk.t_k[j];
        bool match = true;
        for (int a = 1; a < blk.arity && match; ++a) {
          int dec_a;
          decoded_at(e, mapb, e.edge_vs[ebase + a], dec_a, own);
          match = blk.cfg_es[int64_t(k) * blk.arity + a] == blk.edge_off[a] + dec_a;
        }
        if (match) { en = -lpb[blk.pot_base(f) + k]; break; }
      }
    }
    acc += en;
  }
  energy_block_reduce(acc, partial + (int64_t(b) * slots + slot) * kEnergyChunks + blockIdx.x);
}

template <bool kPool>
__global__ void __launch_bounds__(kThreads)
k_energy_logical(EnergyArgs e, LogicalDev w, float* __restrict__ partial, int slots, int slot) {
  const int b = blockIdx.y;
  const int32_t* mapb = e.map + int64_t(b) * e.map_stride;
  float acc = 0.f;
  for (int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; f < w.num_factors; f += int64_t(gridDim.x) * blockDim.x) {
    const int64_t p0 = w.uniform > 0 ? f * w.uniform : w.parent_ptr[f];
    const int64_t p1 = w.uniform > 0 ? p0 + w.uniform : w.parent_ptr[f + 1];
    int dec, own;
    decoded_at(e, mapb, w.children_vs[f], dec, own);
    int lhs;
    if (kPool) {  // children_vs / parents_vs point at state 0: count the choices in state 1
      const int child = dec == own + 1;
      lhs = 0;
      for (int64_t i = p0; i < p1; ++i) {
        decoded_at(e, mapb, w.parents_vs[i], dec, own);
        lhs += dec == own + 1;
      }
      acc += lhs == child ? 0.f : INFINITY;
    } else {
      const int child = dec == own;
      lhs = 1;
      for (int64_t i = p0; i < p1; ++i) {
        decoded_at(e, mapb, w.parents_vs[i], dec, own);
        lhs &= dec == own;
      }
      acc += lhs == child ? 0.f : INFINITY;
    }
  }
  energy_block_reduce(acc, partial + (int64_t(b) * slots + slot) * kEnergyChunks + blockIdx.x);
}

__global__ void k_energy_sum(const float* __restrict__ partial, int n, int64_t batch, float* __restrict__ out) {
  const int64_t b = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  float acc = 0.f;
  for (int i = 0; i < n; ++i) acc += partial[b * n + i];
  out[b] = acc;
}

}  // namespace pgx
