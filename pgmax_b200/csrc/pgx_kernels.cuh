// pgx_kernels.cuh — sm_100a kernels of the loopy-BP hot path.
//
// Internal data layout ("batch-inner"): every per-sample vector x[b][n] of the
// ABI (batch-major, as jax.vmap produces) is held as X[n][ld] with the batch
// index fastest (ld = batch rounded up to 8 floats, or 1 when batch == 1).
// A warp then covers 32 samples of ONE graph element, so
//   * every index load (incidence, wiring, config tables) is warp-uniform, and
//   * every message / evidence / var-sum access is a coalesced row segment,
// whatever the graph's structure.  For batch == 1 the same code degenerates to
// one graph element per lane over the reference's flat vectors.
//
// Arithmetic follows SURVEY.md App. A operation by operation (same order of
// additions, true division by the temperature, expf/logf/log1pf/expm1f, no
// FMA contraction: the library is compiled with -fmad=false) so that
// max-product results are bit-comparable with the CPU oracle.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace pgx {

constexpr float kMsgNegInf = -1e32f;   // pgmax/utils/__init__.py:26
constexpr float kLpMaxAbs = 1e6f;      // pgmax/utils/__init__.py:32
constexpr float kTempStabThre = 0.5f;  // pgmax/factor/logical.py:33
constexpr float kLn2 = 0.69314718055994530942f;
constexpr int kThreads = 256;
constexpr int kSmallMaxNS = 64;        // enum "small" kernel: edge-states per factor

// How threads map onto (graph element, sample) pairs.
struct BatchMap {
  int batch;   // B
  int ld;      // row pitch of batch-inner arrays
  int bx_log;  // log2 of samples covered by one warp (BX = min(32, pow2ceil(B)))
  int nbt;     // number of BX-wide sample tiles
};

// A strided view of a per-sample vector: element n of sample b is p[n*rs + b*bs].
// (rs, bs) = (ld, 1) for batch-inner workspace arrays, (1, 0) for an array shared
// by the whole batch that is read in place.
struct View {
  const float* p;
  int64_t rs;
  int64_t bs;
  __device__ __forceinline__ float at(int64_t n, int b) const { return p[n * rs + b * bs]; }
};

struct UnitLoop {
  int b;
  bool b_ok;
  int64_t u, u_end;
  int step;
};

// Splits `num_units` graph elements over the grid.  Warps are grouped into
// workers of `nbt` warps (one per sample tile); each worker walks a contiguous
// chunk of elements, so consecutive iterations touch consecutive index entries.
__device__ __forceinline__ UnitLoop unit_loop(const BatchMap& mp, int64_t num_units) {
  UnitLoop L;
  const int lane = threadIdx.x & 31;
  const int64_t gwarp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int bx = 1 << mp.bx_log;
  const int upw = 32 >> mp.bx_log;  // elements handled side by side in one warp
  const int64_t worker = gwarp / mp.nbt;
  const int64_t nworkers = nwarps / mp.nbt;
  const int bt = int(gwarp - worker * mp.nbt);
  L.b = bt * bx + (lane & (bx - 1));
  L.b_ok = L.b < mp.batch;
  int64_t chunk = (num_units + nworkers - 1) / nworkers;
  chunk = (chunk + upw - 1) / upw * upw;
  const int64_t u0 = worker * chunk;
  L.u_end = min(u0 + chunk, num_units);
  L.u = u0 + (lane >> mp.bx_log);
  L.step = upw;
  if (worker >= nworkers) L.u = L.u_end;  // warps beyond the last full worker idle
  return L;
}

__device__ __forceinline__ float clip_lp(float x) {
  return fminf(fmaxf(x, -kLpMaxAbs), kLpMaxAbs);  // pgmax/infer/bp.py:85-87
}

// max|m' - m| of one sample, accumulated with an integer atomicMax (valid for
// non-negative floats; NaNs are skipped).
__device__ __forceinline__ void publish_delta(float* deltas, int64_t idx, float d) {
  if (deltas != nullptr && d > 0.f) atomicMax(reinterpret_cast<int*>(deltas) + idx, __float_as_int(d));
}

// ---------------------------------------------------------------------------
// update_utils.py restated (pgmax/factor/update_utils.py:135-190)
// ---------------------------------------------------------------------------
__device__ __forceinline__ float logaddexp_t(float x, float y, float T) {
  const float mx = fmaxf(x, y), mn = fminf(x, y);
  return T * log1pf(expf((mn - mx) / T)) + mx;
}
__device__ __forceinline__ float log1mexp(float u) {
  return (u <= kLn2) ? logf(-expm1f(-u)) : log1pf(-expf(-u));
}
__device__ __forceinline__ float logminusexp_t(float x, float y, float T, float eps) {
  return (x >= y + eps) ? (T * log1mexp((x - y) / T) + x) : -INFINITY;
}

// ---------------------------------------------------------------------------
// Layout conversion: ABI batch-major [B][N]  <->  batch-inner [N][ld]
// ---------------------------------------------------------------------------
__global__ void k_to_batch_inner(const float* __restrict__ src, float* __restrict__ dst,
                                 int64_t N, int B, int ld) {
  __shared__ float tile[32][33];
  const int64_t n0 = int64_t(blockIdx.x) * 32;
  const int b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int b = b0 + r;
    const int64_t n = n0 + threadIdx.x;
    tile[r][threadIdx.x] = (b < B && n < N) ? src[int64_t(b) * N + n] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t n = n0 + r;
    const int b = b0 + threadIdx.x;
    if (n < N && b < ld) dst[n * ld + b] = tile[threadIdx.x][r];
  }
}

__global__ void k_from_batch_inner(const float* __restrict__ src, float* __restrict__ dst,
                                   int64_t N, int B, int ld) {
  __shared__ float tile[32][33];
  const int64_t n0 = int64_t(blockIdx.x) * 32;
  const int b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t n = n0 + r;
    const int b = b0 + threadIdx.x;
    tile[r][threadIdx.x] = (n < N && b < B) ? src[n * ld + b] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int b = b0 + r;
    const int64_t n = n0 + threadIdx.x;
    if (b < B && n < N) dst[int64_t(b) * N + n] = tile[threadIdx.x][r];
  }
}

// Broadcast a shared [N] vector into batch-inner [N][ld] (messages given once for the batch).
__global__ void k_broadcast_rows(const float* __restrict__ src, float* __restrict__ dst,
                                 int64_t N, int ld) {
  const int64_t total = N * ld;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x)
    dst[i] = src[i / ld];
}

// ---------------------------------------------------------------------------
// normalize_and_clip_msgs applied to the INPUT messages (pgmax/infer/bp.py:92-96,
// 249-259): per edge subtract the max over its states, clip below at -1e32.
// In place on the batch-inner buffer.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_normalize_edges(BatchMap mp, int64_t num_edges, const int32_t* __restrict__ edge_msg_start,
                  float* __restrict__ m) {
  UnitLoop L = unit_loop(mp, num_edges);
  if (!L.b_ok) return;
  for (int64_t e = L.u; e < L.u_end; e += L.step) {
    const int64_t s0 = edge_msg_start[e], s1 = edge_msg_start[e + 1];
    float mx = -INFINITY;
    for (int64_t s = s0; s < s1; ++s) mx = fmaxf(mx, m[s * mp.ld + L.b]);
    for (int64_t s = s0; s < s1; ++s) {
      float* p = m + s * mp.ld + L.b;
      *p = fmaxf(*p - mx, kMsgNegInf);
    }
  }
}

// ---------------------------------------------------------------------------
// K1: variable sums  S_v = ev_v + sum_{e incident to v} m_e, accumulated in
// ascending message index starting from the evidence (the order of a serial
// scatter-add, pgmax/infer/bp.py:217).  One thread per (var-state, sample)
// walking the variable's incident-edge list (CSR built by the plan).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_var_sums(BatchMap mp, int64_t num_var_states, const int32_t* __restrict__ vs_var,
           const int32_t* __restrict__ var_first_state, const int32_t* __restrict__ var_ptr,
           const int32_t* __restrict__ var_edge_msg, View ev, const float* __restrict__ m,
           float* __restrict__ S) {
  UnitLoop L = unit_loop(mp, num_var_states);
  if (!L.b_ok) return;
  for (int64_t v = L.u; v < L.u_end; v += L.step) {
    const int var = vs_var[v];
    const int64_t st = v - var_first_state[var];
    const int64_t k0 = var_ptr[var], k1 = var_ptr[var + 1];
    float acc = ev.at(v, L.b);
    int64_t k = k0;
    for (; k + 4 <= k1; k += 4) {  // loads are independent of the running sum
      const float a0 = m[(var_edge_msg[k] + st) * mp.ld + L.b];
      const float a1 = m[(var_edge_msg[k + 1] + st) * mp.ld + L.b];
      const float a2 = m[(var_edge_msg[k + 2] + st) * mp.ld + L.b];
      const float a3 = m[(var_edge_msg[k + 3] + st) * mp.ld + L.b];
      acc += a0; acc += a1; acc += a2; acc += a3;
    }
    for (; k < k1; ++k) acc += m[(var_edge_msg[k] + st) * mp.ld + L.b];
    S[v * mp.ld + L.b] = acc;
  }
}

// ---------------------------------------------------------------------------
// Shared epilogue: damping, per-edge max-normalisation, clip, delta
// (pgmax/infer/bp.py:127-136).  `one_minus_d` is computed on the host in fp32.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float damp(float m_old, float f, float d, float one_minus_d) {
  return d * m_old + one_minus_d * f;
}

// ---------------------------------------------------------------------------
// Pairwise binary EnumFactor with all 4 configurations valid (PairwiseFactorGroup
// over binary variables: Ising, RBM).  Everything in registers.
//   s_k = (q_a + q_b) + lp_k;  M_e = max over the 2 configs containing e;
//   T = 0: f_e = M_e - q_e;  T > 0: f_e = (T log sum exp((s_k - M_e)/T) + M_e) - q_e
// (pgmax/factor/enum.py:451-475, update_utils.py:68-98.)  With two terms the sum is
// exp(0) + exp((min - max)/T) = 1 + e exactly as the reference forms it, so one
// expf per edge-state suffices; (min - max)/T is evaluated as (min - max) * (1/T)
// (exact for T = 1; one extra rounding of the exponent otherwise, far inside the
// 1e-5 sum-product tolerance).
// ---------------------------------------------------------------------------
template <bool kSumProduct>
__device__ __forceinline__ float lse2(float a, float b, float T, float inv_T) {
  const float mx = fmaxf(a, b);
  if (!kSumProduct) return mx;
  const float mn = fminf(a, b);
  return T * logf(1.0f + expf((mn - mx) * inv_T)) + mx;
}

// In: old messages m[4] = (v0s0, v0s1, v1s0, v1s1), var sums S[4] of the same
// var-states, clipped potentials lp[4] in config order (0,0),(0,1),(1,0),(1,1).
// Out: n[4] damped + normalised + clipped; returns max|n - m|.
template <bool kSumProduct>
__device__ __forceinline__ float pw2_update(const float (&m)[4], const float (&Sv)[4],
                                            const float (&lp)[4], float d, float one_minus_d,
                                            float T, float inv_T, float (&n)[4]) {
  const float q0 = Sv[0] - m[0], q1 = Sv[1] - m[1], q2 = Sv[2] - m[2], q3 = Sv[3] - m[3];
  const float s00 = (q0 + q2) + lp[0], s01 = (q0 + q3) + lp[1];
  const float s10 = (q1 + q2) + lp[2], s11 = (q1 + q3) + lp[3];
  const float f0 = lse2<kSumProduct>(s00, s01, T, inv_T) - q0;
  const float f1 = lse2<kSumProduct>(s10, s11, T, inv_T) - q1;
  const float f2 = lse2<kSumProduct>(s00, s10, T, inv_T) - q2;
  const float f3 = lse2<kSumProduct>(s01, s11, T, inv_T) - q3;
  float n0 = damp(m[0], f0, d, one_minus_d), n1 = damp(m[1], f1, d, one_minus_d);
  float n2 = damp(m[2], f2, d, one_minus_d), n3 = damp(m[3], f3, d, one_minus_d);
  const float mxa = fmaxf(n0, n1), mxb = fmaxf(n2, n3);
  n[0] = fmaxf(n0 - mxa, kMsgNegInf); n[1] = fmaxf(n1 - mxa, kMsgNegInf);
  n[2] = fmaxf(n2 - mxb, kMsgNegInf); n[3] = fmaxf(n3 - mxb, kMsgNegInf);
  return fmaxf(fmaxf(fabsf(n[0] - m[0]), fabsf(n[1] - m[1])),
               fmaxf(fabsf(n[2] - m[2]), fabsf(n[3] - m[3])));
}

// K2a: one thread per (factor, sample).
template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_enum_pw2(BatchMap mp, int64_t num_factors, int64_t first_edge, int64_t first_msg,
           int64_t first_pot, const int32_t* __restrict__ edge_vs, View lp,
           const float* __restrict__ S, const float* __restrict__ m_old,
           float* __restrict__ m_new, float d, float one_minus_d, float T,
           float* __restrict__ deltas, int64_t delta_stride, int64_t delta_off) {
  UnitLoop L = unit_loop(mp, num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int b = L.b;
  const int64_t ld = mp.ld;
  const float inv_T = kSumProduct ? 1.0f / T : 0.f;
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t e = first_edge + 2 * f;
    const int64_t vs0 = edge_vs[e], vs1 = edge_vs[e + 1];
    const int64_t mb = (first_msg + 4 * f) * ld + b;
    const float m[4] = {m_old[mb], m_old[mb + ld], m_old[mb + 2 * ld], m_old[mb + 3 * ld]};
    const float Sv[4] = {S[vs0 * ld + b], S[(vs0 + 1) * ld + b], S[vs1 * ld + b],
                         S[(vs1 + 1) * ld + b]};
    const int64_t pb = first_pot + 4 * f;
    const float lpv[4] = {clip_lp(lp.at(pb, b)), clip_lp(lp.at(pb + 1, b)),
                          clip_lp(lp.at(pb + 2, b)), clip_lp(lp.at(pb + 3, b))};
    float n[4];
    dmax = fmaxf(dmax, pw2_update<kSumProduct>(m, Sv, lpv, d, one_minus_d, T, inv_T, n));
    m_new[mb] = n[0]; m_new[mb + ld] = n[1]; m_new[mb + 2 * ld] = n[2]; m_new[mb + 3 * ld] = n[3];
  }
  publish_delta(deltas, int64_t(b) * delta_stride + delta_off, dmax);
}

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
// A warp owns (sample tile of 32 lanes) x (strip of TJ columns) x (chunk of RI
// rows): column sums S_v and the column accumulators live in registers for the
// whole chunk, the row accumulator for one row.  The four warps of a CTA share
// strip and chunk (their potentials are staged once in shared memory) and cover
// four adjacent sample tiles, i.e. 512 contiguous bytes of every message row.
// ---------------------------------------------------------------------------
struct BipDev {
  int64_t first_msg, first_pot;
  int32_t I, J;          // rows, columns
  int32_t NS, NR, RI;    // column strips, row chunks, rows per chunk
  const int32_t* row_vs;   // [I] var-state of state 0 of row variable i
  const int32_t* col_vs;   // [J]
  const int32_t* row_part; // [I] partial-buffer row of (row var i, state 0, strip 0); state s, strip k at +2k+s
  const int32_t* col_part; // [J] same for column variables / row chunks
};

constexpr int kBipWarps = 4;
constexpr int kBipTJ = 16;

template <bool kSumProduct, int TJ>
__global__ void __launch_bounds__(kBipWarps * 32)
k_enum_pw2_bip(int batch, int ld, int nbt_groups, BipDev g, const float* __restrict__ lp,
               const float* __restrict__ S, const float* __restrict__ m_old,
               float* __restrict__ m_new, float* __restrict__ part, float d, float one_minus_d,
               float T, float* __restrict__ deltas, int64_t delta_stride, int64_t delta_off) {
  extern __shared__ float lp_s[];  // [RI][TJ][4] clipped potentials of this (strip, chunk)
  // blockIdx.x = (chunk * NS + strip) * nbt_groups + sample-tile group
  const int grp = blockIdx.x % nbt_groups;
  const int sc = blockIdx.x / nbt_groups;
  const int js = sc % g.NS, rc = sc / g.NS;
  const int j0 = js * TJ, i0 = rc * g.RI;
  const int i1 = min(i0 + g.RI, g.I);
  const int nj = min(TJ, g.J - j0);
  for (int t = threadIdx.x; t < (i1 - i0) * TJ * 4; t += blockDim.x) {
    const int r = t / (TJ * 4), c = t - r * (TJ * 4);
    lp_s[t] = (c < nj * 4) ? clip_lp(lp[g.first_pot + 4 * (int64_t(i0 + r) * g.J + j0) + c]) : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = (grp * kBipWarps + w) * 32 + lane;
  if (b >= batch) return;
  const float inv_T = kSumProduct ? 1.0f / T : 0.f;
  float Sc0[TJ], Sc1[TJ], ac0[TJ], ac1[TJ];
#pragma unroll
  for (int jj = 0; jj < TJ; ++jj) {
    const int64_t vs = g.col_vs[min(j0 + jj, g.J - 1)];
    Sc0[jj] = S[vs * ld + b];
    Sc1[jj] = S[(vs + 1) * ld + b];
    ac0[jj] = 0.f;
    ac1[jj] = 0.f;
  }
  float dmax = 0.f;
  for (int i = i0; i < i1; ++i) {
    const int64_t rvs = g.row_vs[i];
    const float Sr0 = S[rvs * ld + b], Sr1 = S[(rvs + 1) * ld + b];
    const int64_t mb = (g.first_msg + 4 * (int64_t(i) * g.J + j0)) * ld + b;
    const float* lrow = lp_s + (i - i0) * TJ * 4;
    float mo[TJ][4];
#pragma unroll
    for (int jj = 0; jj < TJ; ++jj) {
      if (jj < nj) {
#pragma unroll
        for (int k = 0; k < 4; ++k) mo[jj][k] = m_old[mb + int64_t(4 * jj + k) * ld];
      }
    }
    float ar0 = 0.f, ar1 = 0.f;
#pragma unroll
    for (int jj = 0; jj < TJ; ++jj) {
      if (jj < nj) {
        const float Sv[4] = {Sr0, Sr1, Sc0[jj], Sc1[jj]};
        const float lpv[4] = {lrow[4 * jj], lrow[4 * jj + 1], lrow[4 * jj + 2], lrow[4 * jj + 3]};
        float n[4];
        dmax = fmaxf(dmax, pw2_update<kSumProduct>(mo[jj], Sv, lpv, d, one_minus_d, T, inv_T, n));
#pragma unroll
        for (int k = 0; k < 4; ++k) m_new[mb + int64_t(4 * jj + k) * ld] = n[k];
        ar0 += n[0]; ar1 += n[1];
        ac0[jj] += n[2]; ac1[jj] += n[3];
      }
    }
    const int64_t pr = (int64_t(g.row_part[i]) + 2 * js) * ld + b;
    part[pr] = ar0;
    part[pr + ld] = ar1;
  }
#pragma unroll
  for (int jj = 0; jj < TJ; ++jj) {
    if (jj < nj) {
      const int64_t pc = (int64_t(g.col_part[j0 + jj]) + 2 * rc) * ld + b;
      part[pc] = ac0[jj];
      part[pc + ld] = ac1[jj];
    }
  }
  publish_delta(deltas, int64_t(b) * delta_stride + delta_off, dmax);
}

// ---------------------------------------------------------------------------
// K1-fused: S_v = ev_v + (messages of the edges that no fused block covers, in
// ascending message index) + (partial sums written by the fused blocks, in
// ascending partial row).  One thread per (var-state, sample).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_var_reduce(BatchMap mp, int64_t num_var_states, const int32_t* __restrict__ vs_var,
             const int32_t* __restrict__ var_first_state, const int32_t* __restrict__ rest_ptr,
             const int32_t* __restrict__ rest_edge_msg, const int32_t* __restrict__ part_first,
             const int32_t* __restrict__ part_count, View ev, const float* __restrict__ m,
             const float* __restrict__ part, float* __restrict__ S) {
  UnitLoop L = unit_loop(mp, num_var_states);
  if (!L.b_ok) return;
  const int64_t ld = mp.ld;
  for (int64_t v = L.u; v < L.u_end; v += L.step) {
    const int var = vs_var[v];
    const int64_t st = v - var_first_state[var];
    float acc = ev.at(v, L.b);
    for (int64_t k = rest_ptr[var]; k < rest_ptr[var + 1]; ++k)
      acc += m[(rest_edge_msg[k] + st) * ld + L.b];
    // partial rows of (var, state st): first + 2*k + st for binary variables
    const int64_t p0 = part_first[var];
    const int cnt = part_count[var];
    int k = 0;
    for (; k + 4 <= cnt; k += 4) {
      const float a0 = part[(p0 + 2 * k + st) * ld + L.b];
      const float a1 = part[(p0 + 2 * (k + 1) + st) * ld + L.b];
      const float a2 = part[(p0 + 2 * (k + 2) + st) * ld + L.b];
      const float a3 = part[(p0 + 2 * (k + 3) + st) * ld + L.b];
      acc += a0; acc += a1; acc += a2; acc += a3;
    }
    for (; k < cnt; ++k) acc += part[(p0 + 2 * k + st) * ld + L.b];
    S[v * ld + L.b] = acc;
  }
}

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
};

// ---------------------------------------------------------------------------
// K2b: EnumFactor update, small factors (ns <= 64): one thread per (factor,
// sample); q staged in a per-thread array, edge-state by edge-state walk of
// the transposed configuration lists.  Exact ascending-config order for both
// the max and the sum.
// ---------------------------------------------------------------------------
template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_enum_small(BatchMap mp, EnumBlockDev blk, const int32_t* __restrict__ edge_vs, View lp,
             const float* __restrict__ S, const float* __restrict__ m_old,
             float* __restrict__ m_new, float d, float one_minus_d, float T,
             float* __restrict__ deltas, int64_t delta_stride, int64_t delta_off) {
  UnitLoop L = unit_loop(mp, blk.num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int b = L.b;
  float q[kSmallMaxNS];
  float nv[kSmallMaxNS];
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t mbase = blk.first_msg + f * blk.ns;
    const int64_t ebase = blk.first_edge + f * blk.arity;
    const int64_t pbase = blk.first_pot + f * blk.num_configs;
    for (int a = 0; a < blk.arity; ++a) {
      const int64_t vs = edge_vs[ebase + a];
      for (int s = blk.edge_off[a]; s < blk.edge_off[a + 1]; ++s)
        q[s] = S[(vs + s - blk.edge_off[a]) * mp.ld + b] - m_old[(mbase + s) * mp.ld + b];
    }
    for (int s = 0; s < blk.ns; ++s) {
      const int j0 = blk.t_ptr[s], j1 = blk.t_ptr[s + 1];
      float M = -INFINITY;
      for (int j = j0; j < j1; ++j) {
        const int k = blk.t_k[j];
        float sk = 0.f;
        for (int a = 0; a < blk.arity; ++a) sk += q[blk.cfg_es[k * blk.arity + a]];
        sk += clip_lp(lp.at(pbase + k, b));
        M = fmaxf(M, sk);
      }
      float val = M;
      if (kSumProduct) {
        float sum = 0.f;
        for (int j = j0; j < j1; ++j) {
          const int k = blk.t_k[j];
          float sk = 0.f;
          for (int a = 0; a < blk.arity; ++a) sk += q[blk.cfg_es[k * blk.arity + a]];
          sk += clip_lp(lp.at(pbase + k, b));
          sum += expf((sk - M) / T);
        }
        val = T * logf(sum) + M;
      }
      nv[s] = damp(m_old[(mbase + s) * mp.ld + b], val - q[s], d, one_minus_d);
    }
    for (int a = 0; a < blk.arity; ++a) {
      const int s0 = blk.edge_off[a], s1 = blk.edge_off[a + 1];
      float mx = -INFINITY;
      for (int s = s0; s < s1; ++s) mx = fmaxf(mx, nv[s]);
      for (int s = s0; s < s1; ++s) {
        const float out = fmaxf(nv[s] - mx, kMsgNegInf);
        const int64_t idx = (mbase + s) * mp.ld + b;
        dmax = fmaxf(dmax, fabsf(out - m_old[idx]));
        m_new[idx] = out;
      }
    }
  }
  publish_delta(deltas, int64_t(b) * delta_stride + delta_off, dmax);
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

template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_enum_big(int batch, int ld, EnumBlockDev blk, const int32_t* __restrict__ edge_vs, View lp,
           const float* __restrict__ S, const float* __restrict__ m_old,
           float* __restrict__ m_new, float d, float one_minus_d, float T,
           float* __restrict__ deltas, int64_t delta_stride, int64_t delta_off) {
  extern __shared__ float smem[];
  float* q = smem;
  float* nv = smem + blk.ns;
  float* red = smem + 2 * blk.ns;
  const int64_t total = blk.num_factors * batch;
  for (int64_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
    const int64_t f = unit / batch;
    const int b = int(unit - f * batch);
    const int64_t mbase = blk.first_msg + f * blk.ns;
    const int64_t ebase = blk.first_edge + f * blk.arity;
    const int64_t pbase = blk.first_pot + f * blk.num_configs;
    __syncthreads();  // previous unit done with q / nv
    for (int a = 0; a < blk.arity; ++a) {
      const int64_t vs = edge_vs[ebase + a];
      const int s0 = blk.edge_off[a], s1 = blk.edge_off[a + 1];
      for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x)
        q[s] = S[(vs + s - s0) * ld + b] - m_old[(mbase + s) * ld + b];
    }
    __syncthreads();
    for (int s = threadIdx.x; s < blk.ns; s += blockDim.x) {
      const int j0 = blk.t_ptr[s], j1 = blk.t_ptr[s + 1];
      float M = -INFINITY;
      for (int j = j0; j < j1; ++j) {
        const int k = blk.t_k[j];
        float sk = 0.f;
        for (int a = 0; a < blk.arity; ++a) sk += q[blk.cfg_es[k * blk.arity + a]];
        sk += clip_lp(lp.at(pbase + k, b));
        M = fmaxf(M, sk);
      }
      float val = M;
      if (kSumProduct) {
        float sum = 0.f;
        for (int j = j0; j < j1; ++j) {
          const int k = blk.t_k[j];
          float sk = 0.f;
          for (int a = 0; a < blk.arity; ++a) sk += q[blk.cfg_es[k * blk.arity + a]];
          sk += clip_lp(lp.at(pbase + k, b));
          sum += expf((sk - M) / T);
        }
        val = T * logf(sum) + M;
      }
      nv[s] = damp(m_old[(mbase + s) * ld + b], val - q[s], d, one_minus_d);
    }
    float dmax = 0.f;
    for (int a = 0; a < blk.arity; ++a) {
      const int s0 = blk.edge_off[a], s1 = blk.edge_off[a + 1];
      __syncthreads();
      float mx = -INFINITY;
      for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x) mx = fmaxf(mx, nv[s]);
      mx = block_max(mx, red);
      for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
        const float out = fmaxf(nv[s] - mx, kMsgNegInf);
        const int64_t idx = (mbase + s) * ld + b;
        dmax = fmaxf(dmax, fabsf(out - m_old[idx]));
        m_new[idx] = out;
      }
    }
    if (deltas != nullptr) {
      dmax = block_max(dmax, red);
      if (threadIdx.x == 0) publish_delta(deltas, int64_t(b) * delta_stride + delta_off, dmax);
    }
  }
}

// ---------------------------------------------------------------------------
// Writes the two states of a binary edge whose factor->variable message is
// (0, x) or (x, 0): damping + normalisation + clip + delta.
//   lo = message index of the edge's state 0.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float write_binary_edge(const float* __restrict__ m_old,
                                                   float* __restrict__ m_new, int64_t lo, int ld,
                                                   int b, float f0, float f1, float d,
                                                   float one_minus_d) {
  const int64_t i0 = lo * ld + b, i1 = i0 + ld;
  const float a0 = m_old[i0], a1 = m_old[i1];
  float n0 = damp(a0, f0, d, one_minus_d), n1 = damp(a1, f1, d, one_minus_d);
  const float mx = fmaxf(n0, n1);
  n0 = fmaxf(n0 - mx, kMsgNegInf);
  n1 = fmaxf(n1 - mx, kMsgNegInf);
  m_new[i0] = n0;
  m_new[i1] = n1;
  return fmaxf(fabsf(n0 - a0), fabsf(n1 - a1));
}

// Device-side logical / pool wiring.  parent_ptr[f]..parent_ptr[f+1] indexes the
// parents of factor f; *_msg are global message indices of the wiring's "p_i" /
// "c" state, *_vs the var-state index of that same state.
struct LogicalDev {
  int64_t num_factors;
  const int32_t* parent_ptr;
  const int32_t* parents_msg;
  const int32_t* parents_vs;
  const int32_t* children_msg;
  const int32_t* children_vs;
  int32_t off;  // +1 OR / Pool, -1 AND
};

// ---------------------------------------------------------------------------
// K4: OR / AND update, closed form from per-factor sums and the two largest
// parent differences (pgmax/factor/logical.py:561-779).
// One thread per (factor, sample); two passes over the parents.
// ---------------------------------------------------------------------------
template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_logical(BatchMap mp, LogicalDev w, const float* __restrict__ S,
          const float* __restrict__ m_old, float* __restrict__ m_new, float d,
          float one_minus_d, float T, float* __restrict__ deltas, int64_t delta_stride,
          int64_t delta_off) {
  UnitLoop L = unit_loop(mp, w.num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int b = L.b;
  const int ld = mp.ld;
  const int off = w.off;
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t p0 = w.parent_ptr[f], p1 = w.parent_ptr[f + 1];
    const int64_t c = w.children_msg[f], cvs = w.children_vs[f];
    const float ca = S[(cvs + off) * ld + b] - m_old[(c + off) * ld + b];  // "relevant" state
    const float cb = S[cvs * ld + b] - m_old[c * ld + b];                  // "other" state
    // Pass 1: sums in ascending parent order, first / second max of the differences
    // (first arg-max = LARGEST tied index, update_utils.py:51-63).
    float Sb = 0.f, acc = 0.f, d1 = -INFINITY, d2 = -INFINITY;
    int64_t istar = p0;
    for (int64_t i = p0; i < p1; ++i) {
      const int64_t pm = w.parents_msg[i], pv = w.parents_vs[i];
      const float a_i = S[(pv + off) * ld + b] - m_old[(pm + off) * ld + b];
      const float b_i = S[pv * ld + b] - m_old[pm * ld + b];
      const float dl = a_i - b_i;
      Sb += b_i;
      acc += kSumProduct ? logaddexp_t(a_i, b_i, T) : fmaxf(b_i, a_i);
      if (dl >= d1) { d2 = d1; d1 = dl; istar = i; }
      else if (dl > d2) d2 = dl;
    }
    const bool single = (p1 - p0) == 1;
    float CR;
    if (kSumProduct) {
      CR = logminusexp_t(acc, Sb, T, 1e-4f);
      if (T < kTempStabThre) CR = fmaxf(CR, logaddexp_t(Sb + d1, Sb + d2, T));
    } else {
      CR = acc + fminf(0.f, d1);
    }
    // Pass 2: outgoing messages to the parents.
    for (int64_t i = p0; i < p1; ++i) {
      const int64_t pm = w.parents_msg[i], pv = w.parents_vs[i];
      const float a_i = S[(pv + off) * ld + b] - m_old[(pm + off) * ld + b];
      const float b_i = S[pv * ld + b] - m_old[pm * ld + b];
      float PR, PO;
      if (kSumProduct) {
        const float l_i = logaddexp_t(a_i, b_i, T);
        const float Lw = acc - l_i, Sw = Sb - b_i;
        PR = ca + Lw;
        const float o1 = cb + Sw, o2 = ca + Lw, o3 = ca + Sw;
        PO = logminusexp_t(logaddexp_t(o1, o2, T), o3, T, 1e-4f);
        if (T < kTempStabThre) {
          const float bound = (i == istar) ? (Sw + d2) : (Sw + d1);
          PO = fmaxf(PO, logaddexp_t(o1, ca + bound, T));
        }
      } else {
        const float mu = fmaxf(b_i, a_i);
        PR = (acc + ca) - mu;
        const float o1 = (cb + Sb) - b_i;
        const float o2 = PR + ((i == istar) ? fminf(0.f, d2) : fminf(0.f, d1));
        PO = fmaxf(o1, o2);
      }
      if (single) { PR = ca; PO = cb; }  // logical.py:739-757
      const float x = PR - PO;          // message of the "p_i + off" state; the other state gets 0
      const int64_t lo = (off > 0) ? pm : pm - 1;
      dmax = fmaxf(dmax, write_binary_edge(m_old, m_new, lo, ld, b, off > 0 ? 0.f : x,
                                           off > 0 ? x : 0.f, d, one_minus_d));
    }
    const float xc = CR - Sb;
    const int64_t lo = (off > 0) ? c : c - 1;
    dmax = fmaxf(dmax, write_binary_edge(m_old, m_new, lo, ld, b, off > 0 ? 0.f : xc,
                                         off > 0 ? xc : 0.f, d, one_minus_d));
  }
  publish_delta(deltas, int64_t(b) * delta_stride + delta_off, dmax);
}

// ---------------------------------------------------------------------------
// K5: Pool update (pgmax/factor/pool.py:328-474).
// ---------------------------------------------------------------------------
template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_pool(BatchMap mp, LogicalDev w, const float* __restrict__ S, const float* __restrict__ m_old,
       float* __restrict__ m_new, float d, float one_minus_d, float T,
       float* __restrict__ deltas, int64_t delta_stride, int64_t delta_off) {
  UnitLoop L = unit_loop(mp, w.num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int b = L.b;
  const int ld = mp.ld;
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t p0 = w.parent_ptr[f], p1 = w.parent_ptr[f + 1];
    const int64_t c = w.children_msg[f], cvs = w.children_vs[f];
    const float D = (S[(cvs + 1) * ld + b] - m_old[(c + 1) * ld + b]) -
                    (S[cvs * ld + b] - m_old[c * ld + b]);
    float d1 = -INFINITY, d2 = -INFINITY;
    int64_t istar = p0;
    for (int64_t i = p0; i < p1; ++i) {
      const int64_t pm = w.parents_msg[i], pv = w.parents_vs[i];
      const float dl = (S[(pv + 1) * ld + b] - m_old[(pm + 1) * ld + b]) -
                       (S[pv * ld + b] - m_old[pm * ld + b]);
      if (dl >= d1) { d2 = d1; d1 = dl; istar = i; }
      else if (dl > d2) d2 = dl;
    }
    const bool single = (p1 - p0) == 1;
    float out_ind = d1, G = 0.f, out_star = 0.f;
    if (kSumProduct) {
      // logsumexp over the choices with the precomputed max, and over the set where
      // the arg-max choice is replaced by -D (own max), both in ascending order.
      float sum = 0.f, mx2 = -INFINITY;
      for (int64_t i = p0; i < p1; ++i) {
        const int64_t pm = w.parents_msg[i], pv = w.parents_vs[i];
        const float dl = (S[(pv + 1) * ld + b] - m_old[(pm + 1) * ld + b]) -
                         (S[pv * ld + b] - m_old[pm * ld + b]);
        sum += expf((dl - d1) / T);
        mx2 = fmaxf(mx2, (i == istar) ? -D : dl);
      }
      out_ind = T * logf(sum) + d1;
      G = logaddexp_t(out_ind, -D, T);
      float sum2 = 0.f;
      for (int64_t i = p0; i < p1; ++i) {
        const int64_t pm = w.parents_msg[i], pv = w.parents_vs[i];
        const float dl = (S[(pv + 1) * ld + b] - m_old[(pm + 1) * ld + b]) -
                         (S[pv * ld + b] - m_old[pm * ld + b]);
        sum2 += expf((((i == istar) ? -D : dl) - mx2) / T);
      }
      out_star = -(T * logf(sum2) + mx2);
    }
    for (int64_t i = p0; i < p1; ++i) {
      const int64_t pm = w.parents_msg[i], pv = w.parents_vs[i];
      float x;
      if (kSumProduct) {
        const float dl = (S[(pv + 1) * ld + b] - m_old[(pm + 1) * ld + b]) -
                         (S[pv * ld + b] - m_old[pm * ld + b]);
        x = (i == istar) ? out_star : -logminusexp_t(G, dl, T, 1e-30f);
      } else {
        x = fminf(D, -((i == istar) ? d2 : d1));
      }
      if (single) x = D;  // pool.py:430-450
      dmax = fmaxf(dmax, write_binary_edge(m_old, m_new, pm, ld, b, 0.f, x, d, one_minus_d));
    }
    dmax = fmaxf(dmax, write_binary_edge(m_old, m_new, c, ld, b, 0.f, out_ind, d, one_minus_d));
  }
  publish_delta(deltas, int64_t(b) * delta_stride + delta_off, dmax);
}

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
  int ntie = 0;
  for (int64_t var = L.u; var < L.u_end; var += L.step) {
    const int64_t v0 = var_first_state[var], v1 = var_first_state[var + 1];
    const int64_t k0 = var_ptr[var], k1 = var_ptr[var + 1];
    float best = -INFINITY, second = -INFINITY;
    int arg = 0;
    for (int64_t v = v0; v < v1; ++v) {
      float acc = ev.at(v, b);
      for (int64_t k = k0; k < k1; ++k) acc += m.at(var_edge_msg[k] + (v - v0), b);
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
        float acc = ev.at(v, b);
        for (int64_t k = k0; k < k1; ++k) acc += m.at(var_edge_msg[k] + (v - v0), b);
        sum += expf(acc - best);
      }
      const float lse = best + logf(sum);
      for (int64_t v = v0; v < v1; ++v) {
        float acc = ev.at(v, b);
        for (int64_t k = k0; k < k1; ++k) acc += m.at(var_edge_msg[k] + (v - v0), b);
        marginals[int64_t(b) * num_var_states + v] = expf(acc - lse);
      }
    }
  }
  if (ties != nullptr && ntie > 0) atomicAdd(ties + b, ntie);
}

}  // namespace pgx
