// pgx kernels - K4 / K5: OR, AND and Pool factors.  Part of pgx_kernels.cuh (included in this order).
#pragma once

#include "enum.cuh"

namespace pgx {

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
  int32_t uniform;  // > 0: every factor has exactly this many parents (parent_ptr[f] = f * uniform)
};

// ---------------------------------------------------------------------------
// K4: OR / AND update, closed form from per-factor sums and the two largest
// parent differences (pgmax/factor/logical.py:561-779).
// One thread per (factor, sample).  Factors with <= kRegParents parents (the AND
// factors of the deconvolution graphs have 2) keep the parents' variable->factor
// messages in registers: every message is read once.  Wider factors (ORs with up
// to 180 parents) make two passes over the parents, loading kChunk parents' worth
// of independent gathers at a time.
// ---------------------------------------------------------------------------
constexpr int kRegParents = 4;
constexpr int kParentChunk = 8;

// Arithmetic shared by both paths, exactly App. A.3.
struct LogicalAcc {
  float Sb = 0.f, acc = 0.f, d1 = -INFINITY, d2 = -INFINITY;
  int64_t istar = 0;
  template <bool kSumProduct>
  __device__ __forceinline__ void add(int64_t i, float a_i, float b_i, float T) {
    const float dl = a_i - b_i;
    Sb += b_i;
    acc += kSumProduct ? logaddexp_t(a_i, b_i, T) : fmaxf(b_i, a_i);
    if (dl >= d1) { d2 = d1; d1 = dl; istar = i; }  // first arg-max = LARGEST tied index
    else if (dl > d2) d2 = dl;
  }
  // add() under a predicate, with selects only (no divergent-branch bookkeeping in unrolled loops)
  template <bool kSumProduct>
  __device__ __forceinline__ void add_if(bool on, int64_t i, float a_i, float b_i, float T) {
    const float dl = a_i - b_i;
    const float Sb1 = Sb + b_i;
    const float acc1 = acc + (kSumProduct ? logaddexp_t(a_i, b_i, T) : fmaxf(b_i, a_i));
    const bool ge = on && dl >= d1, gt = on && !(dl >= d1) && dl > d2;
    Sb = on ? Sb1 : Sb;
    acc = on ? acc1 : acc;
    d2 = ge ? d1 : (gt ? dl : d2);
    d1 = ge ? dl : d1;
    istar = ge ? i : istar;
  }
  template <bool kSumProduct>
  __device__ __forceinline__ float child_relevant(float T) const {
    if (kSumProduct) {
      float CR = logminusexp_t(acc, Sb, T, 1e-4f);
      if (T < kTempStabThre) CR = fmaxf(CR, logaddexp_t(Sb + d1, Sb + d2, T));
      return CR;
    }
    return acc + fminf(0.f, d1);
  }
  // message difference (relevant - other) to parent i
  template <bool kSumProduct>
  __device__ __forceinline__ float parent_out(int64_t i, float a_i, float b_i, float ca, float cb,
                                              float T, bool single) const {
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
    return PR - PO;
  }
};

// Groups whose factors all have the same number n <= kRegParents of parents (the AND
// factors of the deconvolution graphs: n = 2).  One thread per (factor, sample), TWO factors
// per iteration: the wiring of both is loaded first, then all their gathers (2 * 2(n + 1)
// message / var-sum pairs in flight), then the closed form of App. A.3 for each.
constexpr int kLogicalUnits = 2;

template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_logical_uniform(BatchMap mp, LogicalDev w, const float* __restrict__ S,
                  const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  UnitLoop L = unit_loop(mp, w.num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int sh = mp.bx_log;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float* mo = m_old + moff;
  float* mn = m_new + moff;
  const float* SL = S + lane_off(mp, a.Vs, L.b);
  const int off = w.off, n = w.uniform;
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  auto write_edge = [&](int64_t pm, float x) {
    const int64_t lo = (off > 0) ? pm : pm - 1;
    dmax = fmaxf(dmax, write_binary_edge(mo, mn, lo, sh, off > 0 ? 0.f : x, off > 0 ? x : 0.f, d, one_minus_d));
  };
  for (int64_t f0 = L.u; f0 < L.u_end; f0 += kLogicalUnits * L.step) {
    int32_t c[kLogicalUnits], cvs[kLogicalUnits], pm[kLogicalUnits][kRegParents], pv[kLogicalUnits][kRegParents];
#pragma unroll
    for (int u = 0; u < kLogicalUnits; ++u) {
      const int64_t f = f0 + u * L.step;
      if (f < L.u_end) {
        c[u] = w.children_msg[f];
        cvs[u] = w.children_vs[f];
#pragma unroll
        for (int j = 0; j < kRegParents; ++j)
          if (j < n) { pm[u][j] = w.parents_msg[f * n + j]; pv[u][j] = w.parents_vs[f * n + j]; }
      }
    }
    float ca[kLogicalUnits], cb[kLogicalUnits], av[kLogicalUnits][kRegParents], bv[kLogicalUnits][kRegParents];
#pragma unroll
    for (int u = 0; u < kLogicalUnits; ++u) {
      if (f0 + u * L.step < L.u_end) {
        ca[u] = SL[int64_t(cvs[u] + off) << sh] - mo[int64_t(c[u] + off) << sh];
        cb[u] = SL[int64_t(cvs[u]) << sh] - mo[int64_t(c[u]) << sh];
#pragma unroll
        for (int j = 0; j < kRegParents; ++j)
          if (j < n) {
            av[u][j] = SL[int64_t(pv[u][j] + off) << sh] - mo[int64_t(pm[u][j] + off) << sh];
            bv[u][j] = SL[int64_t(pv[u][j]) << sh] - mo[int64_t(pm[u][j]) << sh];
          }
      }
    }
#pragma unroll
    for (int u = 0; u < kLogicalUnits; ++u) {
      const int64_t f = f0 + u * L.step;
      if (f < L.u_end) {
        const int64_t p0 = f * n;
        LogicalAcc A;
        A.istar = p0;
#pragma unroll
        for (int j = 0; j < kRegParents; ++j)
          if (j < n) A.add<kSumProduct>(p0 + j, av[u][j], bv[u][j], T);
#pragma unroll
        for (int j = 0; j < kRegParents; ++j)
          if (j < n)
            write_edge(pm[u][j], A.parent_out<kSumProduct>(p0 + j, av[u][j], bv[u][j], ca[u], cb[u], T, n == 1));
        write_edge(c[u], A.child_relevant<kSumProduct>(T) - A.Sb);
      }
    }
  }
  publish_delta(a.deltas, int64_t(L.b) * a.delta_stride + a.delta_off, dmax);
}

// kNarrow: every factor of the group has <= kRegParents parents (the wide path is compiled out,
// which halves the register count and doubles the resident warps).
template <bool kSumProduct, bool kNarrow>
__global__ void __launch_bounds__(kThreads)
k_logical(BatchMap mp, LogicalDev w, const float* __restrict__ S,
          const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  UnitLoop L = unit_loop(mp, w.num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int sh = mp.bx_log;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float* mo = m_old + moff;
  float* mn = m_new + moff;
  const float* SL = S + lane_off(mp, a.Vs, L.b);
  const int off = w.off;
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  // variable->factor message of the wiring's state (at msg index pm / var-state pv) and of
  // the "relevant" state (+off)
  auto q_rel = [&](int64_t pm, int64_t pv) { return SL[(pv + off) << sh] - mo[(pm + off) << sh]; };
  auto q_oth = [&](int64_t pm, int64_t pv) { return SL[pv << sh] - mo[pm << sh]; };
  auto write_edge = [&](int64_t pm, float x) {  // x = message of the "+off" state; the other state gets 0
    const int64_t lo = (off > 0) ? pm : pm - 1;
    dmax = fmaxf(dmax, write_binary_edge(mo, mn, lo, sh, off > 0 ? 0.f : x, off > 0 ? x : 0.f, d, one_minus_d));
  };
  // wiring of one factor; fetched one factor ahead so that its (dependent) index loads
  // overlap the current factor's gathers
  struct Wiring {
    int64_t p0, p1;
    int32_t c, cvs;
    int32_t pm[kRegParents], pv[kRegParents];
  };
  auto load_wiring = [&](int64_t f, Wiring& x) {
    if (w.uniform > 0) { x.p0 = f * w.uniform; x.p1 = x.p0 + w.uniform; }
    else { x.p0 = w.parent_ptr[f]; x.p1 = w.parent_ptr[f + 1]; }
    x.c = w.children_msg[f];
    x.cvs = w.children_vs[f];
    if (x.p1 - x.p0 <= kRegParents) {
#pragma unroll
      for (int j = 0; j < kRegParents; ++j)
        if (x.p0 + j < x.p1) { x.pm[j] = w.parents_msg[x.p0 + j]; x.pv[j] = w.parents_vs[x.p0 + j]; }
    }
  };
  Wiring cur_w, next_w;
  if (L.u < L.u_end) load_wiring(L.u, cur_w);
  for (int64_t f = L.u; f < L.u_end; f += L.step, cur_w = next_w) {
    if (f + L.step < L.u_end) load_wiring(f + L.step, next_w);
    const int64_t p0 = cur_w.p0, p1 = cur_w.p1;
    const int64_t c = cur_w.c, cvs = cur_w.cvs;
    const float ca = q_rel(c, cvs), cb = q_oth(c, cvs);
    const bool single = (p1 - p0) == 1;
    LogicalAcc A;
    A.istar = p0;
    if (kNarrow || p1 - p0 <= kRegParents) {
      float av[kRegParents], bv[kRegParents];
#pragma unroll
      for (int j = 0; j < kRegParents; ++j) {
        if (p0 + j < p1) {
          av[j] = q_rel(cur_w.pm[j], cur_w.pv[j]);
          bv[j] = q_oth(cur_w.pm[j], cur_w.pv[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < kRegParents; ++j)
        if (p0 + j < p1) A.add<kSumProduct>(p0 + j, av[j], bv[j], T);
#pragma unroll
      for (int j = 0; j < kRegParents; ++j)
        if (p0 + j < p1)
          write_edge(cur_w.pm[j], A.parent_out<kSumProduct>(p0 + j, av[j], bv[j], ca, cb, T, single));
    } else if (!kNarrow) {
      // Pass 1: sums in ascending parent order, first / second max of the differences.
      int64_t i = p0;
      for (; i + kParentChunk <= p1; i += kParentChunk) {
        float av[kParentChunk], bv[kParentChunk];
#pragma unroll
        for (int j = 0; j < kParentChunk; ++j) {
          const int64_t pm = w.parents_msg[i + j], pv = w.parents_vs[i + j];
          av[j] = q_rel(pm, pv);
          bv[j] = q_oth(pm, pv);
        }
#pragma unroll
        for (int j = 0; j < kParentChunk; ++j) A.add<kSumProduct>(i + j, av[j], bv[j], T);
      }
      for (; i < p1; ++i) {
        const int64_t pm = w.parents_msg[i], pv = w.parents_vs[i];
        A.add<kSumProduct>(i, q_rel(pm, pv), q_oth(pm, pv), T);
      }
      // Pass 2: outgoing messages to the parents.
      i = p0;
      for (; i + kParentChunk <= p1; i += kParentChunk) {
        int64_t pm[kParentChunk];
        float av[kParentChunk], bv[kParentChunk];
#pragma unroll
        for (int j = 0; j < kParentChunk; ++j) {
          pm[j] = w.parents_msg[i + j];
          const int64_t pv = w.parents_vs[i + j];
          av[j] = q_rel(pm[j], pv);
          bv[j] = q_oth(pm[j], pv);
        }
#pragma unroll
        for (int j = 0; j < kParentChunk; ++j)
          write_edge(pm[j], A.parent_out<kSumProduct>(i + j, av[j], bv[j], ca, cb, T, single));
      }
      for (; i < p1; ++i) {
        const int64_t pm = w.parents_msg[i], pv = w.parents_vs[i];
        write_edge(pm, A.parent_out<kSumProduct>(i, q_rel(pm, pv), q_oth(pm, pv), ca, cb, T, single));
      }
    }
    write_edge(c, A.child_relevant<kSumProduct>(T) - A.Sb);
  }
  publish_delta(a.deltas, int64_t(L.b) * a.delta_stride + a.delta_off, dmax);
}

// ---------------------------------------------------------------------------
// K4-pull: OR / AND update for full sample tiles (TW = 32: a warp = the 32 samples of ONE
// factor, every index warp-uniform) that needs the variable-sum array only for HIGH-degree
// variables.  For a variable with one or two incident edges the kernel re-derives
// S_v = ev_v + (incident messages in ascending message index) itself - the same additions
// in the same order as k_var_sums, so results are bit-identical - which removes the S
// write + gather and the message re-read of k_var_sums for those variables (in the
// deconvolution graphs: every SW and X variable, 2/3 of all var-states); k_var_sums then
// only runs over the listed high-degree var-states (S and W).
// Wiring per edge (EdgeW, one 16-byte load): msg = message index of the state the
// reference's wiring points at (state 0 for OR, state 1 for AND), vs = var-state of that
// state, other = message index (same state) of the variable's only other edge, -1 if the
// variable has no other edge, -2 if its sum is to be read from S.
// All loads of a factor are issued before the first use (no data-dependent branch between
// them); addresses are 32-bit offsets from per-lane bases.
// ---------------------------------------------------------------------------
struct EdgeW {
  int32_t msg, vs, other, pad;
};

struct LogicalPullDev {
  int64_t num_factors;
  const int32_t* parent_ptr;  // [F + 1] (null when uniform)
  const EdgeW* parents;       // [P]
  const EdgeW* children;      // [F]
  int32_t off;                // +1 OR, -1 AND
  int32_t uniform;            // > 0: every factor has this many parents
};

// What stays live per edge between the loads and their use: 8 registers.
struct EdgeIn {
  float m_p, m_r;  // old messages: pointed state, relevant (+off) state
  float a_p, a_r;  // S (kind 0) or evidence
  float o_p, o_r;  // the other edge's messages (kinds 2, 3)
  int32_t msg;
  int32_t kind;    // 0: sums from S; 1: no other edge; 2: other edge first; 3: own edge first
};

// kBin: the message arrays are in binary-difference storage (one float x = n1 - n0 per edge,
// every edge of the graph has two states, edge e holds message rows 2e, 2e + 1; see
// bin_expand): row of an edge = msg >> 1, and (pointed, relevant) = (state 0, state 1) for
// off = +1, (state 1, state 0) for off = -1.
template <bool kBin>
__device__ __forceinline__ size_t msg_rows(const RunArgs& a) { return kBin ? size_t(a.Es) >> 1 : size_t(a.Es); }

// Sum-product closed forms can return an infinite difference (an empty logminusexp); the
// update then leaves BOTH states of the edge at the clip value -1e32 (inf - inf = NaN, and
// fmaxf(NaN, -1e32) = -1e32), the one normalised pair whose maximum is not 0.  The stored
// difference encodes it as NaN (kFloor variants; max-product never produces it).
// load_msg only LOADS (kBin: the raw difference goes to m_p); expand_msg, called on the
// consumer side (edge_q), turns it into the two states - no arithmetic sits between the loads
// of a batch of edges, so they all stay in flight together.
template <bool kBin, bool kFloor>
__device__ __forceinline__ void expand_msg(int off, float& m_p, float& m_r) {
  if (!kBin) return;
  const float x = m_p;
  const float xs = off > 0 ? x : -x;  // relevant - pointed
  m_p = fminf(-xs, 0.f);
  m_r = fminf(xs, 0.f);
  if (kFloor && x != x) m_p = m_r = kMsgNegInf;
}

template <bool kBin, bool kFloor>
__device__ __forceinline__ void load_msg(const float* __restrict__ mo, int32_t msg, int off, float& m_p, float& m_r) {
  if (kBin) {
    m_p = mo[(uint32_t(msg) >> 1) << 5];
    m_r = 0.f;
  } else {
    m_p = mo[uint32_t(msg) << 5];
    m_r = mo[uint32_t(msg + off) << 5];
  }
}

// Issues the (up to) 6 loads of an edge (4 in binary-difference storage); no data-dependent branch.
template <bool kBin, bool kFloor>
__device__ __forceinline__ EdgeIn load_edge(const EdgeW& e, int off, const float* __restrict__ mo,
                                            const float* __restrict__ evq, int esh,
                                            const float* __restrict__ SL) {
  EdgeIn r;
  r.msg = e.msg;
  r.kind = e.other == -2 ? 0 : (e.other == -1 ? 1 : (e.other < e.msg ? 2 : 3));
  load_msg<kBin, kFloor>(mo, e.msg, off, r.m_p, r.m_r);
  const bool from_s = e.other == -2;
  r.a_p = *(from_s ? SL + (uint32_t(e.vs) << 5) : evq + (uint32_t(e.vs) << esh));
  r.a_r = *(from_s ? SL + (uint32_t(e.vs + off) << 5) : evq + (uint32_t(e.vs + off) << esh));
  // unconditional: an edge without a second edge re-reads its own row (a hit) and edge_q ignores it
  load_msg<kBin, kFloor>(mo, e.other >= 0 ? e.other : e.msg, off, r.o_p, r.o_r);
  return r;
}
// variable -> factor messages (pointed state, relevant state): S - m with S accumulated from
// the evidence in ascending message index
template <bool kBin, bool kFloor>
__device__ __forceinline__ void edge_q(EdgeIn& r, int off, float& q_p, float& q_r) {
  // branch-free (selects): the same additions in the same order as the nested ifs it replaces
  expand_msg<kBin, kFloor>(off, r.m_p, r.m_r);
  expand_msg<kBin, kFloor>(off, r.o_p, r.o_r);
  const bool from_s = r.kind == 0, two = r.kind >= 2, other_first = r.kind == 2;
  const float f_p = other_first ? r.o_p : r.m_p, f_r = other_first ? r.o_r : r.m_r;
  const float g_p = other_first ? r.m_p : r.o_p, g_r = other_first ? r.m_r : r.o_r;
  float s_p = from_s ? r.a_p : r.a_p + f_p;
  float s_r = from_s ? r.a_r : r.a_r + f_r;
  s_p = two ? s_p + g_p : s_p;
  s_r = two ? s_r + g_r : s_r;
  q_p = s_p - r.m_p;
  q_r = s_r - r.m_r;
}

// new message (x at the relevant state, 0 at the pointed state): damping, normalisation,
// clip, store; returns max|new - old| when kDelta
template <bool kDelta, bool kBin, bool kFloor>
__device__ __forceinline__ float store_edge(float* __restrict__ mn, int off, const EdgeIn& r, float x, float d,
                                            float one_minus_d) {
  float n_p = damp(r.m_p, 0.f, d, one_minus_d), n_r = damp(r.m_r, x, d, one_minus_d);
  const float mx = fmaxf(n_p, n_r);
  n_p = fmaxf(n_p - mx, kMsgNegInf);
  n_r = fmaxf(n_r - mx, kMsgNegInf);
  if (kBin) {  // one of n_p, n_r is the exact zero: the difference loses nothing
    float xd = off > 0 ? n_r - n_p : n_p - n_r;
    if (kFloor && fmaxf(n_p, n_r) < 0.f) xd = __int_as_float(0x7fc00000);  // both states at the floor
    mn[(uint32_t(r.msg) >> 1) << 5] = xd;
  } else {
    mn[uint32_t(r.msg) << 5] = n_p;
    mn[uint32_t(r.msg + off) << 5] = n_r;
  }
  return kDelta ? fmaxf(fabsf(n_p - r.m_p), fabsf(n_r - r.m_r)) : 0.f;
}

// Factors with <= NP parents (AND factors: NP = 2), everything in registers, U factors per
// warp iteration (their loads are all in flight together).  kUniform: every factor has
// exactly NP parents.
template <bool kSumProduct, bool kDelta, int NP, int U, bool kUniform, bool kBin>
__global__ void __launch_bounds__(kThreads)
k_logical_pull_small(int batch, LogicalPullDev w, View ev, const float* __restrict__ S,
                     const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y * 32 + lane;
  if (b >= batch) return;
  const size_t tile = blockIdx.y;
  const float* mo = m_old + tile * msg_rows<kBin>(a) * 32 + lane;
  float* mn = m_new + tile * msg_rows<kBin>(a) * 32 + lane;
  const float* SL = S + tile * size_t(a.Vs) * 32 + lane;
  const float* evq = ev.kind == 1 ? ev.p + tile * size_t(ev.n_rows) * 32 + lane : ev.p;
  const int esh = ev.kind == 1 ? 5 : 0;
  const int off = w.off;
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  const int64_t gwarp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float dmax = 0.f;
  for (int64_t f0 = gwarp; f0 < w.num_factors; f0 += U * nwarps) {
    EdgeIn ce[U], pe[U][NP];
    int np[U];
    int64_t p0[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t f = f0 + u * nwarps;
      np[u] = 0;
      p0[u] = 0;
      if (f < w.num_factors) {
        if (kUniform) { p0[u] = f * NP; np[u] = NP; }
        else { p0[u] = w.parent_ptr[f]; np[u] = int(w.parent_ptr[f + 1] - p0[u]); }
        ce[u] = load_edge<kBin, kSumProduct>(w.children[f], off, mo, evq, esh, SL);
#pragma unroll
        for (int j = 0; j < NP; ++j)
          if (kUniform || j < np[u]) pe[u][j] = load_edge<kBin, kSumProduct>(w.parents[p0[u] + j], off, mo, evq, esh, SL);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (f0 + u * nwarps < w.num_factors) {
        float c_p, c_r, q_p[NP], q_r[NP];
        edge_q<kBin, kSumProduct>(ce[u], off, c_p, c_r);
        LogicalAcc A;
        A.istar = p0[u];
#pragma unroll
        for (int j = 0; j < NP; ++j)
          if (kUniform || j < np[u]) {
            edge_q<kBin, kSumProduct>(pe[u][j], off, q_p[j], q_r[j]);
            A.add<kSumProduct>(p0[u] + j, q_r[j], q_p[j], T);
          }
#pragma unroll
        for (int j = 0; j < NP; ++j)
          if (kUniform || j < np[u]) {
            const float x = A.parent_out<kSumProduct>(p0[u] + j, q_r[j], q_p[j], c_r, c_p, T, np[u] == 1);
            dmax = fmaxf(dmax, store_edge<kDelta, kBin, kSumProduct>(mn, off, pe[u][j], x, d, one_minus_d));
          }
        dmax = fmaxf(dmax, store_edge<kDelta, kBin, kSumProduct>(mn, off, ce[u], A.child_relevant<kSumProduct>(T) - A.Sb, d, one_minus_d));
      }
    }
  }
  if (kDelta) publish_delta(a.deltas, int64_t(b) * a.delta_stride + a.delta_off, dmax);
}

// The same update for UNIFORM groups (every factor has exactly NP parents) with the wiring
// staged in shared memory: a CTA owns a contiguous range of `chunk` factors and fetches their
// wiring (children and parents are contiguous arrays) with one coalesced sweep.  In the
// kernel above every warp iteration is two dependent round trips - wiring from L2, then the
// rows it points at from HBM; here the first one is paid once per CTA.  Same arithmetic.
constexpr int kPullChunk = 256;  // factors per CTA

template <bool kSumProduct, bool kDelta, int NP, int U, bool kBin>
__global__ void __launch_bounds__(kThreads)
k_logical_pull_small_staged(int batch, LogicalPullDev w, View ev, const float* __restrict__ S,
                            const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  __shared__ int4 wsm_child[kPullChunk];
  __shared__ int4 wsm_par[kPullChunk * NP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y * 32 + lane;
  const int64_t fbase = int64_t(blockIdx.x) * kPullChunk;
  const int nfac = int(min(int64_t(kPullChunk), w.num_factors - fbase));
  {
    const int4* gc = reinterpret_cast<const int4*>(w.children + fbase);
    const int4* gp = reinterpret_cast<const int4*>(w.parents + fbase * NP);
    for (int t = threadIdx.x; t < nfac; t += blockDim.x) wsm_child[t] = gc[t];
    for (int t = threadIdx.x; t < nfac * NP; t += blockDim.x) wsm_par[t] = gp[t];
  }
  __syncthreads();
  if (b >= batch) return;
  const size_t tile = blockIdx.y;
  const float* mo = m_old + tile * msg_rows<kBin>(a) * 32 + lane;
  float* mn = m_new + tile * msg_rows<kBin>(a) * 32 + lane;
  const float* SL = S + tile * size_t(a.Vs) * 32 + lane;
  const float* evq = ev.kind == 1 ? ev.p + tile * size_t(ev.n_rows) * 32 + lane : ev.p;
  const int esh = ev.kind == 1 ? 5 : 0;
  const int off = w.off;
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  const EdgeW* sc = reinterpret_cast<const EdgeW*>(wsm_child);
  const EdgeW* sp = reinterpret_cast<const EdgeW*>(wsm_par);
  constexpr int kWarps = kThreads / 32;
  float dmax = 0.f;
  for (int j0 = warp * U; j0 < nfac; j0 += kWarps * U) {
    EdgeIn ce[U], pe[U][NP];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (j0 + u < nfac) {
        ce[u] = load_edge<kBin, kSumProduct>(sc[j0 + u], off, mo, evq, esh, SL);
#pragma unroll
        for (int j = 0; j < NP; ++j) pe[u][j] = load_edge<kBin, kSumProduct>(sp[(j0 + u) * NP + j], off, mo, evq, esh, SL);
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (j0 + u < nfac) {
        const int64_t p0 = (fbase + j0 + u) * NP;
        float c_p, c_r, q_p[NP], q_r[NP];
        edge_q<kBin, kSumProduct>(ce[u], off, c_p, c_r);
        LogicalAcc A;
        A.istar = p0;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          edge_q<kBin, kSumProduct>(pe[u][j], off, q_p[j], q_r[j]);
          A.add<kSumProduct>(p0 + j, q_r[j], q_p[j], T);
        }
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          const float x = A.parent_out<kSumProduct>(p0 + j, q_r[j], q_p[j], c_r, c_p, T, NP == 1);
          dmax = fmaxf(dmax, store_edge<kDelta, kBin, kSumProduct>(mn, off, pe[u][j], x, d, one_minus_d));
        }
        dmax = fmaxf(dmax, store_edge<kDelta, kBin, kSumProduct>(mn, off, ce[u], A.child_relevant<kSumProduct>(T) - A.Sb, d, one_minus_d));
      }
  }
  if (kDelta) publish_delta(a.deltas, int64_t(b) * a.delta_stride + a.delta_off, dmax);
}

// Any number of parents (OR factors with up to hundreds): two passes over the parents.  The
// wiring of 32 parents is fetched with ONE coalesced load (lane j holds parent i + j) and
// handed out by shuffles; the parents' loads are issued kParentChunk at a time.  All lanes
// stay alive for the shuffles; lanes beyond the batch read a valid sample and store nothing.
template <bool kSumProduct, bool kDelta, bool kBin>
__global__ void __launch_bounds__(128)
k_logical_pull_wide(int batch, LogicalPullDev w, View ev, const float* __restrict__ S,
                    const float* __restrict__ m_old, float* __restrict__ m_new, RunArgs a) {
  const int lane = threadIdx.x & 31;
  const int b_raw = blockIdx.y * 32 + lane;
  const bool live = b_raw < batch;
  const int ll = live ? lane : 0;  // dead lanes shadow sample 0 of the tile (always valid)
  const size_t tile = blockIdx.y;
  const float* mo = m_old + tile * msg_rows<kBin>(a) * 32 + ll;
  float* mn = m_new + tile * msg_rows<kBin>(a) * 32 + ll;
  const float* SL = S + tile * size_t(a.Vs) * 32 + ll;
  const float* evq = ev.kind == 1 ? ev.p + tile * size_t(ev.n_rows) * 32 + ll : ev.p;
  const int esh = ev.kind == 1 ? 5 : 0;
  const int off = w.off;
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  const int64_t gwarp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float dmax = 0.f;
  for (int64_t f = gwarp; f < w.num_factors; f += nwarps) {
    int64_t p0, p1;
    if (w.uniform > 0) { p0 = f * w.uniform; p1 = p0 + w.uniform; }
    else { p0 = w.parent_ptr[f]; p1 = w.parent_ptr[f + 1]; }
    EdgeIn ce = load_edge<kBin, kSumProduct>(w.children[f], off, mo, evq, esh, SL);
    LogicalAcc A;
    A.istar = p0;
    const bool single = (p1 - p0) == 1;
    float c_p = 0.f, c_r = 0.f;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1) edge_q<kBin, kSumProduct>(ce, off, c_p, c_r);
      EdgeW mine = (p0 + lane < p1) ? w.parents[p0 + lane] : EdgeW{0, 0, -1, 0};
      for (int64_t i32 = p0; i32 < p1; i32 += 32) {
        const EdgeW held = mine;
        if (i32 + 32 + lane < p1) mine = w.parents[i32 + 32 + lane];  // next 32, in flight during this block
        const int n32 = int(min(int64_t(32), p1 - i32));
#pragma unroll 1
        for (int c0 = 0; c0 < n32; c0 += kParentChunk) {
          EdgeIn r[kParentChunk];
#pragma unroll
          for (int j = 0; j < kParentChunk; ++j) {
            EdgeW e;
            e.msg = __shfl_sync(0xffffffffu, held.msg, (c0 + j) & 31);
            e.vs = __shfl_sync(0xffffffffu, held.vs, (c0 + j) & 31);
            e.other = __shfl_sync(0xffffffffu, held.other, (c0 + j) & 31);
            if (c0 + j < n32) r[j] = load_edge<kBin, kSumProduct>(e, off, mo, evq, esh, SL);
          }
#pragma unroll
          for (int j = 0; j < kParentChunk; ++j)
            if (c0 + j < n32) {
              float q_p, q_r;
              edge_q<kBin, kSumProduct>(r[j], off, q_p, q_r);
              const int64_t i = i32 + c0 + j;
              if (pass == 0) {
                A.add<kSumProduct>(i, q_r, q_p, T);
              } else {
                const float x = A.parent_out<kSumProduct>(i, q_r, q_p, c_r, c_p, T, single);
                if (live) dmax = fmaxf(dmax, store_edge<kDelta, kBin, kSumProduct>(mn, off, r[j], x, d, one_minus_d));
              }
            }
        }
      }
    }
    if (live)
      dmax = fmaxf(dmax, store_edge<kDelta, kBin, kSumProduct>(mn, off, ce, A.child_relevant<kSumProduct>(T) - A.Sb, d, one_minus_d));
  }
  if (kDelta && live) publish_delta(a.deltas, int64_t(b_raw) * a.delta_stride + a.delta_off, dmax);
}

// The wide update split in two launches, so that only the part that must be serial is:
//   pass 1 (k_logical_wide_reduce): one warp per (factor, sample tile) walks the parents once,
//     accumulating the sums in ascending parent order and the two largest differences, writes
//     the child's message and the factor's aggregates [F][8][32 samples] (tile-blocked);
//   pass 2 (k_logical_wide_emit): one warp per (PARENT, sample tile) - fully parallel,
//     bandwidth-bound - re-derives its own variable -> factor message and emits the message to
//     the parent from the aggregates.
// Same arithmetic as k_logical_pull_wide (LogicalAcc), hence bit-identical.
constexpr int kAggRows = 8;  // acc, Sb, d1, d2, istar - p0 (int bits), c_p, c_r, unused

#ifndef PGX_REDUCE_CHUNK
#define PGX_REDUCE_CHUNK 8
#endif
#ifndef PGX_REDUCE_WARPS
#define PGX_REDUCE_WARPS 22
#endif
constexpr int kReduceChunk = PGX_REDUCE_CHUNK;  // parents gathered together by a reduce warp

template <bool kSumProduct, bool kDelta, bool kBin>
__global__ void __launch_bounds__(32, PGX_REDUCE_WARPS)  // one warp per CTA: a serial chain holds only its own warp
k_logical_wide_reduce(int batch, LogicalPullDev w, View ev, const float* __restrict__ S,
                      const float* __restrict__ m_old, float* __restrict__ m_new, float* __restrict__ agg,
                      RunArgs a) {
  const int lane = threadIdx.x & 31;
  const int b_raw = blockIdx.y * 32 + lane;
  const bool live = b_raw < batch;
  const int ll = live ? lane : 0;
  const size_t tile = blockIdx.y;
  const float* mo = m_old + tile * msg_rows<kBin>(a) * 32 + ll;
  float* mn = m_new + tile * msg_rows<kBin>(a) * 32 + ll;
  const float* SL = S + tile * size_t(a.Vs) * 32 + ll;
  const float* evq = ev.kind == 1 ? ev.p + tile * size_t(ev.n_rows) * 32 + ll : ev.p;
  const int esh = ev.kind == 1 ? 5 : 0;
  float* aggL = agg + tile * size_t(w.num_factors) * kAggRows * 32 + ll;
  const int off = w.off;
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  const int64_t gwarp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float dmax = 0.f;
  for (int64_t f = gwarp; f < w.num_factors; f += nwarps) {
    int64_t p0, p1;
    if (w.uniform > 0) { p0 = f * w.uniform; p1 = p0 + w.uniform; }
    else { p0 = w.parent_ptr[f]; p1 = w.parent_ptr[f + 1]; }
    EdgeIn ce = load_edge<kBin, kSumProduct>(w.children[f], off, mo, evq, esh, SL);
    LogicalAcc A;
    A.istar = p0;
    EdgeW mine = (p0 + lane < p1) ? w.parents[p0 + lane] : EdgeW{0, 0, -1, 0};
    for (int64_t i32 = p0; i32 < p1; i32 += 32) {
      const EdgeW held = mine;
      if (i32 + 32 + lane < p1) mine = w.parents[i32 + 32 + lane];
      const int n32 = int(min(int64_t(32), p1 - i32));
#pragma unroll 1
      for (int c0 = 0; c0 < n32; c0 += kReduceChunk) {
        // the loads of a chunk are unconditional (slots past the last parent re-read the last
        // parent): no divergent-branch bookkeeping between them, the whole chunk stays in registers
        EdgeIn r[kReduceChunk];
#pragma unroll
        for (int j = 0; j < kReduceChunk; ++j) {
          const int src = min(c0 + j, n32 - 1);
          EdgeW e;
          e.msg = __shfl_sync(0xffffffffu, held.msg, src);
          e.vs = __shfl_sync(0xffffffffu, held.vs, src);
          e.other = __shfl_sync(0xffffffffu, held.other, src);
          r[j] = load_edge<kBin, kSumProduct>(e, off, mo, evq, esh, SL);
        }
#pragma unroll
        for (int j = 0; j < kReduceChunk; ++j) {
          float q_p, q_r;
          edge_q<kBin, kSumProduct>(r[j], off, q_p, q_r);
          A.add_if<kSumProduct>(c0 + j < n32, i32 + c0 + j, q_r, q_p, T);
        }
      }
    }
    float c_p, c_r;
    edge_q<kBin, kSumProduct>(ce, off, c_p, c_r);
    if (live) {
      float* g = aggL + size_t(f) * kAggRows * 32;
      g[0] = A.acc; g[32] = A.Sb; g[64] = A.d1; g[96] = A.d2;
      g[128] = __int_as_float(int(A.istar - p0)); g[160] = c_p; g[192] = c_r;
      dmax = fmaxf(dmax, store_edge<kDelta, kBin, kSumProduct>(mn, off, ce, A.child_relevant<kSumProduct>(T) - A.Sb, d, one_minus_d));
    }
  }
  if (kDelta && live) publish_delta(a.deltas, int64_t(b_raw) * a.delta_stride + a.delta_off, dmax);
}

// parent_factor[i] = factor of parent i (ascending).  kEmitUnits parents per warp iteration.
constexpr int kEmitUnits = 2;

template <bool kSumProduct, bool kDelta, bool kBin>
__global__ void __launch_bounds__(kThreads)
k_logical_wide_emit(int batch, LogicalPullDev w, const int32_t* __restrict__ parent_factor, int64_t num_parents,
                    View ev, const float* __restrict__ S, const float* __restrict__ m_old,
                    float* __restrict__ m_new, const float* __restrict__ agg, RunArgs a) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y * 32 + lane;
  if (b >= batch) return;
  const size_t tile = blockIdx.y;
  const float* mo = m_old + tile * msg_rows<kBin>(a) * 32 + lane;
  float* mn = m_new + tile * msg_rows<kBin>(a) * 32 + lane;
  const float* SL = S + tile * size_t(a.Vs) * 32 + lane;
  const float* evq = ev.kind == 1 ? ev.p + tile * size_t(ev.n_rows) * 32 + lane : ev.p;
  const int esh = ev.kind == 1 ? 5 : 0;
  const float* aggL = agg + tile * size_t(w.num_factors) * kAggRows * 32 + lane;
  const int off = w.off;
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  const int64_t gwarp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float dmax = 0.f;
  for (int64_t i0 = gwarp; i0 < num_parents; i0 += kEmitUnits * nwarps) {
    EdgeIn r[kEmitUnits];
    float g[kEmitUnits][7];
    int64_t p0[kEmitUnits], p1[kEmitUnits];
#pragma unroll
    for (int u = 0; u < kEmitUnits; ++u) {
      const int64_t i = i0 + u * nwarps;
      if (i < num_parents) {
        const int f = parent_factor[i];
        if (w.uniform > 0) { p0[u] = int64_t(f) * w.uniform; p1[u] = p0[u] + w.uniform; }
        else { p0[u] = w.parent_ptr[f]; p1[u] = w.parent_ptr[f + 1]; }
        r[u] = load_edge<kBin, kSumProduct>(w.parents[i], off, mo, evq, esh, SL);
        const float* gp = aggL + size_t(f) * kAggRows * 32;
#pragma unroll
        for (int k = 0; k < 7; ++k) g[u][k] = gp[k * 32];
      }
    }
#pragma unroll
    for (int u = 0; u < kEmitUnits; ++u) {
      const int64_t i = i0 + u * nwarps;
      if (i < num_parents) {
        LogicalAcc A;
        A.acc = g[u][0]; A.Sb = g[u][1]; A.d1 = g[u][2]; A.d2 = g[u][3];
        A.istar = p0[u] + __float_as_int(g[u][4]);
        float q_p, q_r;
        edge_q<kBin, kSumProduct>(r[u], off, q_p, q_r);
        const float x = A.parent_out<kSumProduct>(i, q_r, q_p, g[u][6], g[u][5], T, p1[u] - p0[u] == 1);
        dmax = fmaxf(dmax, store_edge<kDelta, kBin, kSumProduct>(mn, off, r[u], x, d, one_minus_d));
      }
    }
  }
  if (kDelta) publish_delta(a.deltas, int64_t(b) * a.delta_stride + a.delta_off, dmax);
}

constexpr int kEmitChunk = 512;  // parents per CTA of the staged emit kernel

template <bool kSumProduct, bool kDelta, bool kBin>
__global__ void __launch_bounds__(kThreads)
k_logical_wide_emit_staged(int batch, LogicalPullDev w, const int32_t* __restrict__ parent_factor, int64_t num_parents,
                    View ev, const float* __restrict__ S, const float* __restrict__ m_old,
                    float* __restrict__ m_new, const float* __restrict__ agg, RunArgs a) {
  // a CTA owns the contiguous parent range [pbase, pbase + kEmitChunk): wiring and factor index
  // are staged in shared memory with one coalesced sweep (see k_logical_pull_small_staged)
  __shared__ int4 wsm[kEmitChunk];
  __shared__ int32_t fsm[kEmitChunk];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y * 32 + lane;
  const int64_t pbase = int64_t(blockIdx.x) * kEmitChunk;
  const int npar = int(min(int64_t(kEmitChunk), num_parents - pbase));
  for (int t = threadIdx.x; t < npar; t += blockDim.x) {
    wsm[t] = reinterpret_cast<const int4*>(w.parents + pbase)[t];
    fsm[t] = parent_factor[pbase + t];
  }
  __syncthreads();
  if (b >= batch) return;
  const EdgeW* sw = reinterpret_cast<const EdgeW*>(wsm);
  const size_t tile = blockIdx.y;
  const float* mo = m_old + tile * msg_rows<kBin>(a) * 32 + lane;
  float* mn = m_new + tile * msg_rows<kBin>(a) * 32 + lane;
  const float* SL = S + tile * size_t(a.Vs) * 32 + lane;
  const float* evq = ev.kind == 1 ? ev.p + tile * size_t(ev.n_rows) * 32 + lane : ev.p;
  const int esh = ev.kind == 1 ? 5 : 0;
  const float* aggL = agg + tile * size_t(w.num_factors) * kAggRows * 32 + lane;
  const int off = w.off;
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  constexpr int kWarps = kThreads / 32;
  float dmax = 0.f;
  for (int j0 = warp * kEmitUnits; j0 < npar; j0 += kWarps * kEmitUnits) {
    EdgeIn r[kEmitUnits];
    float g[kEmitUnits][7];
    int64_t p0[kEmitUnits], p1[kEmitUnits];
#pragma unroll
    for (int u = 0; u < kEmitUnits; ++u) {
      const int j = j0 + u;
      if (j < npar) {
        const int f = fsm[j];
        if (w.uniform > 0) { p0[u] = int64_t(f) * w.uniform; p1[u] = p0[u] + w.uniform; }
        else { p0[u] = w.parent_ptr[f]; p1[u] = w.parent_ptr[f + 1]; }
        r[u] = load_edge<kBin, kSumProduct>(sw[j], off, mo, evq, esh, SL);
        const float* gp = aggL + size_t(f) * kAggRows * 32;
#pragma unroll
        for (int k = 0; k < 7; ++k) g[u][k] = gp[k * 32];
      }
    }
#pragma unroll
    for (int u = 0; u < kEmitUnits; ++u) {
      const int64_t i = pbase + j0 + u;
      if (j0 + u < npar) {
        LogicalAcc A;
        A.acc = g[u][0]; A.Sb = g[u][1]; A.d1 = g[u][2]; A.d2 = g[u][3];
        A.istar = p0[u] + __float_as_int(g[u][4]);
        float q_p, q_r;
        edge_q<kBin, kSumProduct>(r[u], off, q_p, q_r);
        const float x = A.parent_out<kSumProduct>(i, q_r, q_p, g[u][6], g[u][5], T, p1[u] - p0[u] == 1);
        dmax = fmaxf(dmax, store_edge<kDelta, kBin, kSumProduct>(mn, off, r[u], x, d, one_minus_d));
      }
    }
  }
  if (kDelta) publish_delta(a.deltas, int64_t(b) * a.delta_stride + a.delta_off, dmax);
}

// ---------------------------------------------------------------------------
// K4-fused: OR factors whose every parent is the degree-2 child of exactly one two-parent AND
// factor (the deconvolution model of examples/pmp_binary_deconvolution.ipynb: X = OR_i(S_i AND
// W_i); every SW variable sits between one AND factor and one OR factor).  The AND update of
// SW_i and the OR update both need the SAME rows - the two messages into SW_i and its evidence -
// so ONE kernel does both per OR factor and sample tile, and every row is read once:
//
//   phase A (all warps, parent-parallel): per parent i load the two messages of SW_i, its
//            evidence, the two messages and variable sums of the AND factor's parents; finish
//            the AND factor (three stores); park SW_i's variable -> OR-factor messages
//            (a_i, b_i) in shared memory;
//   phase B (warp 0): the OR factor's sums and top-2 differences in ASCENDING parent order from
//            shared memory (the serial order of the reference's segment sums, ~12 instructions
//            per parent and no memory latency), the child's message;
//   phase C (all warps): the OR factor's messages to its parents.
//
// Same helpers (edge_q / LogicalAcc / store_edge) and the same operations in the same order as
// k_logical_pull_small + k_logical_wide_reduce / _emit: bit-identical (tested); the edge kinds are
// compile-time constants here (no selects), nothing is gathered twice, one launch replaces four
// and the second stream.  Binary-difference storage, full sample tiles.  The pairing is found at
// plan time (pgx.cu); graphs that do not have it keep the separate kernels.
// ---------------------------------------------------------------------------
struct FusedW {
  int32_t mO, mA;      // rows (edge indices) of the OR-parent edge and of the AND-child edge of SW_i
  int32_t ev0;         // var-state of SW_i's state 0
  int32_t ms, mw;      // rows of the AND factor's two parent edges
  int32_t Ss0, Sw0;    // var-states (state 0) of those parents: their sums come from S
  int32_t f_and;       // the AND factor (its parents have the global parent indices 2 f, 2 f + 1)
};

struct OrAndFusedDev {
  int64_t num_or;
  const int32_t* parent_ptr;  // [num_or + 1]
  const FusedW* w;            // [P], OR parent order
  const EdgeW* or_children;   // [num_or]
  int32_t max_parents;
};

constexpr int kFusedWarps = 16;
// dynamic shared memory: wiring [P] (32 B each), parked (a_i, b_i, old OR message) [P][3][32], aggregates [20][32]
__host__ __device__ constexpr size_t orand_fused_smem(int max_parents) {
  return size_t(max_parents) * sizeof(FusedW) + (size_t(max_parents) * 3 + 20) * 32 * sizeof(float);
}

// kPack > 1: the batch's last, partial sample tile.  With 32 / kPack or fewer samples in it a warp
// would idle most of its lanes, so kPack OR factors share one CTA: lane = (factor slot, sample),
// every per-factor quantity (parent range, wiring, chains of phase B) becomes per-lane instead of
// warp-uniform, loops run to the longest of the CTA's factors under a predicate.  Same operations
// per (factor, sample) in the same order: bit-identical to the kPack = 1 kernel (tested).
template <bool kSumProduct, bool kDelta, int kPack>
__global__ void __launch_bounds__(kFusedWarps * 32, 2)
k_or_and_fused(int batch, int tile0, OrAndFusedDev g, View ev, const float* __restrict__ S, const float* __restrict__ m_old,
               float* __restrict__ m_new, RunArgs a) {
  extern __shared__ __align__(16) unsigned char fz_raw[];
  constexpr bool kPacked = kPack > 1;
  constexpr int kGroup = 32 / kPack;  // samples per factor slot
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int smp = kPacked ? (lane & (kGroup - 1)) : lane;
  const size_t tile = size_t(tile0) + blockIdx.y;
  const int b = int(tile) * 32 + smp;
  // lanes beyond the batch stay alive for the barriers: they work on the (allocated, padded)
  // sample slots of the last tile and publish no delta
  const float* mo = m_old + tile * (size_t(a.Es) >> 1) * 32 + smp;
  float* mn = m_new + tile * (size_t(a.Es) >> 1) * 32 + smp;
  const float* SL = S + tile * size_t(a.Vs) * 32 + smp;
  const float* evq = ev.kind == 1 ? ev.p + tile * size_t(ev.n_rows) * 32 + smp : ev.p;
  const int esh = ev.kind == 1 ? 5 : 0;
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  const int64_t gf_raw = kPacked ? int64_t(blockIdx.x) * kPack + lane / kGroup : int64_t(blockIdx.x);
  const bool f_ok = !kPacked || gf_raw < g.num_or;
  const int64_t gf = f_ok ? gf_raw : g.num_or - 1;
  const int p0 = g.parent_ptr[gf], p1 = g.parent_ptr[gf + 1];
  const int n = f_ok ? p1 - p0 : 0;
  // trip count of the parent loops: the factor's own count, or the longest of the CTA's factors
  const int n_top = kPacked ? int(__reduce_max_sync(0xffffffffu, unsigned(n))) : n;
  int4* wsm = reinterpret_cast<int4*>(fz_raw);                                             // [max_parents][2]
  float* stash = reinterpret_cast<float*>(fz_raw + size_t(g.max_parents) * sizeof(FusedW)) + lane;  // [parent][3][32]
  float* agg = stash + size_t(g.max_parents) * 96;                                         // [8][32]
  float dmax = 0.f;

  // the factor's wiring: one coalesced sweep (the per-parent round trip below is then data only);
  // packed lanes read their own factor's records from global memory (L2-resident, one address
  // per factor slot)
  const int4* wsrc = reinterpret_cast<const int4*>(g.w + p0);
  if (!kPacked) {
    for (int t = threadIdx.x; t < 2 * n; t += blockDim.x) wsm[t] = wsrc[t];
    __syncthreads();
  }
  auto wire = [&](int t) { return kPacked ? __ldg(wsrc + t) : wsm[t]; };

  // ---- phase A ---------------------------------------------------------------------------------
  // software-pipelined: the ten loads of the warp's NEXT parent are in flight while the current
  // one is computed (the kernel is bound by memory latency, not by bytes)
  struct Raw {
    FusedW w;
    float xO, xA, e0, e1, xs, xw, Ss0, Ss1, Sw0, Sw1;
  };
  auto fetch = [&](int j, Raw& r) {
    const int4 lo = wire(2 * j), hi = wire(2 * j + 1);
    r.w = FusedW{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    r.xO = mo[uint32_t(r.w.mO) << 5];
    r.xA = mo[uint32_t(r.w.mA) << 5];
    r.e0 = evq[uint32_t(r.w.ev0) << esh];
    r.e1 = evq[uint32_t(r.w.ev0 + 1) << esh];
    r.xs = mo[uint32_t(r.w.ms) << 5];
    r.xw = mo[uint32_t(r.w.mw) << 5];
    r.Ss0 = SL[uint32_t(r.w.Ss0) << 5];
    r.Ss1 = SL[uint32_t(r.w.Ss0 + 1) << 5];
    r.Sw0 = SL[uint32_t(r.w.Sw0) << 5];
    r.Sw1 = SL[uint32_t(r.w.Sw0 + 1) << 5];
  };
  Raw cur, nxt;
  if (warp < n) fetch(warp, cur);
  for (int j = warp; j < n_top; j += kFusedWarps) {
    if (j + kFusedWarps < n) fetch(j + kFusedWarps, nxt);
    if (j < n) {
      const Raw& r = cur;
      // SW_i as parent of the OR factor (off = +1: pointed state 0; its own edge has the smaller
      // message index: kind 3) -> (a_i, b_i) = (relevant, pointed) variable -> factor messages
      EdgeIn po;
      po.m_p = r.xO; po.m_r = 0.f; po.a_p = r.e0; po.a_r = r.e1; po.o_p = r.xA; po.o_r = 0.f;
      po.msg = r.w.mO << 1; po.kind = 3;
      float ob, oa;
      edge_q<true, kSumProduct>(po, 1, ob, oa);
      stash[size_t(j) * 96] = oa;
      stash[size_t(j) * 96 + 32] = ob;
      stash[size_t(j) * 96 + 64] = r.xO;
      // the AND factor (off = -1: pointed state 1): child SW_i (the OR edge comes first: kind 2),
      // parents with sums from S (kind 0)
      EdgeIn ce, ps, pw;
      ce.m_p = r.xA; ce.m_r = 0.f; ce.a_p = r.e1; ce.a_r = r.e0; ce.o_p = r.xO; ce.o_r = 0.f;
      ce.msg = (r.w.mA << 1) + 1; ce.kind = 2;
      ps.m_p = r.xs; ps.m_r = 0.f; ps.a_p = r.Ss1; ps.a_r = r.Ss0; ps.o_p = r.xs; ps.o_r = 0.f;
      ps.msg = (r.w.ms << 1) + 1; ps.kind = 0;
      pw.m_p = r.xw; pw.m_r = 0.f; pw.a_p = r.Sw1; pw.a_r = r.Sw0; pw.o_p = r.xw; pw.o_r = 0.f;
      pw.msg = (r.w.mw << 1) + 1; pw.kind = 0;
      float c_p, c_r, s_p, s_r, w_p, w_r;
      edge_q<true, kSumProduct>(ce, -1, c_p, c_r);
      edge_q<true, kSumProduct>(ps, -1, s_p, s_r);
      edge_q<true, kSumProduct>(pw, -1, w_p, w_r);
      const int64_t q0 = int64_t(r.w.f_and) * 2;
      LogicalAcc A;
      A.istar = q0;
      A.add<kSumProduct>(q0, s_r, s_p, T);
      A.add<kSumProduct>(q0 + 1, w_r, w_p, T);
      const float x_s = A.parent_out<kSumProduct>(q0, s_r, s_p, c_r, c_p, T, false);
      const float x_w = A.parent_out<kSumProduct>(q0 + 1, w_r, w_p, c_r, c_p, T, false);
      dmax = fmaxf(dmax, store_edge<kDelta, true, kSumProduct>(mn, -1, ps, x_s, d, one_minus_d));
      dmax = fmaxf(dmax, store_edge<kDelta, true, kSumProduct>(mn, -1, pw, x_w, d, one_minus_d));
      dmax = fmaxf(dmax, store_edge<kDelta, true, kSumProduct>(mn, -1, ce, A.child_relevant<kSumProduct>(T) - A.Sb, d,
                                                                one_minus_d));
    }
    cur = nxt;
  }
  // the OR factor's child edge: requested by warp 6 before the barrier, consumed after it
  EdgeIn child;
  if (warp == 6) child = load_edge<true, kSumProduct>(g.or_children[gf], 1, mo, evq, esh, SL);
  __syncthreads();
  // ---- phase B ---------------------------------------------------------------------------------
  // The two sums must be formed in ascending parent order (fp32 rounding): one warp each, a bare
  // chain of dependent additions.  The two largest differences are exact selections, so four warps
  // scan a quarter of the parents each and the quarters are merged after the barrier (first
  // arg-max = LARGEST tied index, second maximum counting duplicates - LogicalAcc::add's rule).
  if (warp == 0 || warp == 1) {
    float sum = 0.f;
    int j = 0;
    if (!kPacked) {
      for (; j + 8 <= n; j += 8) {
        float va[8], vb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { vb[k] = stash[size_t(j + k) * 96 + 32]; va[k] = warp == 0 ? 0.f : stash[size_t(j + k) * 96]; }
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += warp == 0 ? vb[k] : (kSumProduct ? logaddexp_t(va[k], vb[k], T) : fmaxf(vb[k], va[k]));
      }
    }
    for (; j < n_top; ++j) {
      const float vb = stash[size_t(j) * 96 + 32], va = stash[size_t(j) * 96];
      const float term = warp == 0 ? vb : (kSumProduct ? logaddexp_t(va, vb, T) : fmaxf(vb, va));
      if (j < n) sum += term;
    }
    agg[warp == 0 ? 32 : 0] = sum;  // Sb / acc
  } else if (warp >= 2 && warp < 6) {
    const int q = warp - 2, per = (n + 3) / 4;
    const int per_top = (n_top + 3) / 4;
    const int j0 = q * per, j1 = min(n, j0 + per);
    float d1 = -INFINITY, d2 = -INFINITY;
    int istar = p0;
    for (int t = 0; t < per_top; ++t) {
      const int j = j0 + t;
      if (j < j1) {
        const float dl = stash[size_t(j) * 96] - stash[size_t(j) * 96 + 32];
        if (dl >= d1) { d2 = d1; d1 = dl; istar = p0 + j; }
        else if (dl > d2) d2 = dl;
      }
    }
    float* seg = agg + (8 + 3 * q) * 32;  // rows 8.. of the aggregate area
    seg[0] = d1; seg[32] = d2; seg[64] = __int_as_float(istar);
  } else if (warp == 6) {
    float c_p, c_r;
    edge_q<true, kSumProduct>(child, 1, c_p, c_r);  // (expands child.m_p / m_r in place for store_edge below)
    agg[160] = c_p; agg[192] = c_r;
  }
  __syncthreads();
  // ---- phase C ---------------------------------------------------------------------------------
  {
    LogicalAcc A;
    A.acc = agg[0]; A.Sb = agg[32];
    A.d1 = -INFINITY; A.d2 = -INFINITY; A.istar = p0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // quarters in ascending parent order (empty ones skipped)
      if (q * ((n + 3) / 4) >= n) continue;
      const float* seg = agg + (8 + 3 * q) * 32;
      const float sd1 = seg[0], sd2 = seg[32];
      const int sidx = __float_as_int(seg[64]);
      if (sd1 >= A.d1) { A.d2 = fmaxf(A.d1, sd2); A.d1 = sd1; A.istar = sidx; }
      else A.d2 = fmaxf(A.d2, sd1);
    }
    const float c_p = agg[160], c_r = agg[192];
    const bool single = n == 1;
    if (warp == 6 && f_ok)
      dmax = fmaxf(dmax, store_edge<kDelta, true, kSumProduct>(mn, 1, child, A.child_relevant<kSumProduct>(T) - A.Sb, d,
                                                                one_minus_d));
    for (int j = warp; j < n_top; j += kFusedWarps) {
      if (j < n) {
        const float oa = stash[size_t(j) * 96], ob = stash[size_t(j) * 96 + 32];
        EdgeIn po;
        po.m_p = stash[size_t(j) * 96 + 64]; po.m_r = 0.f; po.msg = wire(2 * j).x << 1;
        expand_msg<true, kSumProduct>(1, po.m_p, po.m_r);
        const float x = A.parent_out<kSumProduct>(p0 + j, oa, ob, c_r, c_p, T, single);
        dmax = fmaxf(dmax, store_edge<kDelta, true, kSumProduct>(mn, 1, po, x, d, one_minus_d));
      }
    }
  }
  if (kDelta && b < batch) publish_delta(a.deltas, int64_t(b) * a.delta_stride + a.delta_off, dmax);
}

// ---------------------------------------------------------------------------
// K5: Pool update (pgmax/factor/pool.py:328-474).
// ---------------------------------------------------------------------------
template <bool kSumProduct>
__global__ void __launch_bounds__(kThreads)
k_pool(BatchMap mp, LogicalDev w, const float* __restrict__ S, const float* __restrict__ m_old,
       float* __restrict__ m_new, RunArgs a) {
  UnitLoop L = unit_loop(mp, w.num_factors);
  if (!L.b_ok) return;
  float dmax = 0.f;
  const int sh = mp.bx_log;
  const int64_t moff = lane_off(mp, a.Es, L.b);
  const float* mo = m_old + moff;
  float* mn = m_new + moff;
  const float* SL = S + lane_off(mp, a.Vs, L.b);
  const float T = a.T, d = a.d, one_minus_d = a.one_minus_d;
  // difference (state 1 - state 0) of the variable->factor message of a binary edge
  auto diff = [&](int64_t pm, int64_t pv) {
    return (SL[(pv + 1) << sh] - mo[(pm + 1) << sh]) - (SL[pv << sh] - mo[pm << sh]);
  };
  for (int64_t f = L.u; f < L.u_end; f += L.step) {
    const int64_t p0 = w.parent_ptr[f], p1 = w.parent_ptr[f + 1];
    const int64_t c = w.children_msg[f], cvs = w.children_vs[f];
    const float D = diff(c, cvs);
    float d1 = -INFINITY, d2 = -INFINITY;
    int64_t istar = p0;
    for (int64_t i = p0; i < p1; ++i) {
      const float dl = diff(w.parents_msg[i], w.parents_vs[i]);
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
        const float dl = diff(w.parents_msg[i], w.parents_vs[i]);
        sum += expf((dl - d1) / T);
        mx2 = fmaxf(mx2, (i == istar) ? -D : dl);
      }
      out_ind = T * logf(sum) + d1;
      G = logaddexp_t(out_ind, -D, T);
      float sum2 = 0.f;
      for (int64_t i = p0; i < p1; ++i) {
        const float dl = diff(w.parents_msg[i], w.parents_vs[i]);
        sum2 += expf((((i == istar) ? -D : dl) - mx2) / T);
      }
      out_star = -(T * logf(sum2) + mx2);
    }
    for (int64_t i = p0; i < p1; ++i) {
      const int64_t pm = w.parents_msg[i];
      float x;
      if (kSumProduct) {
        const float dl = diff(pm, w.parents_vs[i]);
        x = (i == istar) ? out_star : -logminusexp_t(G, dl, T, 1e-30f);
      } else {
        x = fminf(D, -((i == istar) ? d2 : d1));
      }
      if (single) x = D;  // pool.py:430-450
      dmax = fmaxf(dmax, write_binary_edge(mo, mn, pm, sh, 0.f, x, d, one_minus_d));
    }
    dmax = fmaxf(dmax, write_binary_edge(mo, mn, c, sh, 0.f, out_ind, d, one_minus_d));
  }
  publish_delta(a.deltas, int64_t(L.b) * a.delta_stride + a.delta_off, dmax);
}

}  // namespace pgx
