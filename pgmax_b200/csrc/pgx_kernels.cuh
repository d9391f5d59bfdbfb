// pgx_kernels.cuh — sm_100a kernels of the loopy-BP hot path.
//
// Internal data layout ("tile-blocked batch-inner"): every per-sample vector
// x[b][n] of the ABI (batch-major, as jax.vmap produces) is held as
// X[tile][n][TW]: samples are grouped in tiles of TW = min(32, pow2ceil(B))
// consecutive samples; inside a tile the sample index is fastest and element n
// of the tile starts at float offset (tile * N + n) * TW.  A warp covers the TW
// samples of 32 / TW graph elements, so
//   * every index load (incidence, wiring, config tables) is warp-uniform per
//     element, every message / evidence / var-sum access is a full coalesced
//     row, whatever the graph's structure;
//   * consecutive elements of one tile are CONTIGUOUS in memory (128 B apart
//     for TW = 32): a factor's edge-states, and a run of factors, form one
//     contiguous span that a single TMA bulk copy can move, and all address
//     arithmetic is "lane pointer + (n << log2 TW)".
// For batch == 1 this degenerates to the reference's flat vectors, one graph
// element per lane.
//
// Arithmetic follows SURVEY.md App. A operation by operation (same order of
// additions, expf/logf/log1pf/expm1f, no FMA contraction: the library is compiled
// with -fmad=false) so that max-product results are bit-comparable with the CPU
// oracle.  The one deliberate exception is the sum-product path of the
// pairwise-binary kernels (see lse2 below).
#pragma once

#include "kernels/common.cuh"
#include "kernels/var_sums.cuh"
#include "kernels/pairwise.cuh"
#include "kernels/lattice.cuh"
#include "kernels/dense_grid.cuh"
#include "kernels/lattice_bin.cuh"
#include "kernels/enum.cuh"
#include "kernels/logical.cuh"
#include "kernels/decode_energy.cuh"
#include "kernels/sdlp.cuh"
#include "kernels/vjp.cuh"
