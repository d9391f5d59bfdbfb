// pgx_vjp.cuh — reverse mode through bp.run (kernels/vjp.cuh).  Included at the end of pgx.cu.
//
// Replaces what jax.grad gives the reference for free (pgmax/infer/bp.py:98 @jax.checkpoint on the
// update; examples/grid_mrf.ipynb cells 15-16): the cotangent of the final messages is pulled back
// to the log potentials, the evidence and the initial messages.  Sum-product (temperature > 0),
// graphs of EnumFactors with at most kSmallMaxNS edge-states per factor.  The iterations are re-run
// with the generic two-pass kernels (serial summation order) keeping every iterate - memory
// (num_iters + 1) x batch x E_s floats - and walked backwards; scratch is allocated per call.

namespace {

struct VjpScratch {
  std::vector<void*> ptrs;
  ~VjpScratch() {
    for (void* p : ptrs) cudaFree(p);
  }
  int alloc(float** p, size_t floats) {
    *p = nullptr;
    PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(floats, 1) * sizeof(float)));
    ptrs.push_back(*p);
    return PGX_OK;
  }
};

}  // namespace

extern "C" {

int pgx_bp_run_vjp(pgx_plan* plan, void* stream, int64_t batch, const float* log_potentials, int lp_batched,
                   const float* evidence, int ev_batched, const float* ftov_in, int msgs_batched,
                   const float* g_ftov_out, int32_t num_iters, float damping, float temperature, float* g_lp_out,
                   float* g_ev_out, float* g_ftov_in_out) {
  if (!plan) return fail(PGX_ERR_INVALID, "null plan");
  PGX_CHECK(batch >= 1 && batch < (1 << 24), "batch must be in [1, 2^24), got %lld", (long long)batch);
  PGX_CHECK(num_iters >= 1, "num_iters must be >= 1, got %d", num_iters);
  PGX_CHECK(g_ftov_out != nullptr, "g_ftov_out is null");
  if (!(temperature > 0.f))
    return fail(PGX_ERR_UNSUPPORTED, "the reverse pass is defined for sum-product (temperature > 0) only");
  if (plan->or_f.dev.num_factors + plan->and_f.dev.num_factors + plan->pool_f.dev.num_factors > 0)
    return fail(PGX_ERR_UNSUPPORTED, "the reverse pass covers EnumFactors only");
  for (const EnumBlockPlan& eb : plan->enum_blocks)
    if (eb.dev.ns > pgx::kSmallMaxNS)
      return fail(PGX_ERR_UNSUPPORTED, "the reverse pass covers factors of at most %d edge-states (got %d)",
                  pgx::kSmallMaxNS, eb.dev.ns);
  int rc;
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(plan))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const pgx::BatchMap mp = make_map(batch);
  const int64_t Es = plan->num_edge_states, Vs = plan->num_var_states, C = plan->num_potentials;
  if (Es == 0) return PGX_OK;
  const size_t tEs = tiled_floats(mp, Es), tVs = tiled_floats(mp, Vs), tC = tiled_floats(mp, C);
  PGX_CHECK(double(tEs) * (num_iters + 6) * 4 < 64e9, "the reverse pass would keep %.1f GB of iterates",
            double(tEs) * (num_iters + 1) * 4e-9);
  VjpScratch scratch;
  float *traj, *S, *m_raw, *G, *G2, *gq, *gm, *gS, *gev, *glp, *evT = nullptr, *lpT = nullptr, *zeros;
  if ((rc = scratch.alloc(&traj, tEs * size_t(num_iters + 1)))) return rc;
  if ((rc = scratch.alloc(&S, tVs)) || (rc = scratch.alloc(&m_raw, tEs)) || (rc = scratch.alloc(&G, tEs)) ||
      (rc = scratch.alloc(&G2, tEs)) || (rc = scratch.alloc(&gq, tEs)) || (rc = scratch.alloc(&gm, tEs)) ||
      (rc = scratch.alloc(&gS, tVs)) || (rc = scratch.alloc(&gev, tVs)) || (rc = scratch.alloc(&glp, tC)) ||
      (rc = scratch.alloc(&zeros, size_t(Vs))))
    return rc;
  PGX_CUDA(cudaMemsetAsync(S, 0, tVs * sizeof(float), st));
  PGX_CUDA(cudaMemsetAsync(gev, 0, tVs * sizeof(float), st));
  PGX_CUDA(cudaMemsetAsync(glp, 0, std::max<size_t>(tC, 1) * sizeof(float), st));
  PGX_CUDA(cudaMemsetAsync(zeros, 0, std::max<size_t>(size_t(Vs), 1) * sizeof(float), st));
  const bool single = batch == 1;
  pgx::View ev{evidence, Vs, 0}, lp{log_potentials, C, 0};
  if (!single && ev_batched) {
    if ((rc = scratch.alloc(&evT, tVs))) return rc;
    if ((rc = to_tiles(plan, st, evidence, evT, Vs, mp))) return rc;
    ev = pgx::View{evT, Vs, 1};
  }
  if (!single && lp_batched) {
    if ((rc = scratch.alloc(&lpT, tC))) return rc;
    if ((rc = to_tiles(plan, st, log_potentials, lpT, C, mp))) return rc;
    lp = pgx::View{lpT, C, 1};
  }
  // ---- forward, every iterate kept: traj[0] = NC(m_in), traj[t + 1] = update(traj[t]) ---------
  if (ftov_in == nullptr) {
    PGX_CUDA(cudaMemsetAsync(m_raw, 0, tEs * sizeof(float), st));
  } else if (single) {
    PGX_CUDA(cudaMemcpyAsync(m_raw, ftov_in, size_t(Es) * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else if (msgs_batched) {
    if ((rc = to_tiles(plan, st, ftov_in, m_raw, Es, mp))) return rc;
  } else {
    pgx::k_broadcast_rows<<<plan->num_sms * 8, pgx::kThreads, 0, st>>>(ftov_in, m_raw, Es, mp);
    if ((rc = check_launch(plan, "k_broadcast_rows"))) return rc;
  }
  PGX_CUDA(cudaMemcpyAsync(traj, m_raw, tEs * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if ((rc = normalize_edges(plan, st, mp, traj))) return rc;
  pgx::RunArgs a = make_run_args(damping, temperature, nullptr, num_iters, Es, Vs);
  for (int t = 0; t < num_iters; ++t) {
    const float* cur = traj + size_t(t) * tEs;
    pgx::k_var_sums<<<grid_for(plan, mp, Vs), pgx::kThreads, 0, st>>>(mp, Vs, Es, plan->d_vs_csr, plan->d_var_edge_msg, ev,
                                                                     cur, S, 0);
    if ((rc = check_launch(plan, "k_var_sums"))) return rc;
    if ((rc = launch_f2v<true>(plan, st, mp, lp, S, cur, traj + size_t(t + 1) * tEs, a, false, false, ev, nullptr)))
      return rc;
  }
  // ---- backward ----------------------------------------------------------------------------
  if (single) {
    PGX_CUDA(cudaMemcpyAsync(G, g_ftov_out, size_t(Es) * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else {
    if ((rc = to_tiles(plan, st, g_ftov_out, G, Es, mp))) return rc;
  }
  const pgx::View zero_ev{zeros, Vs, 0};
  for (int t = num_iters - 1; t >= 0; --t) {
    const float* cur = traj + size_t(t) * tEs;
    pgx::k_var_sums<<<grid_for(plan, mp, Vs), pgx::kThreads, 0, st>>>(mp, Vs, Es, plan->d_vs_csr, plan->d_var_edge_msg, ev,
                                                                     cur, S, 0);
    if ((rc = check_launch(plan, "k_var_sums"))) return rc;
    for (const EnumBlockPlan& eb : plan->enum_blocks) {
      pgx::k_enum_vjp<<<grid_for(plan, mp, eb.dev.num_factors), pgx::kThreads, 0, st>>>(
          mp, eb.dev, plan->d_edge_vs, lp, int(C), S, cur, G, gq, gm, glp, a);
      if ((rc = check_launch(plan, "k_enum_vjp"))) return rc;
    }
    pgx::k_var_sums<<<grid_for(plan, mp, Vs), pgx::kThreads, 0, st>>>(mp, Vs, Es, plan->d_vs_csr, plan->d_var_edge_msg,
                                                                     zero_ev, gq, gS, 0);
    if ((rc = check_launch(plan, "k_var_sums"))) return rc;
    pgx::k_vjp_accumulate<<<plan->num_sms * 4, pgx::kThreads, 0, st>>>(gev, gS, int64_t(tVs));
    if ((rc = check_launch(plan, "k_vjp_accumulate"))) return rc;
    pgx::k_vjp_combine<<<grid_for(plan, mp, plan->num_edges), pgx::kThreads, 0, st>>>(
        mp, plan->num_edges, Es, Vs, plan->d_edge_msg_start, plan->d_edge_vs, gm, gq, gS, G2);
    if ((rc = check_launch(plan, "k_vjp_combine"))) return rc;
    std::swap(G, G2);
  }
  pgx::k_vjp_normalize<<<grid_for(plan, mp, plan->num_edges), pgx::kThreads, 0, st>>>(mp, plan->num_edges, Es,
                                                                                      plan->d_edge_msg_start, m_raw, G);
  if ((rc = check_launch(plan, "k_vjp_normalize"))) return rc;
  // ---- results in the ABI layout: per sample for batched inputs, summed over the batch otherwise ----
  auto emit = [&](const float* tiles, int64_t N, bool batched, float* out) -> int {
    if (out == nullptr || N == 0) return PGX_OK;
    if (single) {
      PGX_CUDA(cudaMemcpyAsync(out, tiles, size_t(N) * sizeof(float), cudaMemcpyDeviceToDevice, st));
      return PGX_OK;
    }
    if (batched) return from_tiles(plan, st, tiles, out, N, mp);
    pgx::k_vjp_sum_batch<<<plan->num_sms * 4, pgx::kThreads, 0, st>>>(mp, N, tiles, out);
    return check_launch(plan, "k_vjp_sum_batch");
  };
  if ((rc = emit(glp, C, lp_batched != 0, g_lp_out))) return rc;
  if ((rc = emit(gev, Vs, ev_batched != 0, g_ev_out))) return rc;
  if ((rc = emit(G, Es, ftov_in != nullptr && msgs_batched != 0, g_ftov_in_out))) return rc;
  PGX_CUDA(cudaStreamSynchronize(st));  // the scratch is freed on return
  return PGX_OK;
}

}  // extern "C"
