// pgx_sdlp.cuh — launch sequencing of the smooth dual LP-MAP entry points (pgx_sdlp_* in
// include/pgx.h; kernels in kernels/sdlp.cuh).  Included at the end of pgx.cu: same translation
// unit, same helpers, no message arithmetic on the host.

namespace {

void free_sdlp(SdlpWorkspace& w) {
  free_dev(w.eta); free_dev(w.P); free_dev(w.vval); free_dev(w.eval); free_dev(w.grad); free_dev(w.partial);
  w = SdlpWorkspace{};
}

// Slots of the objective's first reduction stage: the concurrent (element) lanes of the launch.
dim3 sdlp_objval_grid(const pgx_plan* plan, const pgx::BatchMap& mp, int64_t* num_slots) {
  const dim3 grid = grid_for(plan, mp, plan->num_vars + plan->num_factors);
  *num_slots = int64_t(grid.x) * (pgx::kThreads / 32) * (32 >> mp.bx_log);
  return grid;
}

int ensure_sdlp(pgx_plan* plan, int64_t batch, bool need_eta, bool need_grad) {
  SdlpWorkspace& w = plan->sdlp;
  const pgx::BatchMap mp = make_map(batch);
  if (w.batch != batch) {
    free_sdlp(w);  // leaves w.batch == 0: a failed allocation below is retried by the next call
    int64_t slots = 0;
    sdlp_objval_grid(plan, mp, &slots);
    const size_t padded = size_t(mp.nbt) << mp.bx_log;
    PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(&w.P), tiled_floats(mp, plan->num_var_states) * sizeof(float)));
    PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(&w.vval), tiled_floats(mp, plan->num_vars) * sizeof(float)));
    PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(&w.eval), tiled_floats(mp, plan->num_edges) * sizeof(float)));
    PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(&w.partial), size_t(slots) * padded * sizeof(double)));
    w.batch = batch;
  }
  if (need_eta && w.eta == nullptr)
    PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(&w.eta), tiled_floats(mp, plan->num_edge_states) * sizeof(float)));
  if (need_grad && w.grad == nullptr)
    PGX_CUDA(cudaMalloc(reinterpret_cast<void**>(&w.grad), tiled_floats(mp, plan->num_edge_states) * sizeof(float)));
  return PGX_OK;
}

// BP updates on vtof = -m, un-normalised, unclipped potentials: the generic kernels only
// (the specialised BP paths bake in damping + normalisation).
template <bool kSum>
int launch_f2v_raw(pgx_plan* plan, cudaStream_t st, const pgx::BatchMap& mp, pgx::View lp, const float* m, float* upd,
                   const pgx::RunArgs& a) {
  int rc;
  for (EnumBlockPlan& eb : plan->enum_blocks) {
    const int64_t F = eb.dev.num_factors;
    if (eb.dev.ns <= pgx::kSmallMaxNS) {
      pgx::k_enum_small<kSum, true><<<grid_for(plan, mp, F), pgx::kThreads, 0, st>>>(mp, eb.dev, plan->d_edge_vs, lp,
                                                                                   nullptr, m, upd, a);
      if ((rc = check_launch(plan, "k_enum_small<raw>"))) return rc;
    } else {
      const size_t smem = size_t(2 * eb.dev.ns + 32) * sizeof(float);
      const int grid = int(std::min<int64_t>(F * mp.batch, int64_t(plan->num_sms) * 8));
      const int group = kSum ? kAttrSdlpSum : kAttrSdlpMax;
      if (!((plan->attr_done >> group) & 1u)) {
        plan->attr_done |= 1u << group;
        PGX_CUDA(cudaFuncSetAttribute(pgx::k_enum_big<kSum, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      }
      pgx::k_enum_big<kSum, true><<<grid, pgx::kThreads, smem, st>>>(mp, eb.dev, plan->d_edge_vs, lp, nullptr, m, upd, a);
      if ((rc = check_launch(plan, "k_enum_big<raw>"))) return rc;
    }
  }
  for (LogicalPlan* lg : {&plan->or_f, &plan->and_f}) {
    if (lg->dev.num_factors == 0) continue;
    pgx::k_logical_raw<kSum><<<grid_for(plan, mp, lg->dev.num_factors), pgx::kThreads, 0, st>>>(mp, lg->dev, m, upd, a);
    if ((rc = check_launch(plan, "k_logical_raw"))) return rc;
  }
  if (plan->pool_f.dev.num_factors > 0) {
    pgx::k_pool_raw<kSum><<<grid_for(plan, mp, plan->pool_f.dev.num_factors), pgx::kThreads, 0, st>>>(
        mp, plan->pool_f.dev, m, upd, a);
    if ((rc = check_launch(plan, "k_pool_raw"))) return rc;
  }
  return PGX_OK;
}

// One evaluation of the smooth dual objective and its gradient at the messages in ws.mA
// (tile-blocked), optionally followed by the step on (ws.mA, sdlp.eta).
int sdlp_eval(pgx_plan* plan, cudaStream_t st, const pgx::BatchMap& mp, pgx::View lp, pgx::View ev, float T,
              bool store_grad, bool do_step, float step, float momentum, float* objvals, int64_t obj_stride,
              int64_t obj_off) {
  int rc;
  Workspace& ws = plan->ws;
  SdlpWorkspace& w = plan->sdlp;
  const int64_t Es = plan->num_edge_states, Vs = plan->num_var_states;
  pgx::RunArgs a{};
  a.T = T;
  a.Es = Es;
  a.Vs = Vs;
  if (Vs > 0) {
    pgx::k_var_sums<<<grid_for(plan, mp, Vs), pgx::kThreads, 0, st>>>(mp, Vs, Es, plan->d_vs_csr, plan->d_var_edge_msg, ev,
                                                                     ws.mA, ws.S, 0);
    if ((rc = check_launch(plan, "k_var_sums"))) return rc;
  }
  if (plan->num_vars > 0) {
    if (T > 0.f)
      pgx::k_sdlp_vars<true><<<grid_for(plan, mp, plan->num_vars), pgx::kThreads, 0, st>>>(
          mp, plan->num_vars, Vs, plan->d_var_first_state, ws.S, w.P, w.vval, T);
    else
      pgx::k_sdlp_vars<false><<<grid_for(plan, mp, plan->num_vars), pgx::kThreads, 0, st>>>(
          mp, plan->num_vars, Vs, plan->d_var_first_state, ws.S, w.P, w.vval, T);
    if ((rc = check_launch(plan, "k_sdlp_vars"))) return rc;
  }
  if (Es > 0) {
    rc = T > 0.f ? launch_f2v_raw<true>(plan, st, mp, lp, ws.mA, ws.mB, a)
                 : launch_f2v_raw<false>(plan, st, mp, lp, ws.mA, ws.mB, a);
    if (rc) return rc;
    float* grad = store_grad ? w.grad : nullptr;
    if (T > 0.f)
      pgx::k_sdlp_edges<true><<<grid_for(plan, mp, plan->num_edges), pgx::kThreads, 0, st>>>(
          mp, plan->num_edges, Es, Vs, plan->d_edge_msg_start, plan->d_edge_vs, ws.mB, w.P, ws.mA, w.eta, grad, w.eval, T,
          do_step ? 1 : 0, step, momentum);
    else
      pgx::k_sdlp_edges<false><<<grid_for(plan, mp, plan->num_edges), pgx::kThreads, 0, st>>>(
          mp, plan->num_edges, Es, Vs, plan->d_edge_msg_start, plan->d_edge_vs, ws.mB, w.P, ws.mA, w.eta, grad, w.eval, T,
          do_step ? 1 : 0, step, momentum);
    if ((rc = check_launch(plan, "k_sdlp_edges"))) return rc;
  }
  if (objvals != nullptr) {
    int64_t slots = 0;
    const dim3 grid = sdlp_objval_grid(plan, mp, &slots);
    const int64_t padded = int64_t(mp.nbt) << mp.bx_log;
    pgx::k_sdlp_objval_partial<<<grid, pgx::kThreads, 0, st>>>(mp, plan->num_vars, plan->num_edges, plan->num_factors,
                                                             plan->d_factor_edge_start, w.vval, w.eval, w.partial, padded);
    if ((rc = check_launch(plan, "k_sdlp_objval_partial"))) return rc;
    pgx::k_sdlp_objval_final<<<unsigned(mp.batch), pgx::kThreads, 0, st>>>(w.partial, slots, padded, objvals, obj_stride,
                                                                         obj_off);
    if ((rc = check_launch(plan, "k_sdlp_objval_final"))) return rc;
  }
  return PGX_OK;
}

// Inputs -> workspace: evidence / potentials views, messages into ws.mA (tile-blocked).
int sdlp_stage(pgx_plan* plan, cudaStream_t st, int64_t batch, const float* log_potentials, int lp_batched,
               const float* evidence, int ev_batched, const float* ftov, int msgs_batched, bool need_eta, bool need_grad,
               pgx::View* lp, pgx::View* ev) {
  PGX_CHECK(batch >= 1 && batch < (1 << 24), "batch must be in [1, 2^24), got %lld", (long long)batch);
  PGX_CHECK(plan->d_factor_edge_start != nullptr || plan->num_edges == 0,
            "pgx_plan_set_factors must be called before pgx_sdlp_*");
  PGX_CHECK(plan->num_var_states == 0 || evidence != nullptr, "evidence is null");
  PGX_CHECK(plan->num_potentials == 0 || log_potentials != nullptr, "log_potentials is null");
  int rc;
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(plan))) return rc;
  const pgx::BatchMap mp = make_map(batch);
  const bool single = batch == 1;
  const bool evT = !single && ev_batched, lpT = !single && lp_batched;
  if ((rc = ensure_workspace(plan, batch, evT, lpT, false))) return rc;
  if ((rc = ensure_sdlp(plan, batch, need_eta, need_grad))) return rc;
  Workspace& ws = plan->ws;
  const int64_t Es = plan->num_edge_states, Vs = plan->num_var_states, C = plan->num_potentials;
  *ev = pgx::View{evidence, Vs, 0};
  *lp = pgx::View{log_potentials, C, 0};
  if (evT) {
    if ((rc = to_tiles(plan, st, evidence, ws.evT, Vs, mp))) return rc;
    *ev = pgx::View{ws.evT, Vs, 1};
  }
  if (lpT) {
    if ((rc = to_tiles(plan, st, log_potentials, ws.lpT, C, mp))) return rc;
    *lp = pgx::View{ws.lpT, C, 1};
  }
  if (Es == 0) return PGX_OK;
  if (ftov == nullptr) {
    PGX_CUDA(cudaMemsetAsync(ws.mA, 0, tiled_floats(mp, Es) * sizeof(float), st));
  } else if (single) {
    PGX_CUDA(cudaMemcpyAsync(ws.mA, ftov, size_t(Es) * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else if (msgs_batched) {
    if ((rc = to_tiles(plan, st, ftov, ws.mA, Es, mp))) return rc;
  } else {
    pgx::k_broadcast_rows<<<plan->num_sms * 8, pgx::kThreads, 0, st>>>(ftov, ws.mA, Es, mp);
    if ((rc = check_launch(plan, "k_broadcast_rows"))) return rc;
  }
  return PGX_OK;
}

int sdlp_unstage(pgx_plan* plan, cudaStream_t st, const pgx::BatchMap& mp, const float* src, float* dst, int64_t rows) {
  if (dst == nullptr || rows == 0) return PGX_OK;
  if (mp.batch == 1) {
    PGX_CUDA(cudaMemcpyAsync(dst, src, size_t(rows) * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return PGX_OK;
  }
  return from_tiles(plan, st, src, dst, rows, mp);
}

}  // namespace

extern "C" {

int pgx_plan_set_factors(pgx_plan* plan, int64_t num_factors, const int32_t* factor_edge_start) {
  if (!plan) return fail(PGX_ERR_INVALID, "null plan");
  PGX_CHECK(num_factors >= 0 && factor_edge_start != nullptr, "bad factor table");
  PGX_CHECK(factor_edge_start[0] == 0 && factor_edge_start[num_factors] == plan->num_edges,
            "factor_edge_start must run from 0 to num_edges (%lld), got %d .. %d", (long long)plan->num_edges,
            factor_edge_start[0], factor_edge_start[num_factors]);
  for (int64_t f = 0; f < num_factors; ++f)
    PGX_CHECK(factor_edge_start[f] <= factor_edge_start[f + 1], "factor_edge_start must be non-decreasing (factor %lld)",
              (long long)f);
  int rc;
  DeviceGuard device_guard;
  if ((rc = device_guard.enter(plan))) return rc;
  free_dev(plan->d_factor_edge_start);
  plan->d_factor_edge_start = nullptr;
  free_sdlp(plan->sdlp);
  std::vector<int32_t> host(factor_edge_start, factor_edge_start + num_factors + 1);
  if ((rc = upload(host, &plan->d_factor_edge_start, &plan->device_bytes))) return rc;
  plan->num_factors = num_factors;
  return PGX_OK;
}

int pgx_sdlp_objval_and_grad(pgx_plan* plan, void* stream, int64_t batch, const float* log_potentials, int lp_batched,
                             const float* evidence, int ev_batched, const float* ftov_msgs, int msgs_batched,
                             float logsumexp_temp, float* objval_out, float* grad_out, float* bp_updates_out,
                             float* edge_vals_out) {
  if (!plan) return fail(PGX_ERR_INVALID, "null plan");
  PGX_CHECK(logsumexp_temp >= 0.f, "logsumexp_temp must be >= 0");
  PGX_CHECK(objval_out != nullptr, "objval_out is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pgx::View lp{}, ev{};
  int rc;
  if ((rc = sdlp_stage(plan, st, batch, log_potentials, lp_batched, evidence, ev_batched, ftov_msgs, msgs_batched, false,
                       grad_out != nullptr, &lp, &ev)))
    return rc;
  const pgx::BatchMap mp = make_map(batch);
  if ((rc = sdlp_eval(plan, st, mp, lp, ev, logsumexp_temp, grad_out != nullptr, false, 0.f, 0.f, objval_out, 1, 0)))
    return rc;
  if ((rc = sdlp_unstage(plan, st, mp, plan->sdlp.grad, grad_out, plan->num_edge_states))) return rc;
  if ((rc = sdlp_unstage(plan, st, mp, plan->ws.mB, bp_updates_out, plan->num_edge_states))) return rc;
  return sdlp_unstage(plan, st, mp, plan->sdlp.eval, edge_vals_out, plan->num_edges);
}

int pgx_sdlp_run(pgx_plan* plan, void* stream, int64_t batch, const float* log_potentials, int lp_batched,
                 const float* evidence, int ev_batched, const float* ftov_in, int msgs_batched, float* ftov_out,
                 float* objvals, int32_t num_iters, const float* steps, const float* momenta, float logsumexp_temp) {
  if (!plan) return fail(PGX_ERR_INVALID, "null plan");
  PGX_CHECK(num_iters >= 0, "num_iters must be >= 0, got %d", num_iters);
  PGX_CHECK(logsumexp_temp >= 0.f && logsumexp_temp <= 1.f, "logsumexp_temp must be in [0, 1]");
  PGX_CHECK(num_iters == 0 || (steps != nullptr && momenta != nullptr), "steps / momenta are null");
  PGX_CHECK(ftov_out != nullptr || plan->num_edge_states == 0, "ftov_out is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pgx::View lp{}, ev{};
  int rc;
  if ((rc = sdlp_stage(plan, st, batch, log_potentials, lp_batched, evidence, ev_batched, ftov_in, msgs_batched, true,
                       false, &lp, &ev)))
    return rc;
  const pgx::BatchMap mp = make_map(batch);
  const int64_t Es = plan->num_edge_states;
  if (Es > 0)  // eta starts as a copy of the messages (dual_lp.py:289)
    PGX_CUDA(cudaMemcpyAsync(plan->sdlp.eta, plan->ws.mA, tiled_floats(mp, Es) * sizeof(float), cudaMemcpyDeviceToDevice, st));
  for (int32_t it = 0; it < num_iters; ++it)
    if ((rc = sdlp_eval(plan, st, mp, lp, ev, logsumexp_temp, false, true, steps[it], momenta[it], objvals, num_iters, it)))
      return rc;
  return sdlp_unstage(plan, st, mp, plan->ws.mA, ftov_out, Es);
}

}  // extern "C"
