"""Gradients through bp.run (pgx_bp_run_vjp + pgmax_b200/infer/grad.py).

The reference differentiates through run_bp with jax.grad (pgmax/infer/bp.py:98;
examples/grid_mrf.ipynb cells 15-16: value_and_grad of a cross-entropy loss of the marginals with
respect to the log potentials of four PairwiseFactorGroups that each share one matrix).  JAX is
not available, so the known answer is a central finite difference of the SAME loss evaluated with
the CPU oracle in float64 (oracle.bp_oracle.precision(np.float64))."""

import numpy as np
import pytest
import torch

from oracle import bp_oracle
from pgmax_b200 import fgraph, fgroup, infer, vgroup
from pgmax_b200.infer import grad

pytestmark = pytest.mark.gpu


def grid_mrf(M=4, N=5, num_states=4):
  """examples/grid_mrf.ipynb cell 12: top-down, left-right and the two diagonal factor groups."""
  variables = vgroup.NDVarArray(num_states=num_states, shape=(M, N))
  fg = fgraph.FactorGraph(variables)
  pairs = [
      [[variables[i, j], variables[i + 1, j]] for i in range(M - 1) for j in range(N)],
      [[variables[i, j], variables[i, j + 1]] for i in range(M) for j in range(N - 1)],
      [[variables[i, j], variables[i + 1, j + 1]] for i in range(M - 1) for j in range(N - 1)],
      [[variables[i, j], variables[i - 1, j + 1]] for i in range(1, M) for j in range(N - 1)],
  ]
  groups = [fgroup.PairwiseFactorGroup(variables_for_factors=p) for p in pairs]
  fg.add_factors(groups)
  return fg, variables, groups


def _flat_potentials(groups, matrices):
  """The flat [C] potential vector of matrices shared per group (what fgroup.flatten does), as a
  differentiable function of the matrices."""
  return torch.cat([m.reshape(-1).repeat(len(g.variables_for_factors)) for g, m in zip(groups, matrices)])


def _oracle_loss(graph, lp_flat, evidence, targets, iters, damping, ns):
  """The notebook's loss (cell 15) on the oracle, for a batch: mean over images of
  -mean_pixels log sum(target * marginals)."""
  total = 0.0
  for ev, tgt in zip(evidence, targets):
    msgs, _ = bp_oracle.run_bp(graph, lp_flat, np.zeros(graph.var_states_for_edge_states.shape[0]), ev, iters, damping, 1.0)
    beliefs = bp_oracle.flat_beliefs(graph, msgs, ev).reshape(-1, ns).astype(np.float64)
    marg = np.exp(beliefs - beliefs.max(-1, keepdims=True))
    marg /= marg.sum(-1, keepdims=True)
    total += -np.mean(np.log(np.sum(tgt.reshape(-1, ns) * marg, axis=-1)))
  return total / len(evidence)


@pytest.mark.parametrize("damping", [0.0, 0.5])
def test_value_and_grad_of_the_grid_mrf_loss(damping):
  M, N, ns, iters, batch = 4, 5, 4, 15, 3
  fg, variables, groups = grid_mrf(M, N, ns)
  bp = infer.build_inferer(fg.bp_state, backend="bp")
  rng = np.random.default_rng(0)
  mats = [0.3 * rng.normal(size=(ns, ns)) for _ in groups]
  evidence = rng.normal(size=(batch, M * N * ns)).astype(np.float32)
  labels = rng.integers(0, ns, size=(batch, M * N))
  targets = np.eye(ns, dtype=np.float32)[labels].reshape(batch, -1)
  dev = torch.device("cuda")
  t_mats = [torch.tensor(m, dtype=torch.float32, device=dev, requires_grad=True) for m in mats]
  t_ev = torch.tensor(evidence, device=dev, requires_grad=True)
  t_msgs = torch.zeros(bp.context.plan.num_edge_states, device=dev)
  msgs = grad.run(bp, _flat_potentials(groups, t_mats), t_ev, t_msgs, iters, damping, 1.0)
  marg = grad.marginals(bp, grad.flat_beliefs(bp, t_ev, msgs))[variables]          # [batch, M, N, ns]
  tgt = torch.tensor(targets, device=dev).reshape(batch, M, N, ns)
  loss = -torch.log((tgt * marg).sum(-1)).mean(dim=(1, 2)).mean()
  loss.backward()
  graph = bp_oracle.graph_from_context(bp.context)
  flat = lambda ms: np.concatenate([np.tile(m.reshape(-1), len(g.variables_for_factors)) for g, m in zip(groups, ms)])
  with bp_oracle.precision(np.float64):
    want = _oracle_loss(graph, flat(mats), evidence.astype(np.float64), targets, iters, damping, ns)
    assert abs(float(loss.detach()) - want) < 1e-5
    h = 1e-4
    checked = 0
    for g_idx in range(len(groups)):
      for (a, b) in ((0, 0), (1, 2), (3, 1)):
        vals = []
        for sign in (1.0, -1.0):
          moved = [m.copy() for m in mats]
          moved[g_idx][a, b] += sign * h
          vals.append(_oracle_loss(graph, flat(moved), evidence.astype(np.float64), targets, iters, damping, ns))
        fd = (vals[0] - vals[1]) / (2 * h)
        got = float(t_mats[g_idx].grad[a, b])
        assert abs(got - fd) < 2e-4 + 2e-3 * abs(fd), (g_idx, a, b, got, fd)
        checked += 1
    # the evidence gradient (per sample), a few entries
    for (s, idx) in ((0, 3), (1, 17), (2, 40)):
      vals = []
      for sign in (1.0, -1.0):
        moved = evidence.astype(np.float64).copy()
        moved[s, idx] += sign * h
        vals.append(_oracle_loss(graph, flat(mats), moved, targets, iters, damping, ns))
      fd = (vals[0] - vals[1]) / (2 * h)
      assert abs(float(t_ev.grad[s, idx]) - fd) < 2e-4 + 2e-3 * abs(fd), (s, idx, float(t_ev.grad[s, idx]), fd)
  assert checked == 12


def test_vjp_of_initial_messages_and_unbatched_run():
  """One sample, un-normalised initial messages, unary + pairwise factors of unequal state counts:
  d<w, msgs_out>/d(msgs_in, evidence, potentials) against float64 finite differences."""
  num_states = np.array([2, 3, 3, 4])
  variables = vgroup.NDVarArray(num_states=num_states, shape=(4,))
  fg = fgraph.FactorGraph(variables)
  rng = np.random.default_rng(1)
  for a, b in ((0, 1), (1, 2), (2, 3), (3, 0), (1, 3)):
    fg.add_factors(fgroup.PairwiseFactorGroup(
        variables_for_factors=[[variables[a], variables[b]]],
        log_potential_matrix=0.5 * rng.normal(size=(num_states[a], num_states[b]))))
  bp = infer.build_inferer(fg.bp_state, backend="bp")
  arrays = bp.init()
  graph = bp_oracle.graph_from_context(bp.context)
  lp = np.asarray(arrays.log_potentials, np.float64)
  ev = rng.normal(size=arrays.evidence.shape)
  m0 = rng.normal(size=arrays.ftov_msgs.shape)
  w = rng.normal(size=arrays.ftov_msgs.shape)
  iters, damping, T = 6, 0.3, 0.7
  dev = torch.device("cuda")
  tens = [torch.tensor(x, dtype=torch.float32, device=dev, requires_grad=True) for x in (lp, ev, m0)]
  out = grad.run(bp, *tens, iters, damping, T)
  (out * torch.tensor(w, dtype=torch.float32, device=dev)).sum().backward()

  def value(lp_, ev_, m0_):
    with bp_oracle.precision(np.float64):
      msgs, _ = bp_oracle.run_bp(graph, lp_, m0_, ev_, iters, damping, T)
    return float(np.dot(w, msgs))

  h = 1e-5
  args = [lp, ev, m0]
  for which, name in enumerate(("log_potentials", "evidence", "ftov_msgs")):
    for idx in rng.choice(args[which].shape[0], size=5, replace=False):
      vals = []
      for sign in (1.0, -1.0):
        moved = [x.copy() for x in args]
        moved[which][idx] += sign * h
        vals.append(value(*moved))
      fd = (vals[0] - vals[1]) / (2 * h)
      got = float(tens[which].grad[idx])
      assert abs(got - fd) < 5e-4 + 5e-3 * abs(fd), (name, int(idx), got, fd)


def test_vjp_rejects_what_it_does_not_cover():
  from pgmax_b200 import _native
  import models
  fg, variables, evidence = models.ising_model(n=4)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, device="cuda", requires_grad=True)
  arrays = bp.init(evidence_updates={variables: evidence})
  with pytest.raises(ValueError, match="sum-product"):
    grad.run(bp, t(arrays.log_potentials), t(arrays.evidence), t(arrays.ftov_msgs), 3, 0.5, 0.0)
  data = models.logical_pair("or", 0)
  bp = infer.BP(data["graphs"][0][0].bp_state, temperature=1.0)
  arrays = models.init_logical(bp, data["graphs"][0], data)
  out = grad.run(bp, t(arrays.log_potentials), t(arrays.evidence), t(arrays.ftov_msgs), 2, 0.5, 1.0)
  with pytest.raises(_native.PgxError, match="EnumFactors only"):
    out.sum().backward()
