"""Differentiating through belief propagation (torch.autograd over the C ABI).

The reference gets gradients of anything computed from ``bp.run`` with ``jax.grad``
(pgmax/infer/bp.py:98 wraps the update in ``jax.checkpoint``; examples/grid_mrf.ipynb cells 15-16
train the log potentials of a grid MRF with ``jax.value_and_grad(batch_loss, argnums=2)``).  Here
the run is an opaque library call, so its vector-Jacobian product is a library call too
(``pgx_bp_run_vjp``), tied into torch's autograd:

  msgs = grad.run(bp, log_potentials, evidence, ftov_msgs, num_iters, damping, temperature)
  beliefs = grad.flat_beliefs(bp, evidence, msgs)          # differentiable
  marginals = grad.marginals(bp, beliefs)[variables]       # differentiable (uniform VarGroups)
  loss(marginals).backward()                               # -> .grad of log_potentials / evidence / messages

All arguments are CUDA float32 tensors in the flat layout of BPArrays ([C] / [V_s] / [E_s], or with
a leading batch axis).  Sum-product only (temperature > 0), graphs of EnumFactors of at most 64
edge-states per factor (include/pgx.h, pgx_bp_run_vjp).
"""

from typing import Any, Dict, Hashable

import numpy as np


def _fn():
  import torch  # pylint: disable=g-import-not-at-top

  class BPRun(torch.autograd.Function):
    """ftov_out = run(lp, ev, msgs); backward = pgx_bp_run_vjp."""

    @staticmethod
    def forward(ctx, lp, ev, msgs, plan, num_iters, damping, temperature):
      lp, ev, msgs = lp.contiguous(), ev.contiguous(), msgs.contiguous()
      sizes = {int(t.shape[0]) for t in (lp, ev, msgs) if t.ndim == 2}
      if len(sizes) > 1:
        raise ValueError(f"Inconsistent batch sizes: {sorted(sizes)}")
      batch = sizes.pop() if sizes else 1
      out = torch.empty((batch, plan.num_edge_states), dtype=torch.float32, device=ev.device)
      stream = torch.cuda.current_stream(ev.device).cuda_stream
      plan.bp_run(stream, batch, lp.data_ptr(), lp.ndim == 2, ev.data_ptr(), ev.ndim == 2, msgs.data_ptr(),
                  msgs.ndim == 2, out.data_ptr(), None, num_iters, damping, temperature)
      ctx.save_for_backward(lp, ev, msgs)
      ctx.plan, ctx.args, ctx.batch, ctx.squeeze = plan, (num_iters, damping, temperature), batch, not sizes and batch == 1
      return out[0] if ctx.squeeze else out

    @staticmethod
    def backward(ctx, g_out):
      lp, ev, msgs = ctx.saved_tensors
      plan, (num_iters, damping, temperature) = ctx.plan, ctx.args
      g_out = g_out.contiguous().to(torch.float32)
      g_lp, g_ev, g_msgs = torch.zeros_like(lp), torch.zeros_like(ev), torch.zeros_like(msgs)
      stream = torch.cuda.current_stream(ev.device).cuda_stream
      plan.bp_run_vjp(stream, ctx.batch, lp.data_ptr(), lp.ndim == 2, ev.data_ptr(), ev.ndim == 2, msgs.data_ptr(),
                      msgs.ndim == 2, g_out.data_ptr(), num_iters, damping, temperature, g_lp.data_ptr(),
                      g_ev.data_ptr(), g_msgs.data_ptr())
      return g_lp, g_ev, g_msgs, None, None, None, None

  return BPRun


_BPRUN = None


def run(bp, log_potentials, evidence, ftov_msgs, num_iters: int, damping: float = 0.5, temperature: float = 1.0):
  """Differentiable ``bp.run``: the messages after ``num_iters`` iterations as a tensor that
  carries gradients back to the three inputs."""
  global _BPRUN
  if _BPRUN is None:
    _BPRUN = _fn()
  if not temperature > 0.0:
    raise ValueError("Gradients through BP are defined for sum-product (temperature > 0)")
  return _BPRUN.apply(log_potentials, evidence, ftov_msgs, bp.context.plan, max(int(num_iters), 1), float(damping),
                      float(temperature))


def flat_beliefs(bp, evidence, ftov_msgs):
  """evidence + incoming messages per var-state (pgmax/infer/inferer.py:218-222) with torch ops."""
  import torch  # pylint: disable=g-import-not-at-top

  context = bp.context
  index = getattr(context, "_vs_index_torch", None)
  if index is None or index.device != ftov_msgs.device:
    index = torch.from_numpy(np.asarray(context.var_states_for_edge_states, dtype=np.int64)).to(ftov_msgs.device)
    context._vs_index_torch = index  # pylint: disable=protected-access
  if ftov_msgs.ndim == 2 and evidence.ndim == 1:
    evidence = evidence.expand(ftov_msgs.shape[0], -1)
  return evidence.index_add(-1, index, ftov_msgs if ftov_msgs.ndim == evidence.ndim else ftov_msgs.expand_as(evidence))


def marginals(bp, beliefs) -> Dict[Hashable, Any]:
  """Softmax of the beliefs per variable (pgmax/infer/bp.py:263-288) as a dict VarGroup -> tensor
  of shape [batch,] group shape + (num_states,); VarGroups with a uniform number of states."""
  import torch  # pylint: disable=g-import-not-at-top

  out, start = {}, 0
  for vg in bp.context.bp_state.fg_state.variable_groups:
    ns = np.asarray(vg.num_states).reshape(-1)
    length = int(ns.sum())
    if ns.size and not np.all(ns == ns[0]):
      raise ValueError("grad.marginals needs VarGroups with a uniform number of states")
    if ns.size:
      block = beliefs[..., start : start + length].reshape(beliefs.shape[:-1] + tuple(np.asarray(vg.num_states).shape) + (int(ns[0]),))
      out[vg] = torch.softmax(block, dim=-1)
    start += length
  return out
