"""Belief propagation front end (host mirror of pgmax/infer/bp.py:36-288).

``BP(bp_state, temperature)`` returns the same bundle of functions as the
reference; ``run`` / ``run_with_diffs`` enqueue the sm_100a kernels through the
C ABI (pgx_bp_run) instead of tracing a lax.scan.
"""

import dataclasses
import warnings
from typing import Any, Callable, Dict, Hashable, Optional, Tuple

import numpy as np

from pgmax_b200.infer.bp_state import BPArrays
from pgmax_b200.infer.bp_state import BPState
from pgmax_b200.infer.bp_state import _is_torch
from pgmax_b200.infer.inferer import DeviceBuffers
from pgmax_b200.infer.inferer import Inferer
from pgmax_b200.infer.inferer import InfererContext
from pgmax_b200.infer.inferer import _tree_map


@dataclasses.dataclass(frozen=True, eq=False)
class BeliefPropagation(Inferer):
  """Belief propagation functions (pgmax/infer/bp.py:36-46) plus the fused
  device-side decode and the host-buffer end-to-end call."""

  run_bp: Callable[..., BPArrays]
  run_with_diffs: Callable[..., Tuple[BPArrays, Any]]
  get_map_states: Callable[..., Dict[Hashable, Any]] = None
  infer_host: Callable[..., Dict[str, Any]] = None
  context: InfererContext = None


def BP(bp_state: BPState, temperature: Optional[float] = 0.0) -> BeliefPropagation:  # pylint: disable=invalid-name
  """Belief propagation functions for ``bp_state``.

  Args:
    bp_state: belief propagation state.
    temperature: default temperature; 1.0 = sum-product, 0.0 = max-product.
  """
  context = InfererContext(bp_state)
  default_temperature = temperature

  def run_with_diffs(
      bp_arrays: BPArrays,
      num_iters: int,
      damping: float = 0.5,
      temperature: float = default_temperature,
  ) -> Tuple[BPArrays, Any]:
    """``num_iters`` damped BP iterations; also returns max|m' - m| per iteration
    ([num_iters] or [B, num_iters]).  Potentials are clipped to +-1e6 and the
    input messages normalised inside the call; the returned BPArrays carries the
    caller's potentials and evidence unchanged (pgmax/infer/bp.py:63-155)."""
    import torch  # pylint: disable=g-import-not-at-top

    # The reference runs one update for num_iters <= 1 (bp.py:142-146).
    num_iters = max(int(num_iters), 1)
    buf = DeviceBuffers(bp_arrays, context._device(), context.lp_cache)
    plan = context.plan
    batch = buf.batch or 1
    out = torch.empty((batch, plan.num_edge_states), dtype=torch.float32, device=buf.device)
    deltas = torch.empty((batch, num_iters), dtype=torch.float32, device=buf.device)
    stream = torch.cuda.current_stream(buf.device).cuda_stream
    plan.bp_run(stream, batch, buf.lp.data_ptr(), buf.lp.ndim == 2, buf.ev.data_ptr(),
                buf.ev.ndim == 2, buf.msgs.data_ptr(), buf.msgs.ndim == 2, out.data_ptr(),
                deltas.data_ptr(), num_iters, float(damping), float(temperature),
                flags=plan.RUN_POTENTIALS_UNCHANGED if buf.lp_unchanged else 0)
    if buf.batch is None:
      out, deltas = out[0], deltas[0]
    new_arrays = BPArrays(
        log_potentials=bp_arrays.log_potentials,
        ftov_msgs=buf.out(out),
        evidence=bp_arrays.evidence,
    )
    return new_arrays, buf.out(deltas)

  def run(
      bp_arrays: BPArrays,
      num_iters: int,
      damping: float = 0.5,
      temperature: float = default_temperature,
  ) -> BPArrays:
    """run_with_diffs without the per-iteration deltas (pgmax/infer/bp.py:157-165)."""
    import torch  # pylint: disable=g-import-not-at-top

    num_iters = max(int(num_iters), 1)
    buf = DeviceBuffers(bp_arrays, context._device(), context.lp_cache)
    plan = context.plan
    batch = buf.batch or 1
    out = torch.empty((batch, plan.num_edge_states), dtype=torch.float32, device=buf.device)
    stream = torch.cuda.current_stream(buf.device).cuda_stream
    plan.bp_run(stream, batch, buf.lp.data_ptr(), buf.lp.ndim == 2, buf.ev.data_ptr(),
                buf.ev.ndim == 2, buf.msgs.data_ptr(), buf.msgs.ndim == 2, out.data_ptr(),
                None, num_iters, float(damping), float(temperature),
                flags=plan.RUN_POTENTIALS_UNCHANGED if buf.lp_unchanged else 0)
    return BPArrays(
        log_potentials=bp_arrays.log_potentials,
        ftov_msgs=buf.out(out if buf.batch is not None else out[0]),
        evidence=bp_arrays.evidence,
    )

  def run_bp(bp_arrays: BPArrays, num_iters: int, damping: float = 0.5) -> BPArrays:
    """Deprecated alias of run (pgmax/infer/bp.py:167-176)."""
    warnings.warn("BP.run_bp is deprecated. Please consider using BP.run instead.")
    return run(bp_arrays, num_iters, damping, default_temperature)

  def get_map_states(bp_arrays: BPArrays, return_ties: bool = False):
    """MAP states per VarGroup straight from BPArrays (fused device decode)."""
    states, _, ties = context.decode(bp_arrays)
    decoded = context.unflatten_states(states)
    return (decoded, ties) if return_ties else decoded

  def infer_host(
      bp_arrays: BPArrays,
      num_iters: int,
      damping: float = 0.5,
      temperature: float = default_temperature,
      marginals: bool = False,
      return_msgs: bool = False,
  ) -> Dict[str, Any]:
    """init -> run -> beliefs -> decode in ONE C-ABI call on host (numpy) buffers
    (pgx_infer_host): the timed region of the reference's benchmark harness
    (benchmark/rbm_lib.py:180-187) including host<->device copies.
    ``bp_arrays.ftov_msgs`` that is all zeros may be passed as None-equivalent by
    constructing BPArrays with ftov_msgs=None."""
    import torch  # pylint: disable=g-import-not-at-top

    context._device()
    plan = context.plan
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
    lp, ev, msgs = f32(bp_arrays.log_potentials), f32(bp_arrays.evidence), f32(bp_arrays.ftov_msgs)
    sizes = {int(a.shape[0]) for a in (lp, ev, msgs) if a is not None and a.ndim == 2}
    if len(sizes) > 1:
      raise ValueError(f"Inconsistent batch sizes: {sorted(sizes)}")
    batched = bool(sizes)
    batch = sizes.pop() if sizes else 1
    num_iters = max(int(num_iters), 1)
    states = np.empty((batch, plan.num_vars), dtype=np.int32)
    ties = np.empty((batch,), dtype=np.int32)
    marg = np.empty((batch, plan.num_var_states), dtype=np.float32) if marginals else None
    out_msgs = np.empty((batch, plan.num_edge_states), dtype=np.float32) if return_msgs else None
    ptr = lambda a: None if a is None else a.ctypes.data
    stream = torch.cuda.current_stream(context._device()).cuda_stream
    plan.infer_host(stream, batch, ptr(lp), lp.ndim == 2, ptr(ev), ev.ndim == 2, ptr(msgs),
                    msgs is not None and msgs.ndim == 2, num_iters, float(damping),
                    float(temperature), ptr(states), ptr(marg), ptr(ties), ptr(out_msgs), None)
    sq = (lambda a: a) if batched else (lambda a: None if a is None else a[0])
    return {
        "map_states": context.unflatten_states(sq(states)),
        "flat_map_states": sq(states),
        "tie_counts": sq(ties),
        "marginals": sq(marg),
        "ftov_msgs": sq(out_msgs),
    }

  return BeliefPropagation(
      init=context.init,
      update=context.update,
      to_bp_state=context.to_bp_state,
      get_beliefs=context.get_beliefs,
      run=run,
      run_bp=run_bp,
      run_with_diffs=run_with_diffs,
      get_map_states=get_map_states,
      infer_host=infer_host,
      context=context,
  )


def get_marginals(beliefs: Dict[Hashable, Any]) -> Dict[Hashable, Any]:
  """Softmax of the beliefs of every variable: exp(x - logsumexp(x, -1)); -inf
  padding of ragged groups maps to probability 0 (pgmax/infer/bp.py:263-288)."""

  def normalise(x):
    if x.size == 0:
      return x
    mx = np.max(x, axis=-1, keepdims=True)
    mx = np.where(np.isfinite(mx), mx, 0.0)
    lse = mx + np.log(np.sum(np.exp(x - mx), axis=-1, keepdims=True))
    return np.exp(x - lse)

  return _tree_map(normalise, beliefs)
