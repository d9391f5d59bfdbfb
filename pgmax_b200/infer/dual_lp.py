"""Smooth dual LP-MAP solver (host mirror of pgmax/infer/dual_lp.py:33-463).

``SDLP(bp_state)`` returns the same bundle of functions as the reference.  Every
(sub)gradient step runs on the device through the C ABI (pgx_sdlp_run /
pgx_sdlp_objval_and_grad); this module only validates arguments, chooses the learning
rate and forms the reference's per-iteration fp32 scalars.
"""

import dataclasses
from typing import Any, Callable, Dict, Hashable, Optional, Tuple

import numpy as np

from pgmax_b200.infer import energy as energy_lib
from pgmax_b200.infer.bp_state import BPArrays
from pgmax_b200.infer.bp_state import BPState
from pgmax_b200.infer.inferer import decode_map_states
from pgmax_b200.infer.inferer import DeviceBuffers
from pgmax_b200.infer.inferer import Inferer
from pgmax_b200.infer.inferer import InfererContext


@dataclasses.dataclass(frozen=True, eq=False)
class SmoothDualLP(Inferer):
  """Smooth Dual LP-MAP solver functions (pgmax/infer/dual_lp.py:33-57)."""

  run_with_objvals: Callable[..., Tuple[BPArrays, Any]]
  decode_primal_unaries: Callable[..., Tuple[Dict[Hashable, Any], Dict[Hashable, Any]]]
  get_primal_upper_bound: Callable[..., Any]
  get_map_lower_bound: Callable[..., Any]
  get_bp_updates: Callable[..., Tuple[Any, Any]]
  # smooth_dual_objval_and_grad itself (a closure in the reference): dict with "objval" and the
  # requested extras among "grad", "bp_updates", "edge_vals"
  objval_and_grad: Callable[..., Dict[str, Any]] = None
  context: InfererContext = None


def step_schedule(num_iters: int, lr: float, logsumexp_temp: float) -> Tuple[np.ndarray, np.ndarray]:
  """fp32 step sizes and Nesterov momenta of every iteration, as the traced reference forms
  them (dual_lp.py:293-309): step = lr (T > 0) or lr / sqrt(it + 1) (T == 0);
  momentum = (it + 1) / (it + 4), with ``it`` promoted to fp32."""
  it = np.arange(num_iters, dtype=np.float32)
  one, four = np.float32(1.0), np.float32(4.0)
  if logsumexp_temp > 0:
    steps = np.full((num_iters,), np.float32(lr), dtype=np.float32)
  else:
    steps = (np.float32(lr) / np.sqrt(it + one, dtype=np.float32)).astype(np.float32)
  momenta = ((it + one) / (it + four)).astype(np.float32)
  return steps, momenta


def SDLP(bp_state: BPState) -> SmoothDualLP:  # pylint: disable=invalid-name
  """Smooth Dual LP-MAP functions for ``bp_state`` (pgmax/infer/dual_lp.py:60-463)."""
  context = InfererContext(bp_state)
  energy_lib.register_context(context)  # get_map_lower_bound reuses this context's device plan
  state = {"factors_set": False}

  def plan():
    """The device plan with the factor membership of its edges registered."""
    p = context.plan
    if not state["factors_set"]:
      p.set_factors(context.factor_edge_start)
      state["factors_set"] = True
    return p

  def objval_and_grad(sdlp_arrays: BPArrays, logsumexp_temp: float, want: Tuple[str, ...] = ()):
    """smooth_dual_objval_and_grad (dual_lp.py:67-237): objval plus the requested extras
    ("grad", "bp_updates", "edge_vals")."""
    import torch  # pylint: disable=g-import-not-at-top

    p = plan()
    buf = DeviceBuffers(sdlp_arrays, context._device())  # pylint: disable=protected-access
    batch = buf.batch or 1
    new = lambda n: torch.empty((batch, n), dtype=torch.float32, device=buf.device)
    objval = torch.empty((batch,), dtype=torch.float32, device=buf.device)
    grad = new(p.num_edge_states) if "grad" in want else None
    updates = new(p.num_edge_states) if "bp_updates" in want else None
    edge_vals = new(context.num_edges) if "edge_vals" in want else None
    ptr = lambda t: None if t is None else t.data_ptr()
    stream = torch.cuda.current_stream(buf.device).cuda_stream
    p.sdlp_objval_and_grad(stream, batch, buf.lp.data_ptr(), buf.lp.ndim == 2, buf.ev.data_ptr(),
                           buf.ev.ndim == 2, buf.msgs.data_ptr(), buf.msgs.ndim == 2,
                           float(logsumexp_temp), objval.data_ptr(), ptr(grad), ptr(updates),
                           ptr(edge_vals))
    sq = (lambda t: t) if buf.batch is not None else (lambda t: t[0])
    out = {"objval": buf.out(objval) if buf.batch is not None else float(objval[0].item())}
    for name, t in (("grad", grad), ("bp_updates", updates), ("edge_vals", edge_vals)):
      if t is not None:
        out[name] = buf.out(sq(t))
    return out

  def run_with_objvals(
      sdlp_arrays: BPArrays,
      logsumexp_temp: float,
      num_iters: int,
      lr: Optional[float] = None,
  ) -> Tuple[BPArrays, Any]:
    """Accelerated gradient descent (or subgradient descent for logsumexp_temp == 0) on the
    dual messages; also returns the objective value at each step ([num_iters] or
    [B, num_iters]).  Raises ValueError as the reference (dual_lp.py:266-277)."""
    import torch  # pylint: disable=g-import-not-at-top

    if logsumexp_temp < 0.0 or logsumexp_temp > 1.0:
      raise ValueError(
          "The log sum-exp temperature of the Dual LP-MAP solver has to be"
          " between 0.0 and 1.0"
      )
    if logsumexp_temp != 0.0 and lr is not None and lr > logsumexp_temp:
      raise ValueError(
          "For gradient descent, the learning rate must be smaller than the"
          " log sum-exp temperature."
      )
    if lr is None:
      lr = logsumexp_temp if logsumexp_temp != 0.0 else 0.01
    num_iters = int(num_iters)
    p = plan()
    buf = DeviceBuffers(sdlp_arrays, context._device())  # pylint: disable=protected-access
    batch = buf.batch or 1
    out = torch.empty((batch, p.num_edge_states), dtype=torch.float32, device=buf.device)
    objvals = torch.empty((batch, max(num_iters, 1)), dtype=torch.float32, device=buf.device)
    steps, momenta = step_schedule(num_iters, lr, logsumexp_temp)
    stream = torch.cuda.current_stream(buf.device).cuda_stream
    p.sdlp_run(stream, batch, buf.lp.data_ptr(), buf.lp.ndim == 2, buf.ev.data_ptr(),
               buf.ev.ndim == 2, buf.msgs.data_ptr(), buf.msgs.ndim == 2, out.data_ptr(),
               objvals.data_ptr(), steps, momenta, float(logsumexp_temp))
    objvals = objvals[:, :num_iters]
    if buf.batch is None:
      out, objvals = out[0], objvals[0]
    new_arrays = BPArrays(
        log_potentials=sdlp_arrays.log_potentials,
        ftov_msgs=buf.out(out),
        evidence=sdlp_arrays.evidence,
    )
    return new_arrays, buf.out(objvals)

  def run(
      sdlp_arrays: BPArrays,
      logsumexp_temp: float,
      num_iters: int,
      lr: Optional[float] = None,
  ) -> BPArrays:
    """run_with_objvals without the objective values (dual_lp.py:326-339)."""
    return run_with_objvals(sdlp_arrays, logsumexp_temp, num_iters, lr)[0]

  def decode_primal_unaries(sdlp_arrays: BPArrays):
    """Local decoding of the primal unaries from the dual beliefs (dual_lp.py:341-364)."""
    dual_beliefs = context.get_beliefs(sdlp_arrays)
    return decode_map_states(dual_beliefs), dual_beliefs

  def get_primal_upper_bound(sdlp_arrays: BPArrays):
    """Upper bound of the optimal LP-MAP objective: the dual objective at temperature 0
    (dual_lp.py:366-378)."""
    return objval_and_grad(sdlp_arrays, 0.0)["objval"]

  def get_map_lower_bound(sdlp_arrays: BPArrays, decoded_primal_unaries, debug_mode=False):
    """Minus the energy of the decoding (dual_lp.py:380-409)."""
    energy = energy_lib.compute_energy(
        bp_state=bp_state, bp_arrays=sdlp_arrays, map_states=decoded_primal_unaries,
        debug_mode=debug_mode)[0]
    return -energy

  def get_bp_updates(sdlp_arrays: BPArrays, logsumexp_temp: float):
    """(BP updates, per-edge logsumexp of the outgoing messages) - the quantities the
    reference exposes for its unit tests (dual_lp.py:411-445)."""
    out = objval_and_grad(sdlp_arrays, logsumexp_temp, want=("bp_updates", "edge_vals"))
    return out["bp_updates"], out["edge_vals"]

  return SmoothDualLP(
      init=context.init,
      update=context.update,
      to_bp_state=context.to_bp_state,
      get_beliefs=context.get_beliefs,
      run=run,
      run_with_objvals=run_with_objvals,
      decode_primal_unaries=decode_primal_unaries,
      get_primal_upper_bound=get_primal_upper_bound,
      get_map_lower_bound=get_map_lower_bound,
      get_bp_updates=get_bp_updates,
      objval_and_grad=objval_and_grad,
      context=context,
  )
