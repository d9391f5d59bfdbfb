"""Inference context: flat layout, device plan, beliefs and decoding
(host mirror of pgmax/infer/inferer.py:36-264).

The context owns the device plan (pgmax_b200._native.Plan), created lazily on
the first call that needs the GPU so that graph construction and BPArrays
manipulation work on a CPU-only box.
"""

import dataclasses
from typing import Any, Callable, Dict, Hashable, Optional, Sequence

import numpy as np

from pgmax_b200 import factor
from pgmax_b200 import vgroup
from pgmax_b200.infer import bp_state as bpstate
from pgmax_b200.infer.bp_state import BPArrays
from pgmax_b200.infer.bp_state import BPState
from pgmax_b200.infer.bp_state import _is_torch


@dataclasses.dataclass(frozen=True, eq=False)
class Inferer:
  """Functions shared by all inferers (pgmax/infer/inferer.py:36-52)."""

  init: Callable[..., BPArrays]
  update: Callable[..., BPArrays]
  to_bp_state: Callable[..., BPArrays]
  get_beliefs: Callable[..., Dict[Hashable, Any]]
  run: Callable[..., BPArrays]


class DeviceBuffers:
  """Moves BPArrays fields to fp32 CUDA tensors and remembers what came from the host."""

  def __init__(self, bp_arrays: BPArrays, device, lp_cache: Optional[dict] = None):
    import torch  # pylint: disable=g-import-not-at-top

    self.torch = torch
    self.device = device
    self.from_host = not _is_torch(bp_arrays.ftov_msgs)
    self.batch = bp_arrays.batch_size
    self.lp, self.lp_unchanged = self._put_potentials(bp_arrays.log_potentials, lp_cache)
    self.ev = self._put(bp_arrays.evidence)
    self.msgs = self._put(bp_arrays.ftov_msgs)

  def _put_potentials(self, arr, cache: Optional[dict]):
    """Potentials on the device + whether they are KNOWN to be what the previous run of this
    inferer used (same immutable host array object -> the cached device copy is reused, no
    upload; same device tensor, same torch version counter): per-run preprocessing of the
    potentials is then skipped (PGX_RUN_POTENTIALS_UNCHANGED)."""
    if cache is None:
      return self._put(arr), False
    if _is_torch(arr):
      t = self._put(arr)
      key = ("torch", t.data_ptr(), int(getattr(arr, "_version", -1)), tuple(t.shape), t is arr or t.data_ptr() == arr.data_ptr())
      unchanged = key[-1] and cache.get("key") == key
      cache.update(key=key, host=None, dev=None)
      return t, bool(unchanged)
    if cache.get("host") is arr and not arr.flags.writeable and cache.get("dev") is not None:
      return cache["dev"], True
    t = self._put(arr)
    cache.update(key=None, host=arr if not arr.flags.writeable else None, dev=t)
    return t, False

  def _put(self, arr):
    torch = self.torch
    if _is_torch(arr):
      t = arr.to(device=self.device, dtype=torch.float32)
    else:
      host = np.ascontiguousarray(arr, dtype=np.float32)
      if not host.flags.writeable:  # BPArrays are read-only views; torch wants a writable buffer
        host = host.copy()
      t = torch.from_numpy(host).to(self.device)
    return t.contiguous()

  def out(self, tensor):
    """Result in the same kind of memory the caller's messages live in."""
    return tensor.cpu().numpy() if self.from_host else tensor


class InfererContext:
  """Flat index arrays of a BPState + the device plan built from them."""

  def __init__(self, bp_state: BPState):
    self.bp_state = bp_state
    fg_state = bp_state.fg_state
    self.wiring = fg_state.wiring
    self.evidence_to_vars = fg_state.evidence_to_vars
    self.factor_type_to_msgs_range = fg_state.factor_type_to_msgs_range
    self.factor_type_to_potentials_range = fg_state.factor_type_to_potentials_range
    self.num_variables = sum(
        int(np.prod(vg.num_states.shape)) for vg in fg_state.variable_groups
    )
    self.num_edges = sum(w.num_edges for w in self.wiring.values())
    self.num_factors = sum(w.num_factors for w in self.wiring.values())
    self._flat = None
    self._plan = None
    self.lp_cache = {}  # DeviceBuffers._put_potentials

  # ---- reference-format views (oracle / tests), pgmax/infer/inferer.py:76-98 ----
  def _flat_arrays(self) -> np.ndarray:
    if self._flat is None:
      self._flat = factor.concatenate_var_states_for_edges(
          [self.wiring[ft].var_states_for_edges for ft in factor.FACTOR_TYPES]
      )
    return self._flat

  @property
  def var_states_for_edge_states(self) -> np.ndarray:
    return self._flat_arrays()[:, 0]

  @property
  def edge_indices_for_edge_states(self) -> np.ndarray:
    return self._flat_arrays()[:, 1]

  @property
  def factor_indices_for_edge_states(self) -> np.ndarray:
    return self._flat_arrays()[:, 2]

  @property
  def factor_edge_start(self) -> np.ndarray:
    """[num_factors + 1] offsets: edges [start[f], start[f + 1]) belong to factor f (the edges of
    a factor are contiguous in every compiled wiring; the per-edge form of column 2 above)."""
    shift, ids = 0, []
    for ft in factor.FACTOR_TYPES:
      w = self.wiring[ft]
      ids.append(np.asarray(w.edge_factor, dtype=np.int64) + shift)
      shift += w.num_factors
    ids = np.concatenate(ids) if ids else np.zeros((0,), dtype=np.int64)
    if ids.size and np.any(np.diff(ids) < 0):
      raise ValueError("The edges of a factor must be contiguous in the flat layout")
    counts = np.bincount(ids, minlength=self.num_factors)
    return np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)

  @property
  def inference_arguments(self) -> Dict[Any, Dict[str, Any]]:
    return {ft: self.wiring[ft].get_inference_arguments() for ft in factor.FACTOR_TYPES}

  # ---- device plan ----
  @property
  def plan(self):
    if self._plan is None:
      from pgmax_b200 import _native  # pylint: disable=g-import-not-at-top

      self._plan = _native.Plan(self.bp_state.fg_state)
    return self._plan

  def _device(self):
    import torch  # pylint: disable=g-import-not-at-top

    if not torch.cuda.is_available():
      from pgmax_b200 import _native  # pylint: disable=g-import-not-at-top

      raise _native.PgxError(
          _native.PGX_ERR_NO_DEVICE,
          "no CUDA device: the BP kernels have no CPU fallback",
      )
    return torch.device("cuda", self.plan.device)

  # ---- BPArrays construction (pgmax/infer/inferer.py:120-189) ----
  def update(
      self,
      bp_arrays: Optional[BPArrays] = None,
      log_potentials_updates: Optional[Dict[Any, Any]] = None,
      ftov_msgs_updates: Optional[Dict[Any, Any]] = None,
      evidence_updates: Optional[Dict[Any, Any]] = None,
  ) -> BPArrays:
    """BPArrays with the given updates applied; a leading batch axis on any update
    makes that array batched."""
    fg_state = self.bp_state.fg_state
    if bp_arrays is not None:
      lp, msgs, ev = bp_arrays.log_potentials, bp_arrays.ftov_msgs, bp_arrays.evidence
    else:
      lp = np.asarray(self.bp_state.log_potentials.value, dtype=np.float32)
      msgs = np.asarray(self.bp_state.ftov_msgs.value, dtype=np.float32)
      ev = np.asarray(self.bp_state.evidence.value, dtype=np.float32)
    if log_potentials_updates is not None:
      lp = bpstate.update_log_potentials(lp, log_potentials_updates, fg_state)
    if ftov_msgs_updates is not None:
      msgs = bpstate.update_ftov_msgs(msgs, ftov_msgs_updates, fg_state)
    if evidence_updates is not None:
      ev = bpstate.update_evidence(ev, evidence_updates, fg_state)
    return BPArrays(log_potentials=lp, ftov_msgs=msgs, evidence=ev)

  def init(
      self,
      log_potentials_updates: Optional[Dict[Any, Any]] = None,
      ftov_msgs_updates: Optional[Dict[Any, Any]] = None,
      evidence_updates: Optional[Dict[Any, Any]] = None,
  ) -> BPArrays:
    return self.update(
        bp_arrays=None,
        log_potentials_updates=log_potentials_updates,
        ftov_msgs_updates=ftov_msgs_updates,
        evidence_updates=evidence_updates,
    )

  def to_bp_state(self, bp_arrays: BPArrays) -> BPState:
    """BPState rebuilt from (unbatched) BPArrays (pgmax/infer/inferer.py:191-209)."""
    fg_state = self.bp_state.fg_state
    host = lambda a: a.detach().cpu().numpy() if _is_torch(a) else np.asarray(a)
    return BPState(
        log_potentials=bpstate.LogPotentials(fg_state=fg_state, value=host(bp_arrays.log_potentials)),
        ftov_msgs=bpstate.FToVMessages(fg_state=fg_state, value=host(bp_arrays.ftov_msgs)),
        evidence=bpstate.Evidence(fg_state=fg_state, value=host(bp_arrays.evidence)),
    )

  # ---- device calls ----
  def flat_beliefs(self, bp_arrays: BPArrays):
    """evidence + incoming messages, flat [V_s] or [B, V_s] (pgx_beliefs)."""
    import torch  # pylint: disable=g-import-not-at-top

    buf = DeviceBuffers(bp_arrays, self._device())
    batch = buf.batch or 1
    out = torch.empty((batch, self.plan.num_var_states), dtype=torch.float32, device=buf.device)
    stream = torch.cuda.current_stream(buf.device).cuda_stream
    self.plan.beliefs(stream, batch, buf.ev.data_ptr(), buf.ev.ndim == 2, buf.msgs.data_ptr(),
                      buf.msgs.ndim == 2, out.data_ptr())
    return buf.out(out if buf.batch is not None else out[0])

  def get_beliefs(self, bp_arrays: BPArrays) -> Dict[Hashable, Any]:
    """Beliefs per VarGroup (pgmax/infer/inferer.py:211-225)."""
    flat = self.flat_beliefs(bp_arrays)
    if _is_torch(flat):
      flat = flat.cpu().numpy()
    return unflatten_beliefs(flat, self.bp_state.fg_state.variable_groups)

  def decode(self, bp_arrays: BPArrays, marginals: bool = False):
    """Fused beliefs + MAP (+ marginals) on the device (pgx_decode).

    Returns (map_states, marginals_or_None, tie_counts): flat per-variable int32
    states [num_vars] or [B, num_vars], flat marginals [V_s] or [B, V_s], and the
    number of variables per sample whose two best beliefs are exactly tied.
    """
    import torch  # pylint: disable=g-import-not-at-top

    buf = DeviceBuffers(bp_arrays, self._device())
    batch = buf.batch or 1
    plan = self.plan
    states = torch.empty((batch, plan.num_vars), dtype=torch.int32, device=buf.device)
    ties = torch.empty((batch,), dtype=torch.int32, device=buf.device)
    marg = (
        torch.empty((batch, plan.num_var_states), dtype=torch.float32, device=buf.device)
        if marginals
        else None
    )
    stream = torch.cuda.current_stream(buf.device).cuda_stream
    plan.decode(stream, batch, buf.ev.data_ptr(), buf.ev.ndim == 2, buf.msgs.data_ptr(),
                buf.msgs.ndim == 2, states.data_ptr(),
                marg.data_ptr() if marg is not None else None, ties.data_ptr())
    squeeze = (lambda t: t) if buf.batch is not None else (lambda t: t[0])
    return (
        buf.out(squeeze(states)),
        buf.out(squeeze(marg)) if marg is not None else None,
        buf.out(squeeze(ties)),
    )

  def unflatten_states(self, flat_states) -> Dict[Hashable, Any]:
    """Flat per-variable values -> dict VarGroup -> structured array."""
    flat_states = flat_states.cpu().numpy() if _is_torch(flat_states) else np.asarray(flat_states)
    out, start = {}, 0
    for vg in self.bp_state.fg_state.variable_groups:
      n = int(np.prod(vg.num_states.shape))
      out[vg] = _unflatten(vg, flat_states[..., start : start + n], False)
      start += n
    return out


def _unflatten(vg, flat, per_state: bool):
  """vg.unflatten, or its batched form for [B, n] data (VarDict.unflatten itself keeps the
  reference's 1-D-only contract)."""
  if getattr(flat, "ndim", 1) == 2 and hasattr(vg, "unflatten_batch"):
    return vg.unflatten_batch(flat, per_state)
  return vg.unflatten(flat, per_state)


def unflatten_beliefs(flat_beliefs, variable_groups: Sequence[vgroup.VarGroup]) -> Dict[Hashable, Any]:
  """Flat beliefs -> dict VarGroup -> structured beliefs (pgmax/infer/inferer.py:228-248)."""
  beliefs, start = {}, 0
  for vg in variable_groups:
    length = int(vg.num_states.sum())
    beliefs[vg] = _unflatten(vg, flat_beliefs[..., start : start + length], True)
    start += length
  return beliefs


def _tree_map(fn, tree):
  if isinstance(tree, dict):
    return {k: _tree_map(fn, v) for k, v in tree.items()}
  return fn(np.asarray(tree))


def decode_map_states(beliefs: Dict[Hashable, Any]) -> Dict[Hashable, Any]:
  """First arg-max over the last axis of every beliefs array; empty arrays decode
  to zeros (pgmax/infer/inferer.py:251-264).  Host-side convenience over the
  dict returned by get_beliefs; the fused device path is BP(...).get_map_states."""
  return _tree_map(
      lambda x: np.argmax(x, axis=-1) if x.size > 0 else np.zeros(x.shape[:-1]),
      beliefs,
  )
