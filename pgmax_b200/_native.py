"""ctypes binding of the C ABI in include/pgx.h (csrc/libpgx.so).

This module is the only place the Python host mirror touches native code.
It never computes messages itself and has no fallback: if the shared library
is missing, or no CUDA device is visible, calls raise ``PgxError``.
"""

import ctypes
import dataclasses
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PGX_LIB: another build of the same library (A/B runs of kernel variants)
LIB_PATH = os.environ.get("PGX_LIB") or os.path.join(_HERE, "csrc", "libpgx.so")
BUILD_SCRIPT = os.path.join(_HERE, "csrc", "build.sh")

PGX_OK = 0
PGX_ERR_INVALID = -1
PGX_ERR_CUDA = -2
PGX_ERR_UNSUPPORTED = -3
PGX_ERR_NO_DEVICE = -4

# Every symbol include/pgx.h declares; tests/test_abi.py checks the .so exports all of them.
EXPORTED_SYMBOLS = (
    "pgx_plan_create",
    "pgx_plan_destroy",
    "pgx_plan_get_info",
    "pgx_bp_run",
    "pgx_bp_run_flags",
    "pgx_beliefs",
    "pgx_decode",
    "pgx_decode_last_run",
    "pgx_energy",
    "pgx_infer_host",
    "pgx_plan_set_factors",
    "pgx_sdlp_objval_and_grad",
    "pgx_sdlp_run",
    "pgx_plan_launch_count",
    "pgx_plan_set_exact_order",
    "pgx_plan_num_fused_blocks",
    "pgx_plan_compressed_edges",
    "pgx_plan_disable_paths",
    "pgx_plan_is_lattice",
    "pgx_plan_enable_graphs",
    "pgx_plan_graph_launch_count",
    "pgx_plan_dominant_edge_states",
    "pgx_plan_dominant_grid",
    "pgx_plan_profile_enable",
    "pgx_plan_profile_read",
    "pgx_last_error",
    "pgx_build_info",
    "pgx_bp_run_vjp",
    "pgx_nccl_unique_id",
    "pgx_strip_create",
    "pgx_strip_destroy",
    "pgx_strip_run",
    "pgx_strip_beliefs",
    "pgx_strip_launch_count",
    "pgx_strip_graph_launch_count",
)


class PgxError(RuntimeError):
  """A pgx_* call failed; ``code`` is the pgx_status."""

  def __init__(self, code: int, message: str):
    super().__init__(f"pgx error {code}: {message}")
    self.code = code


_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)


class EnumBlockC(ctypes.Structure):
  _fields_ = [
      ("num_factors", ctypes.c_int64),
      ("arity", ctypes.c_int32),
      ("num_configs", ctypes.c_int32),
      ("configs", _i32p),
      ("first_edge", ctypes.c_int64),
      ("first_msg", ctypes.c_int64),
      ("first_potential", ctypes.c_int64),
  ]


class LogicalDescC(ctypes.Structure):
  _fields_ = [
      ("num_factors", ctypes.c_int64),
      ("num_parents", ctypes.c_int64),
      ("parents_factor", _i32p),
      ("parents_msg", _i32p),
      ("children_msg", _i32p),
      ("edge_states_offset", ctypes.c_int32),
  ]


class GraphDescC(ctypes.Structure):
  _fields_ = [
      ("num_vars", ctypes.c_int64),
      ("var_num_states", _i32p),
      ("num_edges", ctypes.c_int64),
      ("edge_var_start", _i32p),
      ("edge_num_states", _i32p),
      ("num_potentials", ctypes.c_int64),
      ("num_enum_blocks", ctypes.c_int32),
      ("enum_blocks", ctypes.POINTER(EnumBlockC)),
      ("or_factors", LogicalDescC),
      ("and_factors", LogicalDescC),
      ("pool_factors", LogicalDescC),
  ]


class PlanInfoC(ctypes.Structure):
  _fields_ = [
      ("num_vars", ctypes.c_int64),
      ("num_var_states", ctypes.c_int64),
      ("num_edges", ctypes.c_int64),
      ("num_edge_states", ctypes.c_int64),
      ("num_potentials", ctypes.c_int64),
      ("max_var_states", ctypes.c_int64),
      ("device_bytes", ctypes.c_int64),
      ("device", ctypes.c_int32),
      ("num_sms", ctypes.c_int32),
  ]


_lib = None


def build(verbose: bool = False) -> str:
  """Compiles csrc/libpgx.so for sm_100a with nvcc (cross-compiles without a GPU)."""
  out = subprocess.run(
      ["bash", BUILD_SCRIPT], check=True, capture_output=not verbose, text=True
  )
  if verbose and out.stdout:
    print(out.stdout)
  return LIB_PATH


def load() -> ctypes.CDLL:
  """Loads libpgx.so and declares the signatures of include/pgx.h."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise PgxError(
        PGX_ERR_NO_DEVICE,
        f"{LIB_PATH} is missing: build it with pgmax_b200/csrc/build.sh "
        "(there is no CPU fallback for the BP kernels)",
    )
  lib = ctypes.CDLL(LIB_PATH)
  vp, i64, i32, f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float
  lib.pgx_plan_create.argtypes = [ctypes.POINTER(GraphDescC), ctypes.POINTER(vp)]
  lib.pgx_plan_create.restype = ctypes.c_int
  lib.pgx_plan_destroy.argtypes = [vp]
  lib.pgx_plan_destroy.restype = None
  lib.pgx_plan_get_info.argtypes = [vp, ctypes.POINTER(PlanInfoC)]
  lib.pgx_plan_get_info.restype = ctypes.c_int
  lib.pgx_bp_run.argtypes = [vp, vp, i64, vp, ctypes.c_int, vp, ctypes.c_int, vp,
                             ctypes.c_int, vp, vp, i32, f32, f32]
  lib.pgx_bp_run.restype = ctypes.c_int
  lib.pgx_bp_run_flags.argtypes = lib.pgx_bp_run.argtypes + [ctypes.c_uint32]
  lib.pgx_bp_run_flags.restype = ctypes.c_int
  lib.pgx_beliefs.argtypes = [vp, vp, i64, vp, ctypes.c_int, vp, ctypes.c_int, vp]
  lib.pgx_beliefs.restype = ctypes.c_int
  lib.pgx_decode.argtypes = [vp, vp, i64, vp, ctypes.c_int, vp, ctypes.c_int, vp, vp, vp]
  lib.pgx_decode.restype = ctypes.c_int
  lib.pgx_decode_last_run.argtypes = [vp, vp, i64, vp, vp, vp]
  lib.pgx_decode_last_run.restype = ctypes.c_int
  lib.pgx_energy.argtypes = [vp, vp, i64, vp, ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, vp]
  lib.pgx_energy.restype = ctypes.c_int
  lib.pgx_infer_host.argtypes = [vp, vp, i64, vp, ctypes.c_int, vp, ctypes.c_int, vp,
                                 ctypes.c_int, i32, f32, f32, vp, vp, vp, vp, vp]
  lib.pgx_infer_host.restype = ctypes.c_int
  lib.pgx_plan_set_factors.argtypes = [vp, i64, _i32p]
  lib.pgx_plan_set_factors.restype = ctypes.c_int
  lib.pgx_sdlp_objval_and_grad.argtypes = [vp, vp, i64, vp, ctypes.c_int, vp, ctypes.c_int, vp,
                                           ctypes.c_int, f32, vp, vp, vp, vp]
  lib.pgx_sdlp_objval_and_grad.restype = ctypes.c_int
  lib.pgx_sdlp_run.argtypes = [vp, vp, i64, vp, ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int,
                               vp, vp, i32, _f32p, _f32p, f32]
  lib.pgx_sdlp_run.restype = ctypes.c_int
  lib.pgx_plan_launch_count.argtypes = [vp]
  lib.pgx_plan_launch_count.restype = ctypes.c_int64
  lib.pgx_plan_num_fused_blocks.argtypes = [vp]
  lib.pgx_plan_num_fused_blocks.restype = ctypes.c_int
  lib.pgx_plan_compressed_edges.argtypes = [vp]
  lib.pgx_plan_compressed_edges.restype = ctypes.c_int64
  lib.pgx_plan_set_exact_order.argtypes = [vp, ctypes.c_int]
  lib.pgx_plan_set_exact_order.restype = ctypes.c_int
  lib.pgx_plan_disable_paths.argtypes = [vp, ctypes.c_uint32]
  lib.pgx_plan_disable_paths.restype = ctypes.c_int
  lib.pgx_plan_is_lattice.argtypes = [vp]
  lib.pgx_plan_is_lattice.restype = ctypes.c_int
  lib.pgx_plan_enable_graphs.argtypes = [vp, ctypes.c_int]
  lib.pgx_plan_enable_graphs.restype = ctypes.c_int
  lib.pgx_plan_graph_launch_count.argtypes = [vp]
  lib.pgx_plan_graph_launch_count.restype = ctypes.c_int64
  lib.pgx_plan_dominant_edge_states.argtypes = [vp]
  lib.pgx_plan_dominant_edge_states.restype = ctypes.c_int64
  lib.pgx_plan_dominant_grid.argtypes = [vp]
  lib.pgx_plan_dominant_grid.restype = ctypes.c_int64
  lib.pgx_plan_profile_enable.argtypes = [vp, ctypes.c_int]
  lib.pgx_plan_profile_enable.restype = ctypes.c_int
  lib.pgx_plan_profile_read.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_double),
                                        ctypes.POINTER(ctypes.c_char_p)]
  lib.pgx_plan_profile_read.restype = ctypes.c_int
  lib.pgx_bp_run_vjp.argtypes = [vp, vp, i64, vp, ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, vp, i32, f32, f32,
                                 vp, vp, vp]
  lib.pgx_bp_run_vjp.restype = ctypes.c_int
  lib.pgx_nccl_unique_id.argtypes = [vp]
  lib.pgx_nccl_unique_id.restype = ctypes.c_int
  lib.pgx_strip_create.argtypes = [i64, i64, ctypes.c_int, ctypes.c_int, vp, ctypes.POINTER(vp)]
  lib.pgx_strip_create.restype = ctypes.c_int
  lib.pgx_strip_destroy.argtypes = [vp]
  lib.pgx_strip_destroy.restype = None
  lib.pgx_strip_run.argtypes = [vp, vp, vp, vp, vp, vp, i32, f32, f32, ctypes.c_uint32]
  lib.pgx_strip_run.restype = ctypes.c_int
  lib.pgx_strip_beliefs.argtypes = [vp, vp, vp, vp, vp]
  lib.pgx_strip_beliefs.restype = ctypes.c_int
  lib.pgx_strip_launch_count.argtypes = [vp]
  lib.pgx_strip_launch_count.restype = ctypes.c_int64
  lib.pgx_strip_graph_launch_count.argtypes = [vp]
  lib.pgx_strip_graph_launch_count.restype = ctypes.c_int64
  lib.pgx_last_error.argtypes = []
  lib.pgx_last_error.restype = ctypes.c_char_p
  lib.pgx_build_info.argtypes = []
  lib.pgx_build_info.restype = ctypes.c_char_p
  _lib = lib
  return lib


def check(code: int) -> None:
  if code != PGX_OK:
    raise PgxError(code, load().pgx_last_error().decode())


def _i32(arr) -> np.ndarray:
  arr = np.asarray(arr)
  if arr.size and (arr.max() >= 2**31 or arr.min() < -(2**31)):
    raise ValueError("index does not fit int32")
  return np.ascontiguousarray(arr, dtype=np.int32)


def _ptr(arr: np.ndarray):
  return arr.ctypes.data_as(_i32p)


def _logical_desc(keep: list, wiring, parents, children, offset: int, msg_start: int) -> LogicalDescC:
  """LogicalWiring / PoolWiring arrays (local indices) -> pgx_logical_desc (global)."""
  d = LogicalDescC()
  d.num_factors = int(children.shape[0])
  d.num_parents = int(parents.shape[0])
  d.edge_states_offset = int(offset)
  if d.num_factors:
    pf, pm, cm = _i32(parents[:, 0]), _i32(parents[:, 1] + msg_start), _i32(children + msg_start)
    keep.extend([pf, pm, cm])
    d.parents_factor, d.parents_msg, d.children_msg = _ptr(pf), _ptr(pm), _ptr(cm)
  return d


@dataclasses.dataclass
class FlatEnumBlock:
  """One pgx_enum_block in host arrays (see include/pgx.h)."""

  num_factors: int
  factor_configs: np.ndarray  # [num_configs, arity]
  first_edge: int
  first_potential: int


@dataclasses.dataclass
class FlatLogical:
  """One pgx_logical_desc in host arrays: the OR, AND or Pool factors of a graph
  (pgmax/factor/logical.py:264-291, pool.py:168-180).  Message indices are GLOBAL."""

  parents_factor: np.ndarray  # [P] factor of every parent / pool choice, ascending
  parents_msg: np.ndarray     # [P] message index of the wiring's state (state 0; state 1 for AND)
  children_msg: np.ndarray    # [F] same for the child / pool indicator

  @property
  def num_factors(self) -> int:
    return int(np.asarray(self.children_msg).shape[0])


@dataclasses.dataclass
class FlatGraph:
  """A compiled factor graph as the flat arrays of pgx_graph_desc,
  for graphs generated or partitioned directly in array form (no per-factor Python objects).
  Edges are ordered by factor type (Enum, OR, AND, Pool), as in the reference."""

  var_num_states: np.ndarray   # [num_vars]
  edge_var_start: np.ndarray   # [num_edges] var-state index of the edge's state 0
  edge_num_states: np.ndarray  # [num_edges]
  num_potentials: int
  enum_blocks: List[FlatEnumBlock]
  or_factors: Optional[FlatLogical] = None
  and_factors: Optional[FlatLogical] = None
  pool_factors: Optional[FlatLogical] = None


class Plan:
  """Owns a pgx_plan (device index structures) built from a FactorGraphState or a FlatGraph."""

  def __init__(self, fg_state):
    lib = load()
    keep = []  # numpy arrays that must outlive pgx_plan_create
    desc = GraphDescC()
    if isinstance(fg_state, FlatGraph):
      self._fill_from_flat(desc, keep, fg_state)
    else:
      self._fill_from_state(desc, keep, fg_state)
    self._create(lib, desc)

  @staticmethod
  def _fill_from_flat(desc, keep, flat: FlatGraph):
    var_states = _i32(flat.var_num_states)
    edge_var_start, edge_num_states = _i32(flat.edge_var_start), _i32(flat.edge_num_states)
    keep.extend([var_states, edge_var_start, edge_num_states])
    desc.num_vars = int(var_states.shape[0])
    desc.var_num_states = _ptr(var_states)
    desc.num_edges = int(edge_var_start.shape[0])
    desc.edge_var_start = _ptr(edge_var_start)
    desc.edge_num_states = _ptr(edge_num_states)
    desc.num_potentials = int(flat.num_potentials)
    blocks = (EnumBlockC * max(len(flat.enum_blocks), 1))()
    for i, b in enumerate(flat.enum_blocks):
      cfg = _i32(b.factor_configs)
      keep.append(cfg)
      blocks[i].num_factors = int(b.num_factors)
      blocks[i].arity = int(cfg.shape[1])
      blocks[i].num_configs = int(cfg.shape[0])
      blocks[i].configs = _ptr(cfg)
      blocks[i].first_edge = int(b.first_edge)
      # message offset of the block's first edge
      blocks[i].first_msg = int(edge_num_states[: b.first_edge].sum(dtype=np.int64))
      blocks[i].first_potential = int(b.first_potential)
    keep.append(blocks)
    desc.num_enum_blocks = len(flat.enum_blocks)
    desc.enum_blocks = blocks
    for name in ("or_factors", "and_factors", "pool_factors"):
      d = LogicalDescC()
      d.edge_states_offset = -1 if name == "and_factors" else 1
      lg = getattr(flat, name)
      if lg is not None and lg.num_factors:
        pf, pm, cm = _i32(lg.parents_factor), _i32(lg.parents_msg), _i32(lg.children_msg)
        keep.extend([pf, pm, cm])
        d.num_factors, d.num_parents = int(cm.shape[0]), int(pf.shape[0])
        d.parents_factor, d.parents_msg, d.children_msg = _ptr(pf), _ptr(pm), _ptr(cm)
      setattr(desc, name, d)

  @staticmethod
  def _fill_from_state(desc, keep, fg_state):
    # pylint: disable=g-import-not-at-top
    from pgmax_b200 import factor

    var_states = _i32(
        np.concatenate(
            [vg.num_states.reshape(-1) for vg in fg_state.variable_groups]
            + [np.empty((0,), dtype=np.int64)]
        )
    )
    keep.append(var_states)
    desc.num_vars = int(var_states.shape[0])
    desc.var_num_states = _ptr(var_states)

    wirings = [fg_state.wiring[ft] for ft in factor.FACTOR_TYPES]
    edge_var_start = _i32(np.concatenate([w.edge_var_start for w in wirings]))
    edge_num_states = _i32(np.concatenate([w.edge_num_states for w in wirings]))
    keep.extend([edge_var_start, edge_num_states])
    desc.num_edges = int(edge_var_start.shape[0])
    desc.edge_var_start = _ptr(edge_var_start)
    desc.edge_num_states = _ptr(edge_num_states)
    desc.num_potentials = int(fg_state.log_potentials.shape[0])
    edge_msg_start = np.cumsum(edge_num_states, dtype=np.int64) - edge_num_states

    # Enum blocks: local (edge, config) offsets -> global.  The Enum slice comes
    # first in every vector, so its local edge / message offsets are global.
    enum_w = fg_state.wiring[factor.EnumFactor]
    pot_start = fg_state.factor_type_to_potentials_range[factor.EnumFactor][0]
    blocks = (EnumBlockC * max(len(enum_w.blocks), 1))()
    for i, b in enumerate(enum_w.blocks):
      cfg = _i32(b.factor_configs)
      keep.append(cfg)
      blocks[i].num_factors = b.num_factors
      blocks[i].arity = b.arity
      blocks[i].num_configs = b.num_configs
      blocks[i].configs = _ptr(cfg)
      blocks[i].first_edge = b.first_edge
      blocks[i].first_msg = int(edge_msg_start[b.first_edge])
      blocks[i].first_potential = pot_start + b.first_config
    keep.append(blocks)
    desc.num_enum_blocks = len(enum_w.blocks)
    desc.enum_blocks = blocks

    ranges = fg_state.factor_type_to_msgs_range
    w = fg_state.wiring[factor.ORFactor]
    desc.or_factors = _logical_desc(
        keep, w, w.parents_edge_states, w.children_edge_states, 1, ranges[factor.ORFactor][0]
    )
    w = fg_state.wiring[factor.ANDFactor]
    desc.and_factors = _logical_desc(
        keep, w, w.parents_edge_states, w.children_edge_states, -1, ranges[factor.ANDFactor][0]
    )
    w = fg_state.wiring[factor.PoolFactor]
    desc.pool_factors = _logical_desc(
        keep, w, w.pool_choices_edge_states, w.pool_indicators_edge_states, 1,
        ranges[factor.PoolFactor][0],
    )

  def _create(self, lib, desc):
    handle = ctypes.c_void_p()
    check(lib.pgx_plan_create(ctypes.byref(desc), ctypes.byref(handle)))
    self._lib = lib
    self.handle = handle
    info = PlanInfoC()
    check(lib.pgx_plan_get_info(handle, ctypes.byref(info)))
    self.info = info
    self.num_vars = int(info.num_vars)
    self.num_var_states = int(info.num_var_states)
    self.num_edge_states = int(info.num_edge_states)
    self.num_potentials = int(info.num_potentials)
    self.device = int(info.device)

  def __del__(self):
    handle = getattr(self, "handle", None)
    if handle:
      self._lib.pgx_plan_destroy(handle)
      self.handle = None

  @property
  def launch_count(self) -> int:
    return int(self._lib.pgx_plan_launch_count(self.handle))

  @property
  def has_fused_blocks(self) -> bool:
    """True when the plan found dense-grid pairwise blocks (single-pass path available)."""
    return bool(self._lib.pgx_plan_num_fused_blocks(self.handle))

  @property
  def compressed_edges(self) -> int:
    """Edges per sample the single-pass path stores as one float (binary-difference storage)."""
    return int(self._lib.pgx_plan_compressed_edges(self.handle))

  PATH_LATTICE, PATH_RESIDENT, PATH_PULL, PATH_MERGED_MAX, PATH_LOGICAL_PULL = 1, 2, 4, 8, 16
  PATH_LATTICE_STREAM, PATH_AUX_STREAM, PATH_WIDE_SPLIT, PATH_LOGICAL_BIN = 32, 64, 128, 256
  PATH_PERM_POTENTIALS, PATH_HALF_BATCH, PATH_STAGED_WIRING = 512, 1024, 2048
  PATH_TAIL_SPLIT, PATH_LATTICE_BIN, PATH_ORAND_FUSED, PATH_VARSUM_COOP = 4096, 8192, 16384, 32768
  PATH_ENUM_UNARY, PATH_ENUM_CONFIG_MAJOR, PATH_GENERIC_BIN, PATH_ENUM_DENSE_PAIR = 65536, 131072, 262144, 524288
  PATH_ENUM_PAIR_FEW = 1048576

  def disable_paths(self, mask: int) -> None:
    """Pin the launch path (PGX_PATH_* bits of include/pgx.h); all paths are bit-identical."""
    check(self._lib.pgx_plan_disable_paths(self.handle, int(mask)))

  @property
  def dominant_edge_states(self) -> int:
    return int(self._lib.pgx_plan_dominant_edge_states(self.handle))

  @property
  def dominant_grid(self) -> int:
    """CTAs of the dominant kernel's most recent launch."""
    return int(self._lib.pgx_plan_dominant_grid(self.handle))

  @property
  def is_lattice(self) -> bool:
    return bool(self._lib.pgx_plan_is_lattice(self.handle))

  def enable_graphs(self, enabled: bool) -> None:
    """CUDA-graph replay of repeated pgx_bp_run signatures (on by default)."""
    check(self._lib.pgx_plan_enable_graphs(self.handle, int(enabled)))

  @property
  def graph_launch_count(self) -> int:
    return int(self._lib.pgx_plan_graph_launch_count(self.handle))

  def set_exact_order(self, enabled: bool) -> None:
    """Force the two-pass, serial-summation-order path (see pgx_plan_set_exact_order)."""
    check(self._lib.pgx_plan_set_exact_order(self.handle, int(enabled)))

  def profile_enable(self, enabled: bool) -> None:
    check(self._lib.pgx_plan_profile_enable(self.handle, int(enabled)))

  def profile_read(self):
    """(launches, total_ms, kernel_name) of the dominant kernel since the last read."""
    n, ms, name = ctypes.c_int64(), ctypes.c_double(), ctypes.c_char_p()
    check(self._lib.pgx_plan_profile_read(self.handle, ctypes.byref(n), ctypes.byref(ms),
                                          ctypes.byref(name)))
    return int(n.value), float(ms.value), (name.value or b"").decode()

  # The methods below take raw device pointers (ints) so that any owner of device
  # memory (torch tensors here) can call them.
  RUN_INPUT_NORMALIZED, RUN_NO_GRAPH, RUN_POTENTIALS_UNCHANGED, RUN_FINAL_SUMS, RUN_SKIP_OUTPUT = 1, 2, 4, 8, 16

  def bp_run(self, stream: int, batch: int, lp: int, lp_batched: bool, ev: int, ev_batched: bool,
             msgs_in: Optional[int], msgs_batched: bool, msgs_out: int, deltas: Optional[int],
             num_iters: int, damping: float, temperature: float, flags: int = 0) -> None:
    check(self._lib.pgx_bp_run_flags(self.handle, stream, batch, lp, int(lp_batched), ev, int(ev_batched),
                                     msgs_in, int(msgs_batched), msgs_out, deltas, num_iters,
                                     damping, temperature, int(flags)))

  def bp_run_vjp(self, stream: int, batch: int, lp: int, lp_batched: bool, ev: int, ev_batched: bool,
                 msgs_in: Optional[int], msgs_batched: bool, g_out: int, num_iters: int, damping: float,
                 temperature: float, g_lp: Optional[int], g_ev: Optional[int], g_msgs_in: Optional[int]) -> None:
    """Vector-Jacobian product of bp_run at the cotangent g_out of the final messages."""
    check(self._lib.pgx_bp_run_vjp(self.handle, stream, batch, lp, int(lp_batched), ev, int(ev_batched), msgs_in,
                                   int(msgs_batched), g_out, num_iters, damping, temperature, g_lp, g_ev, g_msgs_in))

  def bp_step(self, stream: int, lp: int, ev: int, msgs_in: int, msgs_out: int, damping: float,
              temperature: float, num_iters: int = 1) -> None:
    """One-sample iterations on messages that are already normalised (the output of an
    earlier run): read in place, no staging copy (PGX_RUN_INPUT_NORMALIZED)."""
    check(self._lib.pgx_bp_run_flags(self.handle, stream, 1, lp, 0, ev, 0, msgs_in, 0, msgs_out,
                                     None, num_iters, damping, temperature, 1))

  def beliefs(self, stream: int, batch: int, ev: int, ev_batched: bool, msgs: int,
              msgs_batched: bool, out: int) -> None:
    check(self._lib.pgx_beliefs(self.handle, stream, batch, ev, int(ev_batched), msgs,
                                int(msgs_batched), out))

  def decode(self, stream: int, batch: int, ev: int, ev_batched: bool, msgs: int,
             msgs_batched: bool, map_out: Optional[int], marginals: Optional[int],
             ties: Optional[int]) -> None:
    check(self._lib.pgx_decode(self.handle, stream, batch, ev, int(ev_batched), msgs,
                               int(msgs_batched), map_out, marginals, ties))

  def energy(self, stream: int, batch: int, lp: int, lp_batched: bool, ev: int, ev_batched: bool,
             map_states: int, map_batched: bool, out: int) -> None:
    check(self._lib.pgx_energy(self.handle, stream, batch, lp, int(lp_batched), ev, int(ev_batched),
                               map_states, int(map_batched), out))

  def set_factors(self, factor_edge_start) -> None:
    """Factor membership of the edges ([num_factors + 1] offsets), needed by the sdlp_* calls."""
    starts = _i32(factor_edge_start)
    check(self._lib.pgx_plan_set_factors(self.handle, int(starts.shape[0]) - 1, _ptr(starts)))

  def sdlp_objval_and_grad(self, stream: int, batch: int, lp: int, lp_batched: bool, ev: int,
                           ev_batched: bool, msgs: int, msgs_batched: bool, logsumexp_temp: float,
                           objval: int, grad: Optional[int], bp_updates: Optional[int],
                           edge_vals: Optional[int]) -> None:
    check(self._lib.pgx_sdlp_objval_and_grad(self.handle, stream, batch, lp, int(lp_batched), ev,
                                             int(ev_batched), msgs, int(msgs_batched),
                                             logsumexp_temp, objval, grad, bp_updates, edge_vals))

  def sdlp_run(self, stream: int, batch: int, lp: int, lp_batched: bool, ev: int, ev_batched: bool,
               msgs_in: Optional[int], msgs_batched: bool, msgs_out: int, objvals: Optional[int],
               steps: np.ndarray, momenta: np.ndarray, logsumexp_temp: float) -> None:
    steps = np.ascontiguousarray(steps, dtype=np.float32)
    momenta = np.ascontiguousarray(momenta, dtype=np.float32)
    check(self._lib.pgx_sdlp_run(self.handle, stream, batch, lp, int(lp_batched), ev,
                                 int(ev_batched), msgs_in, int(msgs_batched), msgs_out, objvals,
                                 int(steps.shape[0]), steps.ctypes.data_as(_f32p),
                                 momenta.ctypes.data_as(_f32p), logsumexp_temp))

  def infer_host(self, stream: int, batch: int, lp: int, lp_batched: bool, ev: int,
                 ev_batched: bool, msgs_in: Optional[int], msgs_batched: bool, num_iters: int,
                 damping: float, temperature: float, map_out: Optional[int],
                 marginals: Optional[int], ties: Optional[int], msgs_out: Optional[int],
                 deltas: Optional[int]) -> None:
    check(self._lib.pgx_infer_host(self.handle, stream, batch, lp, int(lp_batched), ev,
                                   int(ev_batched), msgs_in, int(msgs_batched), num_iters,
                                   damping, temperature, map_out, marginals, ties, msgs_out,
                                   deltas))


NCCL_ID_BYTES = 128
STRIP_NO_GRAPH, STRIP_NO_OVERLAP = 1, 2


def nccl_unique_id() -> bytes:
  """A fresh NCCL unique id (call on ONE rank and broadcast the bytes)."""
  buf = ctypes.create_string_buffer(NCCL_ID_BYTES)
  check(load().pgx_nccl_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
  return buf.raw


class Strip:
  """Owns a pgx_strip: one rank's row strip of a 2-D lattice (include/pgx.h, pgx_strip_*)."""

  def __init__(self, n_cols: int, rows: int, rank: int = 0, world: int = 1, nccl_id: Optional[bytes] = None):
    lib = load()
    if world > 1 and (nccl_id is None or len(nccl_id) != NCCL_ID_BYTES):
      raise ValueError(f"a strip of a world of {world} needs the {NCCL_ID_BYTES}-byte NCCL unique id")
    handle = ctypes.c_void_p()
    id_buf = ctypes.create_string_buffer(nccl_id, NCCL_ID_BYTES) if nccl_id is not None else None
    check(lib.pgx_strip_create(int(n_cols), int(rows), int(rank), int(world),
                               ctypes.cast(id_buf, ctypes.c_void_p) if id_buf is not None else None,
                               ctypes.byref(handle)))
    self._lib, self.handle = lib, handle
    self.n_cols, self.rows, self.rank, self.world = int(n_cols), int(rows), int(rank), int(world)

  def __del__(self):
    handle = getattr(self, "handle", None)
    if handle:
      self._lib.pgx_strip_destroy(handle)
      self.handle = None

  def run(self, stream: int, lp: int, ev_own: int, msgs_in: Optional[int], msgs_out: int, num_iters: int,
          damping: float, temperature: float, flags: int = 0) -> None:
    check(self._lib.pgx_strip_run(self.handle, stream, lp, ev_own, msgs_in, msgs_out, int(num_iters),
                                  float(damping), float(temperature), int(flags)))

  def beliefs(self, stream: int, ev_own: int, msgs: int, out: int) -> None:
    check(self._lib.pgx_strip_beliefs(self.handle, stream, ev_own, msgs, out))

  @property
  def launch_count(self) -> int:
    return int(self._lib.pgx_strip_launch_count(self.handle))

  @property
  def graph_launch_count(self) -> int:
    return int(self._lib.pgx_strip_graph_launch_count(self.handle))
