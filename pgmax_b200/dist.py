"""Partitioning of the BP path across the GPUs of one box.

One process per GPU (torch.distributed; NCCL over NVLink on the box, gloo in the
CPU tests).  Two modes, as the north-star names them:

* independent problems of a batch (RBM samples, deconvolution images) are split by
  batch index: ``shard_bounds`` / ``shard_batch``.  Every rank runs the identical
  kernels on its shard; there is NO data-path collective (``all_gather_batch`` only
  collects results afterwards);
* one giant Ising grid is split into row strips: ``ising_strip`` builds the local
  graph of a rank directly in array form (no per-factor Python objects: the 8192^2
  torus has 134 M factors) and ``StripRunner`` runs BP with ONE halo exchange per
  iteration;
* any other single graph (EnumFactors, OR, AND, Pool) is cut by factors: ``partition_flat`` /
  ``PartitionRunner`` (end of this file), one all-reduce of the shared variables'
  partial sums per iteration.

Halo exchange (torus of n x n variables, variable (i, j) owns the vertical factor
[(i, j), (i+1, j)] and the horizontal factor [(i, j), (i, j+1)], as
examples/ising_model.ipynb cell 12).  Rank g owns rows [r0, r1) and their factors and
keeps a ghost copy of row r1 (owned by rank g+1), the second variable of its last
row's vertical factors.  Both halo quantities depend only on the messages of the
previous iteration, so one simultaneous exchange at the top of every iteration
suffices:

  g -> g+1: the boundary factors' factor->variable messages into row r1 (2n floats);
            g+1 adds them to the evidence of its first row;
  g+1 -> g: for its first row, ev_v + (messages from its OWN factors) (2n floats):
            the evidence of g's ghost copies, so that the local variable sum of a
            ghost is the full S_v and q = S_v - m is the boundary variable->factor
            message.

That is the protocol as the host-side StripRunner states it (and as the gloo tests run it
on CPU): the boundary rows are summed from partial sums, so its results agree with the
single graph to fp32 rounding.  The product path, NativeStripRunner / pgx_strip_*, sends
the g+1 -> g terms UNSUMMED and forms every boundary sum in the single graph's order:
N strips are bit-identical to one graph (csrc/pgx_strip.cuh).
"""

import dataclasses
from typing import Optional, Tuple

import numpy as np

from pgmax_b200 import _native


# ----------------------------------------------------------------------------------------
# batch sharding
# ----------------------------------------------------------------------------------------
def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
  """Contiguous, balanced split of range(n): the first n % world ranks get one extra."""
  if not 0 <= rank < world:
    raise ValueError(f"rank {rank} outside world of size {world}")
  base, extra = divmod(n, world)
  start = rank * base + min(rank, extra)
  return start, start + base + (1 if rank < extra else 0)


def shard_batch(bp_arrays, world: int, rank: int):
  """This rank's slice of the batch axis of every batched field of a BPArrays;
  fields without a batch axis (shared potentials) pass through."""
  from pgmax_b200.infer.bp_state import BPArrays  # pylint: disable=g-import-not-at-top

  batch = bp_arrays.batch_size
  if batch is None:
    raise ValueError("BPArrays has no batch axis to shard")
  lo, hi = shard_bounds(batch, world, rank)
  pick = lambda a: a[lo:hi] if a.ndim == 2 else a
  return BPArrays(log_potentials=pick(bp_arrays.log_potentials),
                  ftov_msgs=pick(bp_arrays.ftov_msgs), evidence=pick(bp_arrays.evidence))


def all_gather_batch(local, total: int, group=None):
  """Concatenates per-rank results along the batch axis on every rank (result
  collection only; shards may differ in size by one)."""
  import torch  # pylint: disable=g-import-not-at-top
  import torch.distributed as dist  # pylint: disable=g-import-not-at-top

  world = dist.get_world_size(group)
  t = local if torch.is_tensor(local) else torch.from_numpy(np.ascontiguousarray(local))
  sizes = [shard_bounds(total, world, r) for r in range(world)]
  width = max(hi - lo for lo, hi in sizes)
  padded = torch.zeros((width,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
  padded[: t.shape[0]] = t
  out = [torch.empty_like(padded) for _ in range(world)]
  dist.all_gather(out, padded, group=group)
  return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


# ----------------------------------------------------------------------------------------
# row strips of an Ising torus
# ----------------------------------------------------------------------------------------
@dataclasses.dataclass
class IsingStrip:
  """Local graph of one rank + the index arrays of its halo exchange."""

  n: int
  rank: int
  world: int
  row0: int
  rows: int                       # owned rows
  flat: _native.FlatGraph         # (rows [+1 ghost]) x n binary variables, 2 * rows * n factors
  log_potentials: np.ndarray      # [4 * num_factors] fp32
  send_down_msg: np.ndarray       # [n, 2] messages of the last row's vertical factors into the ghost row
  first_row_vs: np.ndarray        # [n, 2] var-states of the first owned row
  first_row_own_msgs: np.ndarray  # [n, 3, 2] messages into the first row from this rank's factors, ascending
  ghost_vs: np.ndarray            # [n, 2] var-states of the ghost row

  @property
  def num_msgs(self) -> int:
    return 8 * self.rows * self.n

  @property
  def num_var_states(self) -> int:
    return int(self.flat.var_num_states.shape[0]) * 2

  @property
  def global_msg_range(self) -> Tuple[int, int]:
    """The local message vector is this slice of the single-graph message vector."""
    return 8 * self.row0 * self.n, 8 * (self.row0 + self.rows) * self.n


def ising_strip(n: int, rank: int = 0, world: int = 1, coupling: float = 0.8) -> IsingStrip:
  """Row strip `rank` of `world` of the n x n Ising torus (world == 1: the whole torus,
  identical to the graph examples/ising_model.ipynb builds).  Vectorised."""
  row0, row1 = shard_bounds(n, world, rank)
  rows = row1 - row0
  if world > 1 and rows < 2:
    raise ValueError("every strip needs at least 2 rows")
  local_rows = rows + (1 if world > 1 else 0)
  num_vars = local_rows * n
  ll, jj = np.meshgrid(np.arange(rows), np.arange(n), indexing="ij")
  v0 = ll * n + jj
  below = (ll + 1) if world > 1 else (ll + 1) % n
  v_vert = below * n + jj
  v_horz = ll * n + (jj + 1) % n
  # factor 2*(l*n+j)+t, t = 0 vertical, 1 horizontal; edges (factor, var0), (factor, var1)
  edge_var = np.stack([v0, v_vert, v0, v_horz], axis=-1).reshape(-1)
  num_factors = 2 * rows * n
  flat = _native.FlatGraph(
      var_num_states=np.full((num_vars,), 2, dtype=np.int32),
      edge_var_start=(2 * edge_var).astype(np.int32),
      edge_num_states=np.full((2 * num_factors,), 2, dtype=np.int32),
      num_potentials=4 * num_factors,
      enum_blocks=[_native.FlatEnumBlock(
          num_factors=num_factors,
          factor_configs=np.array([[0, 0], [0, 1], [1, 0], [1, 1]], dtype=np.int32),
          first_edge=0, first_potential=0)],
  )
  lp = np.tile(np.float32(coupling) * np.array([1.0, -1.0, -1.0, 1.0], dtype=np.float32), num_factors)
  cols = np.arange(n)
  s = np.arange(2)
  fvert = lambda l, j: 2 * (l * n + j)
  fhorz = lambda l, j: 2 * (l * n + j) + 1
  send_down = (4 * fvert(rows - 1, cols) + 2)[:, None] + s[None]
  first_vs = (2 * cols)[:, None] + s[None]
  own = np.stack([4 * fvert(0, cols), 4 * fhorz(0, cols), 4 * fhorz(0, (cols - 1) % n) + 2], axis=1)
  own = np.sort(own, axis=1)[:, :, None] + s[None, None]
  ghost = (2 * (rows * n + cols))[:, None] + s[None]
  return IsingStrip(n=n, rank=rank, world=world, row0=row0, rows=rows, flat=flat, log_potentials=lp,
                    send_down_msg=send_down.astype(np.int64), first_row_vs=first_vs.astype(np.int64),
                    first_row_own_msgs=own.astype(np.int64), ghost_vs=ghost.astype(np.int64))


class PgxStepEngine:
  """One BP iteration of a FlatGraph on this rank's GPU through the C ABI."""

  def __init__(self, flat: _native.FlatGraph, device):
    import torch  # pylint: disable=g-import-not-at-top

    self.torch = torch
    self.device = torch.device(device)
    with torch.cuda.device(self.device):
      self.plan = _native.Plan(flat)

  def step(self, lp, ev, msgs_in, msgs_out, damping: float, temperature: float) -> None:
    stream = self.torch.cuda.current_stream(self.device).cuda_stream
    self.plan.bp_step(stream, lp.data_ptr(), ev.data_ptr(), msgs_in.data_ptr(), msgs_out.data_ptr(),
                      damping, temperature)

  def beliefs(self, ev, msgs):
    out = self.torch.empty_like(ev)
    stream = self.torch.cuda.current_stream(self.device).cuda_stream
    self.plan.beliefs(stream, 1, ev.data_ptr(), False, msgs.data_ptr(), False, out.data_ptr())
    return out


class NativeStripRunner:
  """Row strip `rank` of `world` of the n x n Ising torus on this rank's GPU, through
  pgx_strip_* (include/pgx.h): the local iteration, the halo pack, the NCCL send/recv ring on
  a side stream and the overlap with the interior rows all live in the library, and a run of
  `num_iters` iterations is ONE CUDA graph launch.  torch.distributed is used once, to
  broadcast the NCCL unique id of the library's own communicator.

  The strip's arrays never exist on the host: potentials are generated on the device.
  """

  def __init__(self, n: int, rank: int = 0, world: int = 1, device="cuda:0", group=None, coupling: float = 0.8):
    import torch  # pylint: disable=g-import-not-at-top

    self.torch = torch
    self.n, self.rank, self.world = int(n), int(rank), int(world)
    self.device = torch.device(device)
    self.row0, row1 = shard_bounds(n, world, rank)
    self.rows = row1 - self.row0
    if self.rows < 2:
      raise ValueError("every strip needs at least 2 rows")
    nccl_id = None
    if world > 1:
      import torch.distributed as dist  # pylint: disable=g-import-not-at-top

      buf = torch.zeros(_native.NCCL_ID_BYTES, dtype=torch.uint8, device=self.device)
      if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(_native.nccl_unique_id()), dtype=torch.uint8))
      dist.broadcast(buf, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
      nccl_id = bytes(buf.cpu().numpy().tobytes())
    with torch.cuda.device(self.device):
      self.strip = _native.Strip(n, self.rows, rank, world, nccl_id)
      pattern = torch.tensor([1.0, -1.0, -1.0, 1.0], dtype=torch.float32, device=self.device) * float(coupling)
      self.lp = pattern.repeat(2 * self.rows * self.n)
      self._out = torch.empty(self.num_msgs, dtype=torch.float32, device=self.device)

  @property
  def num_msgs(self) -> int:
    return 8 * self.rows * self.n

  @property
  def global_msg_range(self) -> Tuple[int, int]:
    return 8 * self.row0 * self.n, 8 * (self.row0 + self.rows) * self.n

  def _stream(self) -> int:
    return self.torch.cuda.current_stream(self.device).cuda_stream

  def run(self, evidence_own, num_iters: int, damping: float = 0.5, temperature: float = 0.0, msgs=None,
          out=None, flags: int = 0):
    """evidence_own: [rows * n * 2] device tensor (read in place).  Returns (messages
    [8 * rows * n], evidence_own).  Without `out` the messages land in a buffer the runner
    owns and reuses (the next run overwrites it); the same (evidence, msgs, out) buffers and
    scalars replay the same CUDA graph."""
    torch = self.torch
    ev = torch.as_tensor(evidence_own, dtype=torch.float32, device=self.device).reshape(-1)
    if not ev.is_contiguous():
      ev = ev.contiguous()
    dst = self._out if out is None else out
    self.strip.run(self._stream(), self.lp.data_ptr(), ev.data_ptr(), None if msgs is None else msgs.data_ptr(),
                   dst.data_ptr(), max(int(num_iters), 1), float(damping), float(temperature), flags)
    return dst, ev

  def beliefs(self, ev_own, msgs):
    """Beliefs of the owned variables [rows * n * 2] (one more boundary exchange)."""
    out = self.torch.empty(2 * self.rows * self.n, dtype=self.torch.float32, device=self.device)
    self.strip.beliefs(self._stream(), ev_own.data_ptr(), msgs.data_ptr(), out.data_ptr())
    return out


class StripRunner:
  """Loopy BP on one row strip with a per-iteration halo exchange: the HOST-SIDE statement of
  the protocol (index arrays of ising_strip + torch.distributed p2p), kept as the executable
  specification the gloo tests run on CPU with an oracle-backed engine; on GPUs the product
  path is NativeStripRunner.

  `engine` does the local iteration (PgxStepEngine on a GPU; the tests inject an
  oracle-backed engine on CPU tensors, which is how the N > 1 host logic is covered
  without GPUs).  All tensors live on `device`.
  """

  def __init__(self, strip: IsingStrip, engine, device, group=None):
    import torch  # pylint: disable=g-import-not-at-top

    self.torch = torch
    self.strip, self.engine, self.device, self.group = strip, engine, torch.device(device), group
    to = lambda a: torch.from_numpy(a).to(self.device)
    self.lp = to(strip.log_potentials)
    self.send_down_msg = to(strip.send_down_msg.reshape(-1))
    self.first_row_vs = to(strip.first_row_vs.reshape(-1))
    self.own_msgs = [to(np.ascontiguousarray(strip.first_row_own_msgs[:, k, :]).reshape(-1)) for k in range(3)]
    self.ghost_vs = to(strip.ghost_vs.reshape(-1))
    width = 2 * strip.n
    self.recv_up = torch.zeros(width, dtype=torch.float32, device=self.device)
    self.recv_down = torch.zeros(width, dtype=torch.float32, device=self.device)

  def _exchange(self, ev_own, msgs, ev_local):
    """Fills ev_local (evidence of the local graph for the coming iteration)."""
    torch, strip = self.torch, self.strip
    ev_local[: ev_own.shape[0]] = ev_own
    if strip.world == 1:
      return
    import torch.distributed as dist  # pylint: disable=g-import-not-at-top

    down = (strip.rank + 1) % strip.world
    up = (strip.rank - 1) % strip.world
    send_down = msgs[self.send_down_msg]
    send_up = ev_own[self.first_row_vs]
    for idx in self.own_msgs:  # ascending message index
      send_up = send_up + msgs[idx]
    ops = [dist.P2POp(dist.isend, send_down, down, self.group),
           dist.P2POp(dist.isend, send_up, up, self.group),
           dist.P2POp(dist.irecv, self.recv_up, up, self.group),
           dist.P2POp(dist.irecv, self.recv_down, down, self.group)]
    for req in dist.batch_isend_irecv(ops):
      req.wait()
    ev_local[self.first_row_vs] = ev_own[self.first_row_vs] + self.recv_up
    ev_local[self.ghost_vs] = self.recv_down

  def run(self, evidence_own, num_iters: int, damping: float = 0.5, temperature: float = 0.0,
          msgs=None):
    """evidence_own: [rows * n * 2] evidence of the owned variables.  Returns
    (messages [8 * rows * n], local evidence incl. halo terms of the LAST exchange)."""
    torch, strip = self.torch, self.strip
    ev_own = torch.as_tensor(evidence_own, dtype=torch.float32, device=self.device).reshape(-1)
    ev_local = torch.zeros(strip.num_var_states, dtype=torch.float32, device=self.device)
    cur = torch.zeros(strip.num_msgs, dtype=torch.float32, device=self.device) if msgs is None else msgs
    nxt = torch.empty_like(cur)
    for _ in range(max(int(num_iters), 1)):
      self._exchange(ev_own, cur, ev_local)
      self.engine.step(self.lp, ev_local, cur, nxt, float(damping), float(temperature))
      cur, nxt = nxt, cur
    return cur, ev_own

  def beliefs(self, ev_own, msgs):
    """Beliefs of the owned variables [rows * n * 2] (one more boundary exchange)."""
    ev_local = self.torch.zeros(self.strip.num_var_states, dtype=self.torch.float32, device=self.device)
    self._exchange(ev_own, msgs, ev_local)
    return self.engine.beliefs(ev_local, msgs)[: ev_own.shape[0]]


# ----------------------------------------------------------------------------------------
# general partition of one graph of EnumFactors
# ----------------------------------------------------------------------------------------
def flat_from_state(fg_state) -> _native.FlatGraph:
  """FlatGraph of a compiled FactorGraphState: EnumFactors of any arity (ragged numbers of
  states), OR, AND and Pool factors - the arrays Plan builds its pgx_graph_desc from."""
  from pgmax_b200 import factor  # pylint: disable=g-import-not-at-top

  wirings = [fg_state.wiring[ft] for ft in factor.FACTOR_TYPES]
  var_states = np.concatenate([vg.num_states.reshape(-1) for vg in fg_state.variable_groups]
                              + [np.empty((0,), dtype=np.int64)])
  ranges = fg_state.factor_type_to_msgs_range

  def logical(ft, parents, children):
    if children.shape[0] == 0:
      return None
    start = ranges[ft][0]
    return _native.FlatLogical(parents_factor=np.asarray(parents[:, 0], dtype=np.int64),
                               parents_msg=np.asarray(parents[:, 1], dtype=np.int64) + start,
                               children_msg=np.asarray(children, dtype=np.int64) + start)

  w = fg_state.wiring[factor.EnumFactor]
  w_or, w_and, w_pool = (fg_state.wiring[ft] for ft in (factor.ORFactor, factor.ANDFactor, factor.PoolFactor))
  return _native.FlatGraph(
      var_num_states=var_states.astype(np.int32),
      edge_var_start=np.concatenate([np.asarray(x.edge_var_start, dtype=np.int64) for x in wirings]).astype(np.int32),
      edge_num_states=np.concatenate([np.asarray(x.edge_num_states, dtype=np.int64) for x in wirings]).astype(np.int32),
      num_potentials=int(fg_state.log_potentials.shape[0]),
      enum_blocks=[_native.FlatEnumBlock(num_factors=b.num_factors, factor_configs=np.asarray(b.factor_configs),
                                         first_edge=b.first_edge, first_potential=b.first_config)
                   for b in w.blocks],
      or_factors=logical(factor.ORFactor, w_or.parents_edge_states, w_or.children_edge_states),
      and_factors=logical(factor.ANDFactor, w_and.parents_edge_states, w_and.children_edge_states),
      pool_factors=logical(factor.PoolFactor, w_pool.pool_choices_edge_states, w_pool.pool_indicators_edge_states))


@dataclasses.dataclass
class GraphPart:
  """Local graph of one rank of a factor-partitioned graph + the index arrays that tie it to the
  single graph.  The factors (all blocks, in block order) are cut into contiguous, balanced
  ranges; a rank's local graph holds its factors and every variable they touch.  Variables touched by more than
  one rank are "shared": their variable sums need the other ranks' messages."""

  rank: int
  world: int
  flat: _native.FlatGraph
  potential_index: np.ndarray   # [C_local] global potential of every local potential
  msg_index: np.ndarray         # [E_s local] global message of every local message
  var_state_index: np.ndarray   # [V_s local] global var-state of every local var-state
  shared_local_vs: np.ndarray   # local var-states of the shared variables this rank touches
  shared_slot: np.ndarray       # their slot in the dense exchange vector
  num_shared: int               # length of the exchange vector (var-states of ALL shared variables)


_LOGICAL_FIELDS = (("or_factors", 0), ("and_factors", 1), ("pool_factors", 0))  # (name, state the wiring points at)


def _segments(flat: _native.FlatGraph):
  """The factors of the graph as segments in message order: one per enum block, then the OR, AND
  and Pool factors.  Per segment: (kind, factor count, first-edge offsets [count + 1])."""
  edge_ns = np.asarray(flat.edge_num_states, dtype=np.int64)
  edge_msg_start = np.cumsum(edge_ns) - edge_ns
  out = []
  for b in flat.enum_blocks:
    arity = int(np.asarray(b.factor_configs).shape[1])
    out.append(("enum", int(b.num_factors), b.first_edge + arity * np.arange(b.num_factors + 1, dtype=np.int64), b))
  first_edge = int(sum(s[1] * np.asarray(s[3].factor_configs).shape[1] for s in out))
  for name, rel in _LOGICAL_FIELDS:
    lg = getattr(flat, name, None)
    if lg is None or lg.num_factors == 0:
      continue
    pf = np.asarray(lg.parents_factor, dtype=np.int64)
    counts = np.bincount(pf, minlength=lg.num_factors) + 1  # parents + the child
    starts = first_edge + np.concatenate([[0], np.cumsum(counts)])
    # the reference's layout: a factor's edges are contiguous, parents in order, child last
    parent_edge = np.repeat(starts[:-1], counts - 1) + (np.arange(pf.shape[0]) - np.repeat(np.cumsum(counts - 1) - (counts - 1), counts - 1))
    if (np.any(np.diff(pf) < 0) or np.any(edge_msg_start[parent_edge] + rel != np.asarray(lg.parents_msg))
        or np.any(edge_msg_start[starts[1:] - 1] + rel != np.asarray(lg.children_msg))):
      raise NotImplementedError(f"{name}: the edges of a factor are not contiguous (parents, then the child)")
    out.append((name, lg.num_factors, starts, lg))
    first_edge = int(starts[-1])
  return out


def _segment_ranges(segments, world: int, rank: int):
  """Per segment: (factor lo, factor hi) of this rank.  The factors of all segments, in order,
  are cut into `world` contiguous balanced ranges (graphs made of many small factor groups are
  balanced too); a segment's range is its intersection with the rank's range."""
  counts = [s[1] for s in segments]
  lo, hi = shard_bounds(sum(counts), world, rank)
  out, first = [], 0
  for n in counts:
    out.append((min(max(lo - first, 0), n), min(max(hi - first, 0), n)))
    first += n
  return out


def partition_flat(flat: _native.FlatGraph, world: int, rank: int) -> GraphPart:
  """Rank `rank`'s part of `flat` (vectorised; every rank derives the same shared-variable list)."""
  var_ns = np.asarray(flat.var_num_states, dtype=np.int64)
  var_start = np.concatenate([[0], np.cumsum(var_ns)])
  edge_vs = np.asarray(flat.edge_var_start, dtype=np.int64)
  edge_ns = np.asarray(flat.edge_num_states, dtype=np.int64)
  edge_var = np.searchsorted(var_start, edge_vs, side="right") - 1
  edge_msg_start = np.cumsum(edge_ns) - edge_ns
  segments = _segments(flat)

  def edges_of(r):
    """Global edge ids of rank r, segment by segment (ascending)."""
    parts = [np.arange(seg[2][lo], seg[2][hi], dtype=np.int64)
             for seg, (lo, hi) in zip(segments, _segment_ranges(segments, world, r))]
    return np.concatenate(parts) if parts else np.zeros((0,), dtype=np.int64)

  touched = [np.unique(edge_var[edges_of(r)]) for r in range(world)]
  counts = np.zeros((var_ns.shape[0],), dtype=np.int64)
  for t in touched:
    counts[t] += 1
  shared_vars = np.flatnonzero(counts > 1)
  shared_ns = var_ns[shared_vars]
  shared_first_slot = np.cumsum(shared_ns) - shared_ns

  my_edges = edges_of(rank)
  my_vars = touched[rank]                                  # ascending global ids = local order
  local_of_var = np.full((var_ns.shape[0],), -1, dtype=np.int64)
  local_of_var[my_vars] = np.arange(my_vars.shape[0])
  local_ns = var_ns[my_vars]
  local_var_start = np.cumsum(local_ns) - local_ns
  # state 0 of an edge always sits at its variable's first state in this layout
  local_edge_vs = local_var_start[local_of_var[edge_var[my_edges]]] + (edge_vs[my_edges] - var_start[edge_var[my_edges]])

  def expand(starts, sizes):
    """Concatenation of arange(start, start + size) for every (start, size)."""
    total = int(sizes.sum())
    if total == 0:
      return np.zeros((0,), dtype=np.int64)
    offs = np.arange(total) - np.repeat(np.cumsum(sizes) - sizes, sizes)
    return np.repeat(starts, sizes) + offs

  msg_index = expand(edge_msg_start[my_edges], edge_ns[my_edges])   # ascending
  blocks, pot_parts, logical, first_edge, first_pot = [], [], {}, 0, 0
  for seg, (lo, hi) in zip(segments, _segment_ranges(segments, world, rank)):
    kind, _, starts, obj = seg
    if hi <= lo:
      continue
    if kind == "enum":
      cfg = np.asarray(obj.factor_configs)
      k, arity = int(cfg.shape[0]), int(cfg.shape[1])
      blocks.append(_native.FlatEnumBlock(num_factors=hi - lo, factor_configs=cfg, first_edge=first_edge,
                                          first_potential=first_pot))
      pot_parts.append(np.arange(obj.first_potential + lo * k, obj.first_potential + hi * k, dtype=np.int64))
      first_pot += (hi - lo) * k
    else:
      pf = np.asarray(obj.parents_factor, dtype=np.int64)
      sel = (pf >= lo) & (pf < hi)
      to_local = lambda g: np.searchsorted(msg_index, np.asarray(g, dtype=np.int64))
      logical[kind] = _native.FlatLogical(parents_factor=pf[sel] - lo,
                                          parents_msg=to_local(np.asarray(obj.parents_msg)[sel]),
                                          children_msg=to_local(np.asarray(obj.children_msg)[lo:hi]))
    first_edge += int(starts[hi] - starts[lo])
  potential_index = np.concatenate(pot_parts) if pot_parts else np.zeros((0,), dtype=np.int64)
  var_state_index = expand(var_start[my_vars], local_ns)
  mine_shared = shared_vars[np.isin(shared_vars, my_vars)]
  pos = np.searchsorted(shared_vars, mine_shared)
  shared_local_vs = expand(local_var_start[local_of_var[mine_shared]], var_ns[mine_shared])
  shared_slot = expand(shared_first_slot[pos], shared_ns[pos])
  local_flat = _native.FlatGraph(
      var_num_states=local_ns.astype(np.int32), edge_var_start=local_edge_vs.astype(np.int32),
      edge_num_states=edge_ns[my_edges].astype(np.int32), num_potentials=int(potential_index.shape[0]),
      enum_blocks=blocks, or_factors=logical.get("or_factors"), and_factors=logical.get("and_factors"),
      pool_factors=logical.get("pool_factors"))
  return GraphPart(rank=rank, world=world, flat=local_flat, potential_index=potential_index, msg_index=msg_index,
                   var_state_index=var_state_index, shared_local_vs=shared_local_vs, shared_slot=shared_slot,
                   num_shared=int(shared_ns.sum()))


class PartitionRunner:
  """Loopy BP on one part of a factor-partitioned graph.

  Per iteration ONE collective: every rank adds up, per state of every shared variable, the
  messages of its own factors (the engine's beliefs with zero evidence), the partial sums are
  all-reduced over a dense vector of the shared var-states (NCCL over NVLink on the box; gloo in
  the CPU tests), and a rank's local evidence of a shared variable becomes
  ``ev + (total - own partial)`` - so that the local variable sum is the full S_v and the local
  iteration is exactly the single-GPU kernel sequence on the local graph.  As with the row strips
  the summation order of shared variables differs from the single graph's: results agree to fp32
  rounding, not bit for bit.  `engine` as in StripRunner.
  """

  def __init__(self, part: GraphPart, engine, device, group=None):
    import torch  # pylint: disable=g-import-not-at-top

    self.torch = torch
    self.part, self.engine, self.device, self.group = part, engine, torch.device(device), group
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
    self.shared_local_vs = to(part.shared_local_vs)
    self.shared_slot = to(part.shared_slot)
    self.num_local_vs = int(np.asarray(part.flat.var_num_states, dtype=np.int64).sum())
    self.num_msgs = int(np.asarray(part.flat.edge_num_states, dtype=np.int64).sum())
    self.zero_ev = torch.zeros(self.num_local_vs, dtype=torch.float32, device=self.device)
    self.buf = torch.zeros(max(part.num_shared, 1), dtype=torch.float32, device=self.device)

  def _local_evidence(self, ev, msgs):
    """Evidence of the local graph for the coming iteration (ev: evidence of the local variables)."""
    if self.part.world == 1 or self.part.num_shared == 0:
      return ev
    import torch.distributed as dist  # pylint: disable=g-import-not-at-top

    partial = self.engine.beliefs(self.zero_ev, msgs)[self.shared_local_vs]
    self.buf.zero_()
    self.buf[self.shared_slot] = partial
    dist.all_reduce(self.buf, group=self.group)
    ev_local = ev.clone()
    ev_local[self.shared_local_vs] = ev[self.shared_local_vs] + (self.buf[self.shared_slot] - partial)
    return ev_local

  def run(self, lp_local, ev_local_vars, num_iters: int, damping: float = 0.5, temperature: float = 0.0, msgs=None):
    """lp_local: potentials of the local factors (global[potential_index]); ev_local_vars: evidence
    of the local variables (global[var_state_index]).  Returns the local messages."""
    torch = self.torch
    lp = torch.as_tensor(lp_local, dtype=torch.float32, device=self.device).reshape(-1)
    ev = torch.as_tensor(ev_local_vars, dtype=torch.float32, device=self.device).reshape(-1)
    cur = torch.zeros(self.num_msgs, dtype=torch.float32, device=self.device) if msgs is None else msgs
    nxt = torch.empty_like(cur)
    for _ in range(max(int(num_iters), 1)):
      self.engine.step(lp, self._local_evidence(ev, cur), cur, nxt, float(damping), float(temperature))
      cur, nxt = nxt, cur
    return cur

  def beliefs(self, ev_local_vars, msgs):
    """Beliefs of the local variables (full sums: one more exchange)."""
    ev = self.torch.as_tensor(ev_local_vars, dtype=self.torch.float32, device=self.device).reshape(-1)
    return self.engine.beliefs(self._local_evidence(ev, msgs), msgs)
