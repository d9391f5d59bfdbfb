"""GPU tests of the array-form graph route (FlatGraph -> pgx_plan) and of the row-strip
runner: single GPU against the oracle, and 2 ranks over NCCL when 2 GPUs are visible."""

import os
import socket

import numpy as np
import pytest
import torch

from oracle import bp_oracle
from pgmax_b200 import dist as pdist

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("temperature", [0.0, 1.0])
def test_flat_ising_single_gpu_vs_oracle(temperature):
  n, iters = 24, 30
  evidence = np.random.default_rng(0).gumbel(size=(n, n, 2)).astype(np.float32)
  strip = pdist.ising_strip(n)
  runner = pdist.StripRunner(strip, pdist.PgxStepEngine(strip.flat, "cuda:0"), "cuda:0")
  msgs, ev = runner.run(evidence.reshape(-1), iters, 0.5, temperature)
  graph = bp_oracle.graph_from_flat(strip.flat)
  want, _ = bp_oracle.run_bp(graph, strip.log_potentials, np.zeros(strip.num_msgs, np.float32),
                             evidence.reshape(-1), iters, 0.5, temperature)
  if temperature == 0.0:
    np.testing.assert_array_equal(msgs.cpu().numpy(), want)
  np.testing.assert_allclose(msgs.cpu().numpy(), want, atol=1e-5)
  beliefs = runner.beliefs(ev, msgs).cpu().numpy()
  np.testing.assert_allclose(beliefs, bp_oracle.flat_beliefs(graph, want, evidence.reshape(-1)), atol=2e-5)


@pytest.mark.parametrize("temperature", [0.0, 1.0])
@pytest.mark.parametrize("n", [24, 300, 1024])
def test_native_strip_single_gpu_vs_oracle(n, temperature):
  """pgx_strip_* with one rank = the whole torus on binary-difference storage (k_lattice_bin:
  partial tiles at n = 24 / 300, 1024 tiles at n = 1024), all iterations one CUDA graph launch.
  Max-product is bit-exact with the oracle (same additions in the same order); a replayed graph
  and the directly enqueued launches give identical bits; resume == one run."""
  from pgmax_b200 import _native
  iters = 12 if n < 1024 else 4
  evidence = torch.from_numpy(np.random.default_rng(0).gumbel(size=(n, n, 2)).astype(np.float32)).cuda().reshape(-1)
  runner = pdist.NativeStripRunner(n, device="cuda:0")
  msgs, ev = runner.run(evidence, iters, 0.5, temperature)
  got = msgs.cpu().numpy()
  whole = pdist.ising_strip(n)
  graph = bp_oracle.graph_from_flat(whole.flat)
  want, _ = bp_oracle.run_bp(graph, whole.log_potentials, np.zeros(whole.num_msgs, np.float32),
                             evidence.cpu().numpy(), iters, 0.5, temperature)
  if temperature == 0.0:
    np.testing.assert_array_equal(got, want)
  np.testing.assert_allclose(got, want, atol=1e-5)
  assert runner.strip.graph_launch_count == 1
  again = runner.run(evidence, iters, 0.5, temperature)[0].cpu().numpy()     # replay
  np.testing.assert_array_equal(again, got)
  assert runner.strip.graph_launch_count == 2
  eager = torch.empty_like(msgs)
  runner.run(evidence, iters, 0.5, temperature, out=eager, flags=_native.STRIP_NO_GRAPH)
  np.testing.assert_array_equal(eager.cpu().numpy(), got)
  half = torch.empty_like(msgs)
  runner.run(evidence, iters // 2, 0.5, temperature, out=half)
  runner.run(evidence, iters - iters // 2, 0.5, temperature, msgs=half, out=half)
  np.testing.assert_array_equal(half.cpu().numpy(), got)
  beliefs = runner.beliefs(ev, msgs).cpu().numpy()
  np.testing.assert_allclose(beliefs, bp_oracle.flat_beliefs(graph, want, evidence.cpu().numpy()), atol=2e-5)


@pytest.mark.parametrize("temperature", [0.0, 1.0])
def test_lattice_binary_difference_storage_is_bit_identical(temperature):
  """pgx_bp_run on a 1024 x 1024 torus (one sample): the large-lattice path keeps one float per
  edge between iterations (k_lattice_bin); same bits as the reference layout (k_lattice_stream),
  with un-normalised initial messages and deltas."""
  n, iters = 1024, 5
  rng = np.random.default_rng(1)
  strip = pdist.ising_strip(n)
  from pgmax_b200 import _native
  plan = _native.Plan(strip.flat)
  assert plan.is_lattice
  dev = torch.device("cuda:0")
  lp = torch.from_numpy((strip.log_potentials + 0.1 * rng.normal(size=strip.log_potentials.shape)).astype(np.float32)).to(dev)
  ev = torch.from_numpy(rng.gumbel(size=2 * n * n).astype(np.float32)).to(dev)
  m0 = torch.from_numpy(rng.normal(size=strip.num_msgs).astype(np.float32)).to(dev)
  stream = torch.cuda.current_stream(dev).cuda_stream
  outs = []
  for mask in (0, plan.PATH_LATTICE_BIN):
    plan.disable_paths(mask)
    out = torch.empty_like(m0)
    deltas = torch.empty(iters, dtype=torch.float32, device=dev)
    plan.bp_run(stream, 1, lp.data_ptr(), False, ev.data_ptr(), False, m0.data_ptr(), False, out.data_ptr(),
                deltas.data_ptr(), iters, 0.5, temperature)
    torch.cuda.synchronize()
    outs.append((out.cpu().numpy(), deltas.cpu().numpy()))
  np.testing.assert_array_equal(outs[0][0], outs[1][0])
  np.testing.assert_array_equal(outs[0][1], outs[1][1])
  assert np.all(outs[0][0].reshape(-1, 2).max(axis=1) == 0.0)


def _native_worker(rank, world, port, n, temperature, iters, out):
  import torch.distributed as dist
  from pgmax_b200 import _native
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.cuda.set_device(rank)
  dev = torch.device("cuda", rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
  try:
    evidence = torch.from_numpy(np.random.default_rng(0).gumbel(size=(n, n, 2)).astype(np.float32)).to(dev)
    runner = pdist.NativeStripRunner(n, rank, world, dev)
    ev_own = evidence[runner.row0 : runner.row0 + runner.rows].reshape(-1).contiguous()
    msgs, _ = runner.run(ev_own, iters, 0.5, temperature)
    got = msgs.cpu().numpy()
    # the same run without the CUDA graph and without the overlap: identical bits
    eager = torch.empty_like(msgs)
    runner.run(ev_own, iters, 0.5, temperature, out=eager, flags=_native.STRIP_NO_GRAPH | _native.STRIP_NO_OVERLAP)
    same = bool(np.array_equal(eager.cpu().numpy(), got))
    beliefs = runner.beliefs(ev_own, msgs).cpu().numpy()
    # N = 1 on this rank's own GPU: the whole torus through the same kernels
    single = pdist.NativeStripRunner(n, 0, 1, dev)
    whole, ev_all = single.run(evidence.reshape(-1), iters, 0.5, temperature)
    whole_b = single.beliefs(ev_all, whole).cpu().numpy()
    lo, hi = runner.global_msg_range
    want = whole.cpu().numpy()[lo:hi]
    out[rank] = (float(np.max(np.abs(got - want))), same, float(np.max(np.abs(beliefs - whole_b[lo // 4 : hi // 4]))))
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,temperature,iters", [(2, 32, 0.0, 20), (2, 600, 1.0, 20), (2, 1024, 1.0, 50),
                                                       (4, 1024, 1.0, 50), (8, 1024, 1.0, 50), (8, 1024, 0.0, 50)])
def test_native_row_strips_match_one_gpu(world, n, temperature, iters):
  """N ranks (NCCL halo ring, interior rows overlapped with the exchange, one CUDA graph per run)
  against N = 1 on the same kernels: BIT-IDENTICAL messages and beliefs at every horizon - the halo
  terms travel unsummed and every boundary variable's sum is formed in the single graph's order
  (first versions exchanged partial sums and drifted by 1e-6 .. 2e-5 over 50 iterations);
  graph + overlap == eager + no overlap bit for bit."""
  if torch.cuda.device_count() < world:
    pytest.skip(f"needs {world} GPUs")
  import torch.multiprocessing as mp
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
  out = mp.Manager().dict()
  mp.spawn(_native_worker, args=(world, port, n, temperature, iters, out), nprocs=world, join=True)
  assert sorted(out.keys()) == list(range(world))
  for rank in range(world):
    err, same, err_b = out[rank]
    assert err == 0.0 and same and err_b == 0.0, (rank, err, same, err_b)


def _worker(rank, world, port, n, temperature, iters, out):
  import torch.distributed as dist
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  try:
    evidence = np.random.default_rng(0).gumbel(size=(n, n, 2)).astype(np.float32)
    strip = pdist.ising_strip(n, rank, world)
    dev = f"cuda:{rank}"
    runner = pdist.StripRunner(strip, pdist.PgxStepEngine(strip.flat, dev), dev)
    msgs, _ = runner.run(evidence[strip.row0 : strip.row0 + strip.rows].reshape(-1), iters, 0.5, temperature)
    whole = pdist.ising_strip(n)
    graph = bp_oracle.graph_from_flat(whole.flat)
    want, _ = bp_oracle.run_bp(graph, whole.log_potentials, np.zeros(whole.num_msgs, np.float32),
                               evidence.reshape(-1), iters, 0.5, temperature)
    lo, hi = strip.global_msg_range
    out[rank] = float(np.max(np.abs(msgs.cpu().numpy() - want[lo:hi])))
  finally:
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("temperature", [0.0, 1.0])
def test_row_strips_two_gpus_match_single_graph(temperature):
  import torch.multiprocessing as mp
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
  out = mp.Manager().dict()
  mp.spawn(_worker, args=(2, port, 32, temperature, 20, out), nprocs=2, join=True)
  assert max(out.values()) <= 1e-5, dict(out)
