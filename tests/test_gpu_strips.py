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
