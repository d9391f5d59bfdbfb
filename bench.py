#!/usr/bin/env python
"""bench.py — BP messages updated per second on the configurations BASELINE.json names.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rbm|ising50|ising_big]
  python bench.py --impl reference ...      # the CPU arm (NumPy oracle on the host cores)

A "step" is one call of the hot path over one batch of synthetic input: bp.run for
the workload's iteration count (pgx_bp_run: all iterations enqueued back to back on
one stream).  Default workload (N=1): BASELINE.json configs[1], the RBM 784 x 500,
sum-product with Gumbel evidence, batch 1024, 200 iterations, damping 0.5.

One JSON line on stdout (rank 0):
  value        messages (edge-states) updated / s, whole job, inputs resident in HBM
  e2e          the same through pgx_infer_host with pinned HOST buffers: H2D copies,
               all iterations, fused decode, D2H of the MAP states inside the timing
  roofline     dominant kernel: algorithmic bytes per launch / its CUDA-event time,
               against MEASURED_PEAKS.json's hbm_gbs (fallback 6650 GB/s)
  cpu_baseline the NumPy oracle on the host cores (bounded sample), rank 0, N=1 only

Multi-GPU (torchrun, one rank per GPU): independent samples are split by batch
index, every rank runs the identical kernels on its own shard, no data-path
collective ("scaling": "weak": per-GPU batch is fixed).  Timing: barrier +
synchronize on both sides, CUDA events, max over ranks.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
  if _p not in sys.path:
    sys.path.insert(0, _p)

import numpy as np  # noqa: E402

# stdout carries the ONE JSON line and nothing else: everything other code (NCCL's version banner,
# library warnings) writes to fd 1 goes to stderr instead.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
  _JSON_OUT.write(json.dumps(line) + "\n")
  _JSON_OUT.flush()


METRIC = "bp_messages_updated_per_sec"
UNIT = "edge_states/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md


# ----------------------------------------------------------------------------------------
# Workloads (synthetic data of the named shapes)
# ----------------------------------------------------------------------------------------
def build_workload(name: str, batch_override=None, iters_override=None):
  """Returns dict(fg groups, evidence (host fp32), temperature, iters, damping, batch, ...)."""
  import models
  from pgmax_b200 import infer

  if name in ("rbm", "rbm_max"):
    # benchmark/rbm_lib.py:138-169 shape; examples/rbm.ipynb sizes; np.random.seed(0) weights.
    # rbm_max: the same run as max-product (T = 0: the reference's own use of this model,
    # benchmark/rbm_lib.py:173-185; BASELINE's metric is quoted on sum-product)
    nh, nv, batch, iters, temperature = 500, 784, 1024, 200, (1.0 if name == "rbm" else 0.0)
    rs = np.random.RandomState(0)
    W, bh, bv = rs.normal(size=(nh, nv)), rs.logistic(size=nh), rs.logistic(size=nv)
    fg, hidden, visible = models.rbm_model(W, bh, bv)
    batch = batch_override or batch
    rng = np.random.default_rng(0)
    bp = infer.BP(fg.bp_state, temperature=temperature)
    evidence = {hidden: rng.gumbel(size=(batch, nh, 2)).astype(np.float32),
                visible: rng.gumbel(size=(batch, nv, 2)).astype(np.float32)}
    label = (f"RBM {nv}x{nh} pairwise EnumFactors, " + ("sum-product T=1" if name == "rbm" else "max-product T=0") +
             ", Gumbel evidence")
  elif name == "rbm_small":  # CPU-sized stand-in used by the tests of bench.py itself
    nh, nv, batch, iters, temperature = 20, 30, 8, 10, 1.0
    rs = np.random.RandomState(0)
    W, bh, bv = rs.normal(size=(nh, nv)), rs.logistic(size=nh), rs.logistic(size=nv)
    fg, hidden, visible = models.rbm_model(W, bh, bv)
    batch = batch_override or batch
    rng = np.random.default_rng(0)
    bp = infer.BP(fg.bp_state, temperature=temperature)
    evidence = {hidden: rng.gumbel(size=(batch, nh, 2)).astype(np.float32),
                visible: rng.gumbel(size=(batch, nv, 2)).astype(np.float32)}
    label = f"RBM {nv}x{nh} (test size)"
  elif name == "ising50":
    # examples/ising_model.ipynb: 50x50 torus, coupling 0.8, 1000 iters, T=0.05
    batch, iters, temperature = batch_override, 1000, 0.05
    fg, variables, ev = models.ising_model(n=50, batch=batch)
    bp = infer.BP(fg.bp_state, temperature=temperature)
    evidence = {variables: ev.astype(np.float32)}
    label = "Ising 50x50 torus, pairwise EnumFactors, T=0.05"
  elif name == "ising50_batch":
    # a NON-special-cased pairwise graph: the same torus at batch 1024 (no lattice / resident path:
    # those are one-sample paths) - generic variable-sum + pairwise kernels, sum-product
    batch, iters, temperature = batch_override or 1024, 200, 1.0
    fg, variables, ev = models.ising_model(n=50, batch=batch)
    bp = infer.BP(fg.bp_state, temperature=temperature)
    evidence = {variables: ev.astype(np.float32)}
    label = "Ising 50x50 torus batched (generic pairwise path), sum-product T=1"
  elif name == "heretic":
    # tests/test_pgmax.py:424-475: 17-state x 3-state pairwise EnumFactors (multi-state enum kernel)
    batch, iters, temperature = batch_override or 256, 100, 1.0
    fg, pixel_vars, hidden_vars = models.heretic_model()
    bp = infer.BP(fg.bp_state, temperature=temperature)
    rng = np.random.default_rng(0)
    evidence = {pixel_vars: rng.gumbel(size=(batch, 30, 30, 3)).astype(np.float32),
                hidden_vars: rng.gumbel(size=(batch, 28, 28, 17)).astype(np.float32)}
    label = "'heretic' model: 7 056 pairwise EnumFactors of 17 x 3 states (generic enum path), sum-product T=1"
  elif name == "deconv":
    # examples/pmp_binary_deconvolution.ipynb shape: one 28x28 image per graph, 5 features 6x6,
    # 100 synthetic images batched, max-product, 100 iterations
    batch, iters, temperature = batch_override or 100, 100, 0.0
    fg, groups = models.deconv_model()
    bp = infer.BP(fg.bp_state, temperature=temperature)
    evidence = {k: v.astype(np.float32) for k, v in models.deconv_evidence(groups, batch).items()}
    label = "binary deconvolution 28x28, 95 220 AND + 784 OR factors, max-product"
  elif name in ("rcn", "rcn_sum"):
    # examples/rcn.ipynb shape: 20 models x 80 variables of 625 states, large-config pairwise
    # EnumFactors (perturb radius 2..8), max-product, 30 iterations; rcn_sum: the same graph,
    # sum-product at T = 1 (not a reference configuration: the T > 0 sibling of the RCN kernel)
    batch, iters, temperature = batch_override, 30, (0.0 if name == "rcn" else 1.0)
    fg, groups, evidence = models.rcn_model()
    bp = infer.BP(fg.bp_state, temperature=temperature)
    evidence = {k: v.astype(np.float32) for k, v in evidence.items()}
    label = ("RCN-shaped: 20 x 80 variables of 625 states, 3 180 large-config EnumFactors, " +
             ("max-product" if name == "rcn" else "sum-product T=1"))
  else:
    raise ValueError(f"unknown workload {name}")
  iters = iters_override or iters
  arrays = bp.init(evidence_updates=evidence)
  return dict(name=name, label=label, bp=bp, arrays=arrays, iters=iters, damping=0.5,
              temperature=temperature, batch=batch or 1)


def strip_record(args, dev, rank, world, n, iters, steps, warmup, sampler_index=None, flags=0):
  """One n x n Ising torus (BASELINE.json configs[4]: n = 8192), sum-product T = 1, damping 0.5,
  split into row strips across `world` ranks through pgx_strip_* (NCCL halo ring per iteration on
  a side stream, interior rows overlapped, all iterations of a run ONE CUDA graph launch).
  Strong scaling: the graph is fixed.  Returns the timing record of this rank (max over ranks is
  taken by the caller)."""
  import torch
  import torch.distributed as dist
  from pgmax_b200 import dist as pdist

  T = 1.0
  runner = pdist.NativeStripRunner(n, rank, world, dev)
  gen = torch.Generator(device=dev).manual_seed(rank)
  u = torch.rand(runner.rows * n * 2, generator=gen, device=dev).clamp_(1e-7, 1 - 1e-7)
  ev_own = -torch.log(-torch.log(u))  # Gumbel(0, 1), generated on the device

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for _ in range(warmup):
    msgs, _ = runner.run(ev_own, iters, 0.5, T, flags=flags)
  barrier()
  launches0 = runner.strip.launch_count
  sampler = ClockSampler(sampler_index) if sampler_index is not None else None
  if sampler:
    sampler.start()
    sampler.wait_first()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  e0.record()
  for _ in range(steps):
    msgs, _ = runner.run(ev_own, iters, 0.5, T, flags=flags)
  e1.record()
  barrier()
  ms = e0.elapsed_time(e1)
  clocks = sampler.stop() if sampler else None
  launches = runner.strip.launch_count - launches0
  # end to end: every step copies the strip's evidence from pinned host memory and reads the
  # step's metric (max |message|) back; device time between barriers
  ev_host = ev_own.cpu().pin_memory()
  ev_dev = torch.empty_like(ev_own)
  for _ in range(min(warmup, 1)):
    runner.run(ev_dev.copy_(ev_host, non_blocking=True), iters, 0.5, T, flags=flags)
  f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  f0.record()
  for _ in range(steps):
    ev_dev.copy_(ev_host, non_blocking=True)
    out, _ = runner.run(ev_dev, iters, 0.5, T, flags=flags)
    metric = float(out.abs().max().item())  # 4-byte device -> host read, synchronises the step
  f1.record()
  barrier()
  e2e_ms = f0.elapsed_time(f1)
  checksum = float(msgs.abs().max().item())
  del runner
  return dict(ms=ms, e2e_ms=e2e_ms, launches=launches, clocks=clocks, checksum=checksum,
              h2d=ev_host.numel() * 4, d2h=4, metric=metric)


def strip_line_fields(n, iters, steps, world, ms, e2e_ms, rec):
  """Metric, roofline and e2e fields of a strip run (ms already max over ranks)."""
  es = 8 * n * n
  bytes_iter = 17 * es   # 4 * (2 + 0.25 + 1 + 1) * E_s, SURVEY.md 8(d)
  moved_iter = 9 * es    # binary-difference storage: 16 + 16 B messages, 32 B potentials, 8 B evidence per cell
  iter_s = ms * 1e-3 / steps / iters
  peak, peak_src = hbm_peak()
  # DRAM bytes ncu measured for one launch over the whole torus (one GPU, same grid): only then
  entry = traffic_entry("ising_big", "k_lattice_bin", 1, None) if (world == 1 and n == 8192) else None
  return {
      "traffic": entry["bytes"] if entry else None,
      "physical_frac": (entry["bytes"] / iter_s / 1e9 / peak) if entry else None,
      "value": es * iters * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "ms_per_step": ms / steps,
      "iter_ms": iter_s * 1e3, "edge_states": es, "iters": iters,
      "frac": bytes_iter / iter_s / 1e9 / world / peak,
      "layout_frac": moved_iter / iter_s / 1e9 / world / peak,
      "algorithmic_bytes_per_iter_per_gpu": bytes_iter // world, "layout_bytes_per_iter_per_gpu": moved_iter // world,
      "peak": peak, "peak_source": peak_src,
      "e2e": {"value": es * iters * steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": rec["h2d"] * world,
              "d2h_bytes_per_step": rec["d2h"] * world, "ms_per_step": e2e_ms / steps},
      "gpu_launches": rec["launches"], "graph_launches_per_step": 1,
      "checksum_max_abs_msg": rec["checksum"],
      "parallelism": (f"row strips x{world}, NCCL send/recv halo ring per iteration on a side stream, interior rows "
                      "overlapped, one CUDA graph per run") if world > 1 else "one GPU, one CUDA graph per run",
  }


def run_ising_big(args):
  """`--workload ising_big`: the strip workload as the main line (strong scaling)."""
  import torch
  import torch.distributed as dist

  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  n, iters = args.size or 8192, args.iters or 200
  rec = strip_record(args, dev, rank, world, n, iters, args.steps, args.warmup, sampler_index=local_rank,
                     flags=args.strip_flags)
  ms, e2e_ms = rec["ms"], rec["e2e_ms"]
  if world > 1:
    t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()
  if rank == 0:
    f = strip_line_fields(n, iters, args.steps, world, ms, e2e_ms, rec)
    line = {
        "metric": METRIC, "value": f["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": f["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Ising {n}x{n} torus, single graph, sum-product T=1", "iters": iters, "damping": 0.5,
                   "temperature": 1.0, "edge_states": f["edge_states"], "parallelism": f["parallelism"],
                   "l2": "working set larger than L2" if 4 * f["edge_states"] > 126e6 else "L2-resident working set"},
        "e2e": f["e2e"], "gpu_launches": f["gpu_launches"], "clocks": rec["clocks"],
        "roofline": {"bound": "hbm", "achieved": f["algorithmic_bytes_per_iter_per_gpu"] / (f["iter_ms"] * 1e-3) / 1e9,
                     "peak": f["peak"], "unit": "GB/s", "frac": f["frac"], "traffic": f["traffic"],
                     "physical_frac": f["physical_frac"],
                     "kernel": "k_lattice_bin (whole iteration incl. the halo exchange when N > 1), per GPU",
                     "algorithmic_bytes_per_launch": f["algorithmic_bytes_per_iter_per_gpu"], "peak_source": f["peak_source"],
                     "iter_ms": f["iter_ms"], "layout_frac": f["layout_frac"],
                     "layout_bytes_per_iter": f["layout_bytes_per_iter_per_gpu"],
                     "storage": "binary-difference, 1 float per two-state edge between iterations"},
        "checksum_max_abs_msg": f["checksum_max_abs_msg"],
    }
    emit(line)
  if world > 1:
    dist.destroy_process_group()


def hbm_peak():
  """(GB/s, source): MEASURED_PEAKS.json's copy bandwidth, else the profiling guide's fallback."""
  try:
    return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
  except (OSError, KeyError, ValueError):
    return FALLBACK_HBM_GBS, "fallback"


def algorithmic_bytes_per_iter(plan, batch, lp_batched):
  """4 * (2*E_s*B + V_s*B + C*B_lp + E_s)."""
  es, vs, c = plan.num_edge_states, plan.num_var_states, plan.num_potentials
  return 4 * (2 * es * batch + vs * batch + c * (batch if lp_batched else 1) + es)


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
class ClockSampler:
  """Samples nvidia-smi SM clocks and throttle reasons while the timed region runs."""

  QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
           "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
           "clocks_event_reasons.sw_power_cap")

  def __init__(self, index: int):
    self.index, self.proc, self.lines = index, None, []

  def wait_first(self, timeout=2.0):
    """Blocks until nvidia-smi has delivered its first sample (its start-up takes a few hundred
    ms), then marks the start of the timed region: stop() reports the samples after the mark."""
    deadline = time.time() + timeout
    while self.proc is not None and not self.lines and time.time() < deadline:
      time.sleep(0.01)
    self.mark = len(self.lines)

  def start(self):
    self.mark = 0
    try:
      self.proc = subprocess.Popen(
          ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
           "--format=csv,noheader,nounits", "-lms", "100"],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    late = not self.lines
    if late:  # timed region shorter than nvidia-smi's start-up + one period: take the first sample now
      deadline = time.time() + 1.5
      while not self.lines and time.time() < deadline:
        time.sleep(0.02)
    self.proc.terminate()
    self.thread.join(timeout=2)
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    lines = self.lines[self.mark:] if len(self.lines) > self.mark else self.lines[-1:]
    for line in lines:
      parts = [p.strip() for p in line.split(",")]
      if len(parts) < 6:
        continue
      try:
        sm.append(float(parts[0])); mx.append(float(parts[1]))
      except ValueError:
        continue
      for n, v in zip(names, parts[2:6]):
        if v.lower().startswith("active"):
          reasons.add(n)
    return {"sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
            "samples": len(sm),
            **({"note": "timed region shorter than one sampling period: sampled right after it"} if late else {})}


# ----------------------------------------------------------------------------------------
# CPU arm: the NumPy oracle on the host cores
# ----------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(name):
  wl = build_workload(name, batch_override=1)
  from oracle import bp_oracle
  _W["graph"] = bp_oracle.graph_from_context(wl["bp"].context)
  _W["wl"] = wl


def _cpu_worker_run(args):
  seed, iters = args
  from oracle import bp_oracle
  wl, graph = _W["wl"], _W["graph"]
  a = wl["arrays"]
  rng = np.random.default_rng(1000 + seed)
  ev = rng.gumbel(size=graph.num_var_states).astype(np.float32)
  msgs, _ = bp_oracle.run_bp(graph, np.asarray(a.log_potentials).reshape(-1)[: a.log_potentials.shape[-1]],
                             np.zeros(a.ftov_msgs.shape[-1], np.float32), ev, iters,
                             wl["damping"], wl["temperature"])
  return float(msgs.sum())


def cpu_arm(name, steps, warmup, iters_per_sample, budget_s=8.0):
  """Oracle throughput with one process per host core; each step = `cores` independent
  samples x `iters_per_sample` iterations of the workload's graph."""
  import multiprocessing as mp
  cores = len(os.sched_getaffinity(0))
  ctx = mp.get_context("spawn")
  with ctx.Pool(cores, initializer=_cpu_worker_init, initargs=(name,)) as pool:
    pool.map(_cpu_worker_run, [(i, 1) for i in range(cores)])  # touch everything once
    t1 = time.perf_counter()
    for _ in range(max(warmup, 1)):
      pool.map(_cpu_worker_run, [(i, 1) for i in range(cores)])
    t1 = (time.perf_counter() - t1) / max(warmup, 1)
    # bounded sample: about `budget_s` seconds of oracle work per step
    iters_per_sample = max(1, min(iters_per_sample, int(budget_s / max(t1, 1e-6))))
    t0 = time.perf_counter()
    for s in range(steps):
      pool.map(_cpu_worker_run, [(s * cores + i, iters_per_sample) for i in range(cores)])
    dt = time.perf_counter() - t0
  wl = build_workload(name)
  es = int(sum(int(w.edge_num_states.sum()) for w in wl["bp"].context.wiring.values()))
  total = es * cores * iters_per_sample * steps
  return dict(value=total / dt, seconds=dt, cores=cores, es=es,
              sample=f"{cores} samples x {iters_per_sample} iterations per step, {steps} steps "
                     f"(one oracle process per core; the workload is {wl['batch']} samples x {wl['iters']} iterations)",
              label=wl["label"], temperature=wl["temperature"], iters=wl["iters"])


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  iters = 3 if args.workload == "rbm" else 50
  res = cpu_arm(args.workload, max(args.steps, 1), min(args.warmup, 1), iters)
  line = {
      "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT,
      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
      "ms_per_step": 1e3 * res["seconds"] / max(args.steps, 1), "higher_is_better": True,
      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      "config": {"workload": res["label"], "iters": res["iters"], "damping": 0.5,
                 "temperature": res["temperature"]},
      "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port",
                       "sample": res["sample"]},
      "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
      "note": "NumPy fp32 restatement of the reference's run_bp (oracle/bp_oracle.py); JAX is not installable here",
  }
  emit(line)


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
def traffic_entry(workload, kernel, batch, grid):
  """DRAM bytes per launch of `kernel` from this round's ncu --set full captures
  (profiles/r02_traffic.json, written by profiles/make_traffic.py from the .ncu-rep files): only
  if the capture was taken on the same kernel, batch and launch grid; otherwise None."""
  try:
    table = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
  except (OSError, ValueError):
    return None
  # (rbm_max launches the same kernel on the same buffers as rbm: the capture of one serves both)
  entry = table.get(f"{workload}:{kernel}") or (table.get(f"rbm:{kernel}") if workload == "rbm_max" else None)
  if not entry or entry.get("batch") != batch or (grid and entry.get("grid") not in (None, grid)):
    return None
  return entry


def oracle_parity(wl, host, dev_run, iters=2, sample=0):
  """Checker, outside every timed region: `iters` iterations of the workload on the device
  (the timed batch, the benchmarked path) against the CPU oracle for ONE sample."""
  from oracle import bp_oracle  # test infrastructure: used here as the checker only
  bp = wl["bp"]
  graph = bp_oracle.graph_from_context(bp.context)
  got = dev_run(iters)
  pick = (lambda a: np.asarray(a)[sample]) if host.batch_size else (lambda a: np.asarray(a))
  row = lambda a: np.asarray(a)[sample] if np.asarray(a).ndim == 2 else np.asarray(a)
  want, _ = bp_oracle.run_bp(graph, row(host.log_potentials), row(host.ftov_msgs), row(host.evidence), iters,
                             wl["damping"], wl["temperature"])
  got = pick(got)
  floor = want <= -1e31
  return {"parity_max_abs": float(np.max(np.abs(got[~floor] - want[~floor]))) if (~floor).any() else 0.0,
          "parity_iters": iters, "parity_sample": sample,
          "parity_floor_mismatch": int(np.sum((got <= -1e31) != floor))}


def measure(args, name, dev, rank, world, local_rank, batch=None, iters=None, steps=None, warmup=None,
            shard_seed=None, with_clocks=True, parity_iters=2, min_seconds=0.0):
  """Times one plan-based workload on this rank: device-resident steps (CUDA events), the
  dominant kernel's launches (events around each one, separate pass), the end-to-end leg through
  pgx_infer_host with pinned host buffers, and the oracle spot-check.  Returns this rank's
  record; the caller reduces times over ranks."""
  import torch
  import torch.distributed as dist
  from pgmax_b200.infer.bp_state import BPArrays

  steps, warmup = steps or args.steps, warmup if warmup is not None else args.warmup
  wl = build_workload(name, batch_override=batch, iters_override=iters)
  bp, iters, damping, T = wl["bp"], wl["iters"], wl["damping"], wl["temperature"]
  host = wl["arrays"]
  if shard_seed is not None and host.evidence.ndim == 2:  # every rank draws its own shard of samples
    rng = np.random.default_rng(shard_seed)
    host = BPArrays(log_potentials=host.log_potentials, ftov_msgs=host.ftov_msgs,
                    evidence=rng.gumbel(size=host.evidence.shape).astype(np.float32))
  plan = bp.context.plan
  plan.set_exact_order(args.exact_order)
  plan.disable_paths(args.disable_paths)
  batch = host.batch_size or 1
  put = lambda a: torch.from_numpy(np.array(a, dtype=np.float32, order="C")).to(dev)
  dev_arrays = BPArrays(log_potentials=put(host.log_potentials), ftov_msgs=put(host.ftov_msgs),
                        evidence=put(host.evidence))
  es = plan.num_edge_states

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  plan.enable_graphs(not args.no_graph)
  # (the previous result is dropped before every run: the allocator then hands the same output
  # buffer back, the call signature repeats and the run is one CUDA graph replay)
  out = None
  for _ in range(max(warmup, 3)):
    out = None
    out = bp.run(dev_arrays, num_iters=iters, damping=damping, temperature=T)
  barrier()
  if min_seconds:  # short workloads: enough steps for >= 3 clock samples (nvidia-smi period 100 ms)
    t0 = time.perf_counter()
    out = None
    out = bp.run(dev_arrays, num_iters=iters, damping=damping, temperature=T)
    torch.cuda.synchronize()
    steps = max(steps, min(2000, int(min_seconds / max(time.perf_counter() - t0, 1e-5)) + 1))
  launches0, graphs0 = plan.launch_count, plan.graph_launch_count
  sampler = ClockSampler(local_rank) if with_clocks else None
  if sampler:
    sampler.start()
    sampler.wait_first()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  e0.record()
  for _ in range(steps):
    out = None
    out = bp.run(dev_arrays, num_iters=iters, damping=damping, temperature=T)
  e1.record()
  barrier()
  ms = e0.elapsed_time(e1)
  clocks = sampler.stop() if sampler else None
  launches, graph_launches = plan.launch_count - launches0, plan.graph_launch_count - graphs0
  # roofline pass: the same step once more with CUDA events around every launch of the dominant
  # kernel (kept out of the timed region above)
  plan.profile_enable(True)
  bp.run(dev_arrays, num_iters=iters, damping=damping, temperature=T)
  torch.cuda.synchronize()
  n_prof, prof_ms, prof_name = plan.profile_read()
  prof_grid = plan.dominant_grid  # CTAs of the launches just timed (later runs may use another launch shape)
  plan.profile_enable(False)
  checksum = float(out.ftov_msgs.float().abs().max().item())

  # end to end through the C ABI with host buffers
  pin = lambda a: torch.from_numpy(np.array(a, dtype=np.float32, order="C")).pin_memory()
  h_lp, h_ev = pin(host.log_potentials), pin(host.evidence)
  h_map = torch.empty((batch, plan.num_vars), dtype=torch.int32).pin_memory()
  h_ties = torch.empty((batch,), dtype=torch.int32).pin_memory()
  stream = torch.cuda.current_stream(dev).cuda_stream

  def e2e_step():
    plan.infer_host(stream, batch, h_lp.data_ptr(), h_lp.ndim == 2, h_ev.data_ptr(), h_ev.ndim == 2,
                    None, False, iters, damping, T, h_map.data_ptr(), None, h_ties.data_ptr(), None, None)

  for _ in range(min(warmup, 2)):
    e2e_step()
  barrier()
  e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e2e_steps = min(steps, 50)
  t0 = time.perf_counter()
  e2.record()
  for _ in range(e2e_steps):
    e2e_step()
  e3.record()
  barrier()
  e2e_ms = max(e2.elapsed_time(e3), 1e3 * (time.perf_counter() - t0)) * steps / e2e_steps  # scaled to `steps`
  parity = None
  if rank == 0 and parity_iters:
    try:
      parity = oracle_parity(wl, host, lambda k: bp.run(dev_arrays, num_iters=k, damping=damping, temperature=T)
                             .ftov_msgs.cpu().numpy(), iters=parity_iters)
    except Exception as err:  # pylint: disable=broad-except
      parity = {"parity_max_abs": None, "parity_note": "check failed: " + repr(err)[-200:]}
  lp_batched = host.log_potentials.ndim == 2
  fused_run = plan.has_fused_blocks and not args.exact_order and batch > 16 and not lp_batched
  return dict(wl=wl, plan=plan, ms=ms, e2e_ms=e2e_ms, launches=launches, graph_launches=graph_launches, clocks=clocks,
              prof_grid=prof_grid, n_prof=n_prof, prof_ms=prof_ms,
              prof_name=prof_name, checksum=checksum, batch=batch, iters=iters, es=es, steps=steps, warmup=warmup,
              h2d=4 * (h_lp.numel() + h_ev.numel()), d2h=4 * (h_map.numel() + h_ties.numel()), parity=parity,
              lp_batched=lp_batched, fused_run=fused_run, damping=damping, temperature=T)


def roofline_of(name, rec, ms):
  """The roofline object of a measured workload (ms: max over ranks of the timed region)."""
  plan, batch, iters, es, steps = rec["plan"], rec["batch"], rec["iters"], rec["es"], rec["steps"]
  peak, peak_src = hbm_peak()
  bytes_iter = algorithmic_bytes_per_iter(plan, batch, rec["lp_batched"])
  n_prof, prof_name = rec["n_prof"], rec["prof_name"]
  kernel_ms = rec["prof_ms"] / max(n_prof, 1)
  # the dominant launch updates dom_es of the es edge-states: its share of the iteration's
  # algorithmic bytes (single-kernel iterations: all of them)
  single_kernel = prof_name in ("k_enum_pw2_bip", "k_lattice", "k_lattice_stream", "k_lattice_bin", "k_enum_pw2_pull",
                                "k_or_and_fused")
  dom_es = es if single_kernel else min(plan.dominant_edge_states, es)
  kernel_bytes = bytes_iter * dom_es // es
  achieved = kernel_bytes / (kernel_ms * 1e-3) / 1e9 if n_prof else None
  # bytes the kernels really have to move: binary-difference storage keeps ONE float per two-state
  # edge between iterations, i.e. one float less to read and one less to write per such edge,
  # sample and iteration than the reference layout
  if rec["fused_run"]:
    saved = 8 * plan.compressed_edges * batch
  elif prof_name == "k_enum_pw2_bin":
    saved = 8 * (es // 2) * batch   # generic two-pass path on binary-difference storage
  elif prof_name == "k_lattice_bin":
    saved = 8 * (es // 2) + 4 * es  # + the incidence index this kernel does not read
  else:
    saved = 0
  layout_bytes = bytes_iter - saved
  iter_ms = ms / steps / iters
  iter_gbs = bytes_iter / (iter_ms * 1e-3) / 1e9
  entry = traffic_entry(name, prof_name, batch, rec["prof_grid"]) if n_prof else None
  traffic = entry["bytes"] if entry else None
  return {
      "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
      "frac": (achieved / peak) if achieved else None, "traffic": traffic,
      # first-class: the DRAM bytes ncu measured for this kernel (same name, batch and grid) over the
      # launch time measured in THIS run, against the same peak - a true fraction of bandwidth
      "physical_frac": (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if (traffic and n_prof) else None,
      "traffic_source": entry.get("source") if entry else None,
      "kernel": prof_name, "kernel_ms": kernel_ms, "launches_timed": n_prof, "kernel_grid": rec["prof_grid"],
      "algorithmic_bytes_per_launch": kernel_bytes, "peak_source": peak_src,
      "edge_states_per_launch": dom_es * batch,
      "iter_ms": iter_ms, "iter_algorithmic_bytes": bytes_iter, "iter_achieved": iter_gbs, "iter_frac": iter_gbs / peak,
      "storage": ("binary-difference, 1 float per two-state edge between iterations" if saved
                  else "reference layout, 1 float per edge-state"),
      "note": ("frac follows SURVEY 8(d): ALGORITHMIC bytes (two floats read + written per two-state edge) / time / "
               "peak; the kernels keep ONE float per such edge between iterations (bit-identical values), so frac "
               "can exceed 1 - layout_frac is the same on the bytes of the layout actually used, physical_frac on "
               "the DRAM bytes ncu measured") if saved else None,
      "layout_bytes_per_iter": layout_bytes,
      "layout_frac": (layout_bytes * (kernel_bytes / bytes_iter) / (kernel_ms * 1e-3) / 1e9 / peak) if n_prof else None,
      "layout_iter_frac": layout_bytes / (iter_ms * 1e-3) / 1e9 / peak,
  }


def compact_record(name, rec, world=1):
  """Short form of a measured workload for the `other_workloads` / `rbm_strong` sub-records."""
  r = roofline_of(name, rec, rec["ms"])
  msgs = rec["es"] * rec["batch"] * rec["iters"] * rec["steps"] * world
  out = {
      "workload": rec["wl"]["label"], "batch": rec["batch"], "iters": rec["iters"], "temperature": rec["temperature"],
      "value": msgs / (rec["ms"] * 1e-3), "unit": UNIT, "ms_per_step": rec["ms"] / rec["steps"], "iter_ms": r["iter_ms"],
      "e2e_value": msgs / (rec["e2e_ms"] * 1e-3), "gpu_launches": rec["launches"],
      "graph_launches": rec["graph_launches"], "steps": rec["steps"],
      "kernel": r["kernel"], "kernel_ms": r["kernel_ms"], "frac": r["frac"], "iter_frac": r["iter_frac"],
      "layout_frac": r["layout_frac"], "layout_iter_frac": r["layout_iter_frac"], "physical_frac": r["physical_frac"],
      "traffic": r["traffic"], "storage": r["storage"], "clocks": rec["clocks"], "checksum_max_abs_msg": rec["checksum"],
  }
  if rec["parity"]:
    out.update(rec["parity"])
  return out


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=3)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="pgx", choices=["pgx", "reference"])
  ap.add_argument("--workload", default="rbm")
  ap.add_argument("--batch", type=int, default=None, help="per-GPU batch override")
  ap.add_argument("--iters", type=int, default=None, help="BP iterations per step override")
  ap.add_argument("--size", type=int, default=None, help="grid side of the ising_big workload")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-extras", action="store_true",
                  help="skip the sub-records (other_workloads at N = 1; rbm_strong and strips at N > 1)")
  ap.add_argument("--no-graph", action="store_true",
                  help="enqueue every launch directly instead of replaying one CUDA graph per run (A/B)")
  ap.add_argument("--strip-flags", type=int, default=0, help="PGX_STRIP_* flags (A/B: 1 no graph, 2 no overlap)")
  ap.add_argument("--disable-paths", type=int, default=0, help="PGX_PATH_* mask (A/B runs of launch paths)")
  ap.add_argument("--exact-order", action="store_true",
                  help="force the two-pass serial-summation-order path (pgx_plan_set_exact_order)")
  args = ap.parse_args()
  if args.impl == "reference":
    run_reference(args)
    return

  if args.workload == "ising_big":
    run_ising_big(args)
    return

  import torch
  import torch.distributed as dist
  from pgmax_b200 import _native

  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise _native.PgxError(_native.PGX_ERR_NO_DEVICE, "bench.py needs a CUDA device (no CPU fallback)")
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)

  def reduce_max(*vals):
    if world == 1:
      return list(vals)
    t = torch.tensor(list(vals), device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()

  # ---- the main line: weak scaling, per-GPU batch fixed, every rank its own shard of samples -----
  rec = measure(args, args.workload, dev, rank, world, local_rank, batch=args.batch, iters=args.iters,
                shard_seed=(100 + rank) if world > 1 else None)
  ms, e2e_ms = reduce_max(rec["ms"], rec["e2e_ms"])
  wl, plan, batch, iters, es = rec["wl"], rec["plan"], rec["batch"], rec["iters"], rec["es"]
  msgs_per_step = es * batch * iters
  line = None
  if rank == 0:
    line = {
        "metric": METRIC, "value": msgs_per_step * args.steps * world / (ms * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["label"], "batch_per_gpu": batch, "global_batch": batch * world,
                   "iters": iters, "damping": rec["damping"], "temperature": rec["temperature"], "edge_states": es,
                   "l2": "working set (2 x %.2f GB of messages) larger than L2" % (4e-9 * es * batch)
                   if 8 * es * batch > 126e6 else "L2-resident working set (reported, not an HBM figure)",
                   "summation_order": "tiled partial sums (single pass)" if rec["fused_run"] else "serial (two-pass)",
                   "parallelism": f"batch-sharded x{world}, no collective"},
        "e2e": {"value": msgs_per_step * args.steps * world / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": rec["h2d"], "d2h_bytes_per_step": rec["d2h"],
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": rec["launches"],
        "graph_launches": rec["graph_launches"],
        "clocks": rec["clocks"],
        "roofline": roofline_of(args.workload, rec, ms),
        "checksum_max_abs_msg": rec["checksum"],
    }
    if rec["parity"]:
      line.update(rec["parity"])
      line["parity_note"] = ("sample 0 of the timed batch, benchmarked path, vs oracle/bp_oracle.py (NumPy fp32, serial "
                             "order), outside the timed region; the oracle's own distance to an fp64 run of this "
                             "recursion is 1.6e-5 after 2 iterations on the RBM (tests/test_gpu_config1_rbm.py)")
  del rec, plan, wl
  torch.cuda.empty_cache()

  # ---- sub-records ---------------------------------------------------------------------------
  if not args.no_extras and args.workload == "rbm" and args.batch is None and args.iters is None:
    quick = dict(steps=2, warmup=3, min_seconds=0.8)
    if world == 1:
      others = {}
      for name in ("rbm_max", "ising50", "deconv", "rcn", "rcn_sum", "ising50_batch", "heretic"):
        try:
          r = measure(args, name, dev, rank, world, local_rank, parity_iters=0 if name.startswith("rcn") else 2, **quick)
          others[name] = compact_record(name, r)
          del r
        except Exception as err:  # pylint: disable=broad-except
          others[name] = {"error": repr(err)[-300:]}
        torch.cuda.empty_cache()
      try:
        n_big, it_big = 8192, 200
        r = strip_record(args, dev, 0, 1, n_big, it_big, 2, 2, sampler_index=local_rank)
        f = strip_line_fields(n_big, it_big, 2, 1, r["ms"], r["e2e_ms"], r)
        f["clocks"] = r["clocks"]
        f["workload"] = f"Ising {n_big}x{n_big} torus, single graph, sum-product T=1, one GPU"
        others["ising_big"] = f
      except Exception as err:  # pylint: disable=broad-except
        others["ising_big"] = {"error": repr(err)[-300:]}
      line["other_workloads"] = others
    else:
      # (1) BASELINE configs[1] as it is worded: batch 1024 sharded over the GPUs (strong scaling);
      #     T1 = the main line's per-GPU time (every rank ran the full batch of 1024 there)
      try:
        per = (1024 + world - 1) // world
        r = measure(args, "rbm", dev, rank, world, local_rank, batch=per, shard_seed=200 + rank, parity_iters=0,
                    with_clocks=rank == 0, **quick)
        s_ms, s_e2e = reduce_max(r["ms"], r["e2e_ms"])
        if rank == 0:
          r["ms"], r["e2e_ms"] = s_ms, s_e2e
          c = compact_record("rbm", r, world)
          c["global_batch"], c["scaling"] = per * world, "strong"
          c["efficiency_vs_same_session_n1"] = (ms / args.steps) / (world * s_ms / r["steps"]) * (per * world / 1024.0)
          c["parallelism"] = f"batch 1024 sharded x{world} ({per} samples per GPU), no collective"
          line["rbm_strong"] = c
        del r
      except Exception as err:  # pylint: disable=broad-except
        if rank == 0:
          line["rbm_strong"] = {"error": repr(err)[-300:]}
      torch.cuda.empty_cache()
      # (2) BASELINE configs[4]: Ising 8192^2 in row strips with the NCCL halo ring (strong scaling);
      #     T1 = the whole torus on rank 0's GPU in the same session
      try:
        n_big, it_big = 8192, 200
        r = strip_record(args, dev, rank, world, n_big, it_big, 2, 2, sampler_index=local_rank if rank == 0 else None)
        s_ms, s_e2e = reduce_max(r["ms"], r["e2e_ms"])
        one = None
        if rank == 0:
          one = strip_record(args, dev, 0, 1, n_big, it_big, 2, 1)
        dist.barrier()
        if rank == 0:
          f = strip_line_fields(n_big, it_big, 2, world, s_ms, s_e2e, r)
          f["clocks"], f["scaling"] = r["clocks"], "strong"
          f["workload"] = f"Ising {n_big}x{n_big} torus, single graph, sum-product T=1"
          f["n1_iter_ms_same_session"] = one["ms"] / 2 / it_big
          f["efficiency_vs_same_session_n1"] = one["ms"] / (world * s_ms)
          line["strips"] = f
      except Exception as err:  # pylint: disable=broad-except
        if rank == 0:
          line["strips"] = {"error": repr(err)[-300:]}

  if rank == 0:
    if world == 1 and not args.no_cpu_baseline:
      try:
        sub = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference",
                              "--workload", args.workload, "--steps", "2", "--warmup", "1"],
                             capture_output=True, text=True, timeout=240)
        ref = json.loads(sub.stdout.strip().splitlines()[-1])
        line["cpu_baseline"] = ref["cpu_baseline"]
      except (IndexError, ValueError, KeyError, subprocess.TimeoutExpired) as err:
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port",
                                "sample": "failed: " + repr(err)[-300:]}
    emit(line)
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
