#!/usr/bin/env bash
# A/B on the GPU box: one CUDA graph replay per run (default) against directly enqueued launches
# (--no-graph), device-resident steps; workloads with short kernels and many iterations gain.
for w in ${@:-ising50_batch heretic rcn deconv}; do
  for flag in "" "--no-graph"; do
    python bench.py --workload $w --no-cpu-baseline --no-extras --steps 5 --warmup 3 $flag 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-14s %-10s value %.4g  ms/step %.3f  iter_ms %.5f  launches %d  graph_launches %d  e2e ms/step %.3f' % ('$w', '$flag' or 'graph', d['value'], d['ms_per_step'], d['roofline']['iter_ms'], d['gpu_launches'], d['graph_launches'], d['e2e']['ms_per_step']))"
  done
done
