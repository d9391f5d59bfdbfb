#!/usr/bin/env bash
# A/B of the k_lattice_bin tile variants (PGX_LB_VARIANT) on the GPU box: parity test + Ising 8192^2 timing.
for v in ${@:-0 1 2 3 4 5}; do
  r=$(PGX_LB_VARIANT=$v python -m pytest tests/test_gpu_strips.py -q -k "native_strip_single or binary_difference" 2>&1 | tail -1)
  t=$(PGX_LB_VARIANT=$v python bench.py --workload ising_big --steps 2 --warmup 1 --iters 100 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['roofline']['iter_ms'],4), round(d['roofline']['layout_frac'],3))" 2>&1 | tail -1)
  echo "variant $v: $t | $r"
done
