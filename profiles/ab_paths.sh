#!/usr/bin/env bash
# A/B of launch paths on the GPU box: bash profiles/ab_paths.sh <workload> mask1 mask2 ...
w=$1; shift
for m in "$@"; do
  echo "== $w --disable-paths $m"
  timeout 300 python bench.py --no-cpu-baseline --workload $w --disable-paths $m 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.4g e2e %.4g kernel %s kernel_ms %.4f iter_ms %.4f iter_frac %.3f checksum %r sm_mhz %s %s' % (d['value'], d['e2e']['value'], r['kernel'], r['kernel_ms'], r['iter_ms'], r['iter_frac'], d['checksum_max_abs_msg'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"
done
