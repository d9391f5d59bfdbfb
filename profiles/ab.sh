#!/usr/bin/env bash
# A/B of library builds on the GPU box: bash profiles/ab.sh <bench args> -- lib1.so lib2.so ...
args=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do args+=("$1"); shift; done
shift
for lib in "$@"; do
  echo "== $lib"
  PGX_LIB="$PWD/$lib" timeout 300 python bench.py --no-cpu-baseline "${args[@]}" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.4g e2e %.4g kernel_ms %.4f iter_ms %.4f checksum %r sm_mhz %s %s' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['iter_ms'], d['checksum_max_abs_msg'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"
done
