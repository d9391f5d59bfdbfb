for lib in pgmax_b200/csrc/variants/*.so; do
  PGX_LIB=$PWD/$lib timeout 200 python bench.py --workload ising_big --steps 1 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$lib', 'iter_ms %.4f frac %.3f' % (r['iter_ms'], r['frac']))"
done
