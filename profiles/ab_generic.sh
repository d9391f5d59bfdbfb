# generic-path A/B (runs on the GPU box): bin-storage test + ising50_batch with and without PATH_GENERIC_BIN
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_zz_late_additions.py -x -q -k "binary_difference" 2>&1 | tail -15
for dp in 0 262144; do python bench.py --workload ising50_batch --disable-paths $dp --no-cpu-baseline --no-extras --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    rf=r.get('roofline',{})
    print(r.get('ms_per_step'), rf.get('kernel'), rf.get('kernel_ms'), rf.get('iter_ms'), rf.get('frac'), rf.get('iter_frac'), rf.get('layout_iter_frac'), r.get('parity_max_abs'))
"; done
