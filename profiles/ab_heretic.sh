# heretic A/B: k_enum_pair_few (default) / k_enum_pair_dense (1048576) / k_enum_small_cm (524288) / k_enum_small (131072)
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_late_additions.py tests/test_gpu_full_size.py tests/test_gpu_graphs.py -x -q 2>&1 | tail -4
for dp in 0 1048576 524288 131072; do python bench.py --workload heretic --disable-paths $dp --no-cpu-baseline --no-extras --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    rf=r.get('roofline',{})
    print(r.get('ms_per_step'), rf.get('kernel'), rf.get('kernel_ms'), rf.get('iter_ms'), rf.get('frac'), rf.get('iter_frac'), rf.get('physical_frac'), r.get('parity_max_abs'))
"; done
for w in rcn rcn_sum; do python bench.py --workload $w --no-cpu-baseline --no-extras --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    rf=r.get('roofline',{})
    print(r.get('ms_per_step'), rf.get('kernel'), rf.get('kernel_ms'), rf.get('iter_ms'), rf.get('frac'), rf.get('iter_frac'), rf.get('physical_frac'))
"; done
