# A/B of k_enum_big_maxprod_all builds (trip length x CTAs per SM), RCN B = 1
for lib in libpgx.so libpgx_t12_c5.so libpgx_t16_c4.so libpgx_t10_c5.so libpgx_t12_c6.so libpgx_t8_c6.so libpgx_t16_c5.so; do
  echo "== $lib" >> gpurun_out/ab_rcn.txt
  PGX_LIB=$PWD/pgmax_b200/csrc/$lib timeout 300 python bench.py --workload rcn --no-cpu-baseline --no-extras --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    rf=r.get('roofline',{})
    print(r.get('ms_per_step'), rf.get('kernel_ms'), rf.get('iter_ms'), rf.get('frac'), rf.get('iter_frac'))
" >> gpurun_out/ab_rcn.txt 2>&1
done
