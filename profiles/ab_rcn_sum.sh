# A/B of k_enum_big_sumprod_all builds (trip length x CTAs per SM), RCN graph, T = 1, B = 1;
# first line: the generic kernel (merged launch disabled) for comparison
out=gpurun_out/ab_rcn_sum.txt
run() {
  python bench.py --workload rcn_sum --no-cpu-baseline --no-extras --steps 3 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    rf=r.get('roofline',{})
    print(r.get('ms_per_step'), rf.get('kernel'), rf.get('kernel_ms'), rf.get('iter_ms'), rf.get('frac'), rf.get('iter_frac'), r.get('parity_max_abs_err'))
" >> $out 2>&1
}
echo "== generic (PGX_PATH_MERGED_MAX disabled)" >> $out
run --disable-paths 8
for lib in libpgx.so libpgx_s8_c4.so libpgx_s6_c4.so libpgx_s4_c4.so libpgx_s10_c3.so libpgx_s12_c3.so; do
  echo "== $lib" >> $out
  PGX_LIB=$PWD/pgmax_b200/csrc/$lib run
done
