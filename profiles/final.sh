#!/usr/bin/env bash
# Runs ON THE GPU BOX (gpurun -- 'bash profiles/final.sh <tag>'): the whole GPU test suite, the default
# bench line (all five configurations + generic-path workloads + the CPU arm), smoke(), and one
# `ncu --set full` capture of the dominant kernel of every configuration (named so that
# profiles/make_traffic.py can turn them into profiles/r02_traffic.json) plus launch lists.
# Ordered by importance: the box time may run out before the last steps.
set -u
tag="${1:-r02_z}"
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*"; }
if [ "${2:-all}" != "captures" ]; then
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee $out/${tag}_pytest_gpu.txt
stamp pytest
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2; stamp smoke
fi
B="--steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-graph"
cap() {  # workload batch kernel-regex launch-skip extra-args
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$3" --launch-skip $4 -c 1 -f \
    -o $out/${tag}_$1_b$2 python bench.py $B --workload $1 ${5:-} > $out/${tag}_$1_ncu.log 2>&1
  stamp "full $1"
}
cap rbm 1024 k_enum_pw2_bip 20 "--disable-paths 1024"  # one launch per iteration over the whole batch, as in bench.py's roofline pass
cap deconv 100 k_or_and_fused 11   # odd index: the full-tile launch (the packed tail launch precedes it in every iteration)
cap rcn 1 k_enum_big_maxprod_all 5
cap ising_big 1 k_lattice_bin 3 "--iters 10 --strip-flags 1"
cap rcn_sum 1 k_enum_big_sumprod_all 5
cap heretic 256 k_enum_pair_few 5 "--iters 10"
cap ising50_batch 1024 k_enum_pw2_bin 5 "--iters 10"
for w in rbm deconv rcn; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 40 -c 120 --csv \
    --log-file $out/${tag}_launches_$w.csv python bench.py $B --workload $w > /dev/null 2>&1; stamp "launches $w"
done
# gpurun copies at most 64 MiB back: the reports are summarised here and only the text travels
python profiles/make_traffic.py $out/${tag}_*_b*.ncu-rep > $out/${tag}_traffic.log 2>&1; cp profiles/r02_traffic.json $out/r02_traffic.json
for w in rbm_b1024 deconv_b100 rcn_b1 ising_big_b1 rcn_sum_b1 heretic_b256 ising50_batch_b1024; do
  python profiles/summarize.py $out/${tag}_$w.ncu-rep > $out/${tag}_$w.txt 2>&1
  python profiles/source_hot.py $out/${tag}_$w.ncu-rep 30 > $out/${tag}_source_$w.txt 2>&1
done
rm -f $out/${tag}_*.ncu-rep
# the bench line last: its physical_frac fields read the traffic table the captures above just wrote
if [ "${2:-all}" != "captures" ]; then
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; stamp "bench"; head -c 400 $out/${tag}_bench.json; echo
fi
ls -la $out | tail -24
