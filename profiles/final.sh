#!/usr/bin/env bash
# Runs ON THE GPU BOX (gpurun -- 'bash profiles/final.sh <tag>'): the whole GPU test suite, the
# bench line of every workload (the default one with its CPU arm), smoke(), and the ncu launch
# lists of the workloads whose launch sequence changed.  Ordered by importance: the box time
# may run out before the last steps.
set -u
tag="${1:-r01_o}"
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*"; }
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee $out/${tag}_pytest_gpu.txt
stamp pytest
timeout 200 python bench.py > $out/${tag}_bench_rbm.json 2> $out/${tag}_bench_rbm.err; stamp "bench rbm"; head -c 600 $out/${tag}_bench_rbm.json; echo
for w in deconv rcn ising50; do
  timeout 200 python bench.py --no-cpu-baseline --workload $w > $out/${tag}_bench_$w.json 2> /dev/null; stamp "bench $w"; head -c 300 $out/${tag}_bench_$w.json; echo
done
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2; stamp smoke
B="--steps 1 --warmup 1 --no-cpu-baseline"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_deconv.csv python bench.py $B --workload deconv > /dev/null 2>&1; stamp "launches deconv"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $out/${tag}_launches_rcn.csv python bench.py $B --workload rcn > /dev/null 2>&1; stamp "launches rcn"
timeout 240 ncu --set full --clock-control none --import-source on --launch-skip 60 -c 14 -f -o $out/${tag}_deconv python bench.py $B --workload deconv > $out/${tag}_full_deconv.log 2>&1; stamp "full deconv"
timeout 300 python bench.py --no-cpu-baseline --workload ising_big > $out/${tag}_bench_ising_big.json 2> /dev/null; stamp "bench ising_big"; head -c 300 $out/${tag}_bench_ising_big.json; echo
ls -la $out | tail -12
