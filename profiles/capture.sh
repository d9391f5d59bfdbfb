#!/usr/bin/env bash
# Runs ON THE GPU BOX (gpurun -- 'bash profiles/capture.sh <tag>'): ncu launch lists
# (gpu__time_duration only) and --set full captures of the dominant kernels of every
# workload.  Numbers printed by bench.py under ncu are never bench values.
set -u
tag="${1:-r01_f}"
out=gpurun_out
mkdir -p $out
B="--steps 1 --warmup 1 --no-cpu-baseline"
list() {  # name, count, bench args...
  local name=$1 count=$2; shift 2
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c "$count" --csv \
    --log-file $out/${tag}_launches_${name}.csv python bench.py $B "$@" > $out/${tag}_launches_${name}.log 2>&1
}
full() {  # name, skip, count, bench args...
  local name=$1 skip=$2 count=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on --launch-skip "$skip" -c "$count" -f \
    -o $out/${tag}_${name} python bench.py $B "$@" > $out/${tag}_full_${name}.log 2>&1
}
list rbm 700
list deconv 300 --workload deconv
list rcn 100 --workload rcn
list ising_big 60 --workload ising_big --iters 10
list ising50 40 --workload ising50
full deconv 40 8 --workload deconv
full rcn 10 4 --workload rcn
full ising_big 12 1 --workload ising_big --iters 10
ls -la $out
