#!/bin/bash
# ncu captures of the dynamics kernel (with Jacobians: launch 2; state only: launch 7) at K = 10^5
set -x
python scripts/bench_dynamics.py 100000 > gpurun_out/dyn_plain.json 2> gpurun_out/dyn_plain.err || exit 1
cat gpurun_out/dyn_plain.json
for w in jac:1 state:6; do
  n=${w%%:*}; s=${w##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dynamics_bulk -s $s -c 1 -f -o gpurun_out/dyn_$n python scripts/bench_dynamics.py 100000 > gpurun_out/dyn_ncu_$n.log 2>&1
  ncu -i gpurun_out/dyn_$n.ncu-rep --page raw --csv > gpurun_out/dyn_${n}_raw.csv 2>/dev/null
  ncu -i gpurun_out/dyn_$n.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/dyn_${n}_source.csv 2>/dev/null
done
ls -la gpurun_out/dyn_*
