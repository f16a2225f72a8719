# round-2 final check on one B200: whole GPU suite, the four bench lines, the reference arm, ncu launch list and two
# ncu --set full captures of the sweep kernel (FP64 and F32).  Reports stay on the box; CSV pages / summaries come back.
O=gpurun_out
T=${1:-r02_final}
python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; tail -3 $O/${T}_tests.log
python bench.py > $O/${T}_bench_generic_f64.json 2> $O/${T}_bench_generic_f64.err
for c in "known f64" "generic f32" "known f32"; do set -- $c
  python bench.py --model $1 --dtype $2 --no-secondary --no-cpu-baseline > $O/${T}_bench_$1_$2.json 2> $O/${T}_bench_$1_$2.err
done
( time python bench.py --impl reference ) > $O/${T}_bench_reference_arm.json 2> $O/${T}_bench_reference_arm.err
for f in generic_f64 known_f64 generic_f32 known_f32; do tail -c 300 $O/${T}_bench_$f.err; python -c "
import json
d=json.loads(open('$O/${T}_bench_$f.json').read().strip().splitlines()[-1])
print('$f', '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'ms %.1f'%d['ms_per_step'], 'frac %.4f'%d['roofline']['frac'], d['resampling'], d.get('cpu_baseline',{}).get('value'))
"; done
tail -4 $O/${T}_bench_reference_arm.err; head -c 600 $O/${T}_bench_reference_arm.json; echo
B="python bench.py --T 24 --steps 1 --warmup 3 --no-secondary --no-cpu-baseline"
$B > $O/${T}_plain_T24.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_generic_T24.csv $B > $O/${T}_ncu1.log 2>&1
for c in "generic f64" "generic f32"; do set -- $c
  ncu --set full --clock-control none --import-source on -k regex:k_sweep_v2 -s 3 -c 1 -f -o /tmp/${T}_$1_$2 $B --model $1 --dtype $2 > $O/${T}_ncu_$1_$2.log 2>&1
  ncu -i /tmp/${T}_$1_$2.ncu-rep --page raw --csv > $O/${T}_v2_$1_$2_T24_raw.csv 2>/dev/null
  ncu -i /tmp/${T}_$1_$2.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src.csv 2>/dev/null
  python scripts/ncu_lines.py /tmp/src.csv 50 > $O/${T}_v2_$1_$2_T24_lines.txt 2>&1
done
ls -la $O | tail -20
