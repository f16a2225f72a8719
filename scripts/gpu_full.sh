# the whole GPU suite, then the bench lines (FP64 generic with all side legs, FP64 known, F32 generic / known)
O=gpurun_out
T=${1:-r02}
python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; tail -4 $O/${T}_tests.log
python bench.py --steps 2 --warmup 3 > $O/${T}_bench_generic_f64.json 2> $O/${T}_bench_generic_f64.err
for c in "known f64" "generic f32" "known f32"; do set -- $c
  python bench.py --model $1 --dtype $2 --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > $O/${T}_bench_$1_$2.json 2> $O/${T}_bench_$1_$2.err
done
for f in generic_f64 known_f64 generic_f32 known_f32; do tail -c 300 $O/${T}_bench_$f.err; python -c "
import json
d=json.loads(open('$O/${T}_bench_$f.json').read().strip().splitlines()[-1])
print('$f', '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'ms %.1f'%d['ms_per_step'], 'frac %.4f'%d['roofline']['frac'], d['resampling'], d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('single_core_value'))
"; done
