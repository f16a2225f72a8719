O=gpurun_out
python -m pytest tests/test_gpu_f32.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -30 > $O/r02h_tests.log; cat $O/r02h_tests.log
for m in generic known; do
  python bench.py --model $m --dtype f32 --steps 2 --no-secondary --no-cpu-baseline > $O/r02h_bench_${m}_f32.json 2> $O/r02h_bench_${m}_f32.err; tail -c 300 $O/r02h_bench_${m}_f32.err
  python -c "
import json
d=json.loads(open('$O/r02h_bench_${m}_f32.json').read().strip().splitlines()[-1])
print('$m f32', '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'ms %.1f'%d['ms_per_step'], 'frac %.4f'%d['roofline']['frac'], d['resampling'])
"
done
