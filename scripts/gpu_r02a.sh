# round-2 baseline of the chunk-stationary sweep kernel: tests, bench lines, ncu launch list, two ncu --set full captures
# (reports stay on the box: only CSV pages / summaries come back, gpurun_out is capped at 64 MiB)
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02a_tests.log 2>&1
python bench.py --steps 2 --warmup 3 > $O/r02a_bench_generic.json 2> $O/r02a_bench_generic.err
python bench.py --model known --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > $O/r02a_bench_known.json 2> $O/r02a_bench_known.err
python bench.py --T 24 --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > $O/r02a_plain_T24.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02a_launches_generic_T24.csv python bench.py --T 24 --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > $O/r02a_ncu1.log 2>&1
for m in generic known; do
  ncu --set full --clock-control none --import-source on -k regex:k_sweep_v2 -s 3 -c 1 -f -o /tmp/r02a_v2_${m}_T24 python bench.py --model $m --T 24 --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > $O/r02a_ncu_$m.log 2>&1
  ncu -i /tmp/r02a_v2_${m}_T24.ncu-rep --page raw --csv > $O/r02a_v2_${m}_T24_raw.csv 2>/dev/null
  ncu -i /tmp/r02a_v2_${m}_T24.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_$m.csv 2>/dev/null
  python scripts/ncu_lines.py /tmp/src_$m.csv 60 > $O/r02a_v2_${m}_T24_lines.txt 2>&1
  gzip -c /tmp/src_$m.csv > $O/r02a_v2_${m}_T24_source.csv.gz
done
ls -la $O
