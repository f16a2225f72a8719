# BASELINE.json configs[3] / configs[4] through pgb_multi_* (one process, in-process NCCL) at 1, 2, 4, 8 GPUs of one box,
# preceded by the multi-GPU parity tests on all visible devices
O=gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 > $O/r02_multi_tests_8gpu.log; cat $O/r02_multi_tests_8gpu.log
for n in 1 2 4 8; do
  python bench.py --workload c5 --gpus $n --steps 5 2> $O/r02_c5_${n}gpu.err | grep '^{' > $O/r02_c5_${n}gpu.json
  python bench.py --workload c4 --gpus $n --kept 8 2> $O/r02_c4_${n}gpu.err | grep '^{' > $O/r02_c4_${n}gpu.json
done
tail -c 300 $O/r02_c4_8gpu.err $O/r02_c5_8gpu.err
cat $O/r02_c5_*gpu.json $O/r02_c4_*gpu.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['workload'], d['n_gpus'], '%.4g' % d['value'], 'call ms', d.get('host_call_ms'), 'kernel ms', d.get('rollout_kernel_ms_max_over_devices'), 'gather ms', d.get('nccl_gather_ms'), 's', d.get('seconds'), 'create s', d.get('handle_creation_s'))
"
