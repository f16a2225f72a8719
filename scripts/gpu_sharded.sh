# BASELINE.json configs[3] / configs[4] through pgb_multi_* (one process, in-process NCCL) at 1, 2, 4, 8 GPUs of one box
O=gpurun_out
for n in 1 2 4 8; do
  python bench.py --workload c5 --gpus $n --steps 5 > $O/r02_c5_${n}gpu.json 2> $O/r02_c5_${n}gpu.err
  python bench.py --workload c4 --gpus $n --kept 8 > $O/r02_c4_${n}gpu.json 2> $O/r02_c4_${n}gpu.err
done
tail -c 300 $O/r02_c4_8gpu.err $O/r02_c5_8gpu.err
cat $O/r02_c5_*gpu.json $O/r02_c4_*gpu.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['workload'], d['n_gpus'], '%.4g' % d['value'], d.get('host_call_ms'), d.get('rollout_kernel_ms_max_over_devices'), d.get('nccl_gather_ms'), d.get('seconds'), d.get('handle_creation_s'))
"
