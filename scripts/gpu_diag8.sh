#!/bin/bash
# diagnosis of the 8-process slowdown of k_sweep_v2 on some devices
O=gpurun_out
B="python bench.py --steps 1 --warmup 2 --no-secondary --no-cpu-baseline"
val() { python -c "
import json,sys
d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$1', '%.4g'%d['value'], 'ms %.1f'%d['ms_per_step'], [ (p['rank'], p['ms_per_step'][0]) for p in (d.get('per_rank') or []) ])"; }
echo "== A: one process alone on device 0, then device 2"
for dv in 0 2; do CUDA_VISIBLE_DEVICES=$dv $B > $O/diag_alone_$dv.json 2>/dev/null; val $O/diag_alone_$dv.json; done
echo "== B: 8 independent processes at once (no torch.distributed), one per device"
for dv in 0 1 2 3 4 5 6 7; do CUDA_VISIBLE_DEVICES=$dv $B > $O/diag_conc_$dv.json 2>/dev/null & done; wait
for dv in 0 1 2 3 4 5 6 7; do val $O/diag_conc_$dv.json; done
echo "== C: torchrun x 8 with NCCL_NVLS_ENABLE=0"
NCCL_NVLS_ENABLE=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 1 --warmup 2 --no-secondary > $O/diag_nonvls.json 2>/dev/null; val $O/diag_nonvls.json
echo "== D: torchrun x 8 with CUDA_DEVICE_MAX_CONNECTIONS / P2P disabled (NCCL_P2P_DISABLE=1 NCCL_SHM_DISABLE=0)"
NCCL_P2P_DISABLE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 1 --warmup 2 --no-secondary > $O/diag_nop2p.json 2>/dev/null; val $O/diag_nop2p.json
