#!/bin/bash
# Round-2 GPU call (4 GPUs): bench at N = 4 (weak scaling, both gather transports, sharded config 4).
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -8
W=${1:-4}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $W --steps 30 --warmup 5 > gpurun_out/n${W}_bench.json 2> gpurun_out/n${W}_bench.err
echo "rc=$?"; tail -c 500 gpurun_out/n${W}_bench.err; python - <<PY
import json
txt=open('gpurun_out/n${W}_bench.json').read()
lines=[l for l in txt.splitlines() if l.startswith('{')]
if lines:
    d=json.loads(lines[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
    g=d['with_frame_gather'] or {}
    print('gather', {k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','receive_gb_per_s_per_rank','error')}) for k,v in g.items() if k!='what'})
    print('c4', d['c4_sharded'])
PY
echo "== done"
