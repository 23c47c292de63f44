#!/bin/bash
# Round-2 call V9 (2 GPUs): the two-GPU tests (sharded render + packed peer frame exchange) and a short 2-GPU bench line on
# the final tree.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_peer.py -m gpu -q -rs -p no:cacheprovider > gpurun_out/v9_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v9_pytest_2gpu.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 5 --skip-cpu --skip-host-abi --skip-others > gpurun_out/v9_bench_n2.json 2> gpurun_out/v9_bench_n2.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/v9_bench_n2.json')); print(d['n_gpus'], round(d['value']), d['ms_per_step'], d['e2e']['value'], d.get('value_with_gather'), (d.get('c4_sharded') or {}).get('frames_per_s'))"
echo "== done"
