#!/bin/bash
# Round-2 GPU call P (2 GPUs): the two-GPU parity test, bench at N = 2 (weak scaling, frame gather, sharded config 4).
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest 2-gpu tests"
timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_configs.py -m gpu -q -rs -p no:cacheprovider > gpurun_out/p_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/p_pytest.log
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/p_bench_n2.json 2> gpurun_out/p_bench_n2.err
echo "rc=$?"; tail -c 600 gpurun_out/p_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/p_bench_n2.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value']); print('gather',json.dumps(d['with_frame_gather'])[:1500]); print('c4',json.dumps(d['c4_sharded'])[:800])"
echo "== bench N=1 (same box)"
timeout 900 python bench.py --steps 30 --warmup 5 --skip-others --skip-cpu --skip-host-abi > gpurun_out/p_bench_n1.json 2> gpurun_out/p_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/p_bench_n1.json')); print('value',d['value'],'e2e',d['e2e']['value'],'c4',d['c4_sharded']['frames_per_s'])"
echo "== done"
