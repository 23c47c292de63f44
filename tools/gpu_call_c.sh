#!/bin/bash
# Round-2 GPU call C (2 GPUs): full parity suite (incl. the 2-GPU NCCL / peer-memory test), bench at N=1 and N=2.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/c_smi.csv 2>&1
nvidia-smi topo -m > gpurun_out/c_topo.txt 2>&1
echo "== pytest gpu (2 GPUs visible)"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 40 gpurun_out/c_pytest.log
echo "== bench N=1"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/c_bench_n1.json 2> gpurun_out/c_bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/c_bench_n1.err
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/c_bench_n2.json 2> gpurun_out/c_bench_n2.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/c_bench_n2.err
echo "== launch list ns"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c_launches_ns.csv python tools/profile_step.py ns 2 > gpurun_out/c_l_ns.log 2>&1
echo "== done"
