#!/bin/bash
# Round-2 GPU call E (2 GPUs): exchange context + staged unpack, capacity fix (C5), bench N=1 / N=2.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > gpurun_out/e_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 25 gpurun_out/e_pytest.log
echo "== diag gather"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/diag_gather.py > gpurun_out/e_diag_gather.log 2>&1; tail -n 18 gpurun_out/e_diag_gather.log
echo "== bench N=1"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/e_bench_n1.json 2> gpurun_out/e_bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/e_bench_n1.err
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/e_bench_n2.json 2> gpurun_out/e_bench_n2.err
echo "bench rc=$?"; tail -c 800 gpurun_out/e_bench_n2.err
echo "== ncu c5 + ns"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mesh_fine_smemq" -s 0 -c 1 -o gpurun_out/e_prof_c5 -f python tools/profile_step.py c5 1 > gpurun_out/e_ncu_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mesh_fine_kernel" -s 2 -c 1 -o gpurun_out/e_prof_ns -f python tools/profile_step.py ns 3 > gpurun_out/e_ncu_ns.log 2>&1
echo "== done"
