#!/bin/bash
# Round-2 GPU call D (2 GPUs): diagnostics of the frame exchange and of the module mismatch; bench N=1; quick parity.
set -u
mkdir -p gpurun_out
echo "== diag modules"
timeout 600 python tools/diag_modules.py > gpurun_out/d_diag_modules.log 2>&1; tail -n 12 gpurun_out/d_diag_modules.log
echo "== diag gather"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/diag_gather.py > gpurun_out/d_diag_gather.log 2>&1; tail -n 22 gpurun_out/d_diag_gather.log
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider -x > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 5 gpurun_out/d_pytest.log
echo "== bench N=1"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/d_bench_n1.json 2> gpurun_out/d_bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/d_bench_n1.err
echo "== ncu ns fine"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mesh_fine_kernel" -s 2 -c 1 -o gpurun_out/d_prof_ns -f python tools/profile_step.py ns 3 > gpurun_out/d_ncu_ns.log 2>&1
echo "== done"
