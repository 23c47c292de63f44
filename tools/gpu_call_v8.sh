#!/bin/bash
# Round-2 call V8 (1 GPU): end-to-end A/B in one box: default library (counters zeroed by a chained kernel) against the
# build with the memset node, through the torch extension (the library file is swapped on the box).
set -u
mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps 50 --warmup 5 --skip-others --skip-cpu --skip-host-abi --skip-c4 2> gpurun_out/v8_$1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'], d['e2e']['modes'])"; }
run default_a
cp pytorch3d_b200/lib/libb200raster.so /tmp/lib_default.so
cp tools/_variants/lib_memset_node.so pytorch3d_b200/lib/libb200raster.so
run memset_node_a
cp /tmp/lib_default.so pytorch3d_b200/lib/libb200raster.so
run default_b
cp tools/_variants/lib_memset_node.so pytorch3d_b200/lib/libb200raster.so
run memset_node_b
cp /tmp/lib_default.so pytorch3d_b200/lib/libb200raster.so
echo "== done"
