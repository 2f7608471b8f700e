#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "dense or full_size or temporal or block" > gpurun_out/t_dense.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_dense.log
tail -3 gpurun_out/t_dense.log
{
tools/quick_bench.sh dense
tools/quick_bench.sh dense_smooth
tools/quick_bench.sh block
tools/quick_bench.sh linear
} > gpurun_out/d.log 2>&1
cat gpurun_out/d.log
