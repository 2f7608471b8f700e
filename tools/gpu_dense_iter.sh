#!/bin/bash
# one iteration of the dense-kernel loop: parity first, then timings (graph replay, 4 clips x 3 intervals per step)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "dense or full_size or graph or limits" > gpurun_out/t_dense.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_dense.log
tail -3 gpurun_out/t_dense.log
{
timeout 300 python tools/exp_streams.py dense 1
timeout 300 python tools/exp_streams.py dense 2
timeout 300 python tools/exp_streams.py dense_smooth 1
timeout 300 python tools/exp_streams.py dense_smooth 2
} > gpurun_out/d.log 2>&1
cat gpurun_out/d.log
