#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "dense or full_size" > gpurun_out/t_dense.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_dense.log
tail -3 gpurun_out/t_dense.log
{
python tools/strip_prof.py dense 2>&1 | grep "launch%16=[4567] " | cut -c1-150
tools/quick_bench.sh dense
tools/quick_bench.sh dense FUVS_DENSE_KERNEL=plane
tools/quick_bench.sh dense_smooth
tools/quick_bench.sh dense_smooth FUVS_DENSE_KERNEL=plane
} > gpurun_out/d.log 2>&1
cat gpurun_out/d.log
