#!/bin/bash
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
tools/quick_bench.sh dense
tools/quick_bench.sh dense_smooth
for S in 1 2 3; do timeout 300 python tools/exp_streams.py dense $S; done
FUVS_STRIP_PDL=0 timeout 300 python tools/exp_streams.py dense 1
FUVS_STRIP_PDL=0 timeout 300 python tools/exp_streams.py dense 2
FUVS_STRIP_PDL=0 timeout 300 python tools/exp_streams.py dense 3
timeout 300 python tools/exp_streams.py dense_smooth 1
timeout 300 python tools/exp_streams.py dense_smooth 2
FUVS_STRIP_PDL=0 timeout 300 python tools/exp_streams.py dense_smooth 2
FUVS_STRIP_TROWS=2 timeout 300 python tools/exp_streams.py dense 1
FUVS_STRIP_TROWS=2 timeout 300 python tools/exp_streams.py dense 2
FUVS_STRIP_TROWS=2 timeout 300 python tools/exp_streams.py dense_smooth 1
timeout 300 python tools/exp_streams.py block 1
timeout 300 python tools/exp_streams.py block 2
timeout 300 python tools/exp_streams.py linear 1
timeout 300 python tools/exp_streams.py linear 2
} > gpurun_out/exp1.log 2>&1
cat gpurun_out/exp1.log
