# developer loop for the dense strip kernel: parity tests, then us per interval with 1 and 2 streams
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_golden_gpu.py tests/test_flowmodel_gpu.py -x -q -m gpu -k "dense or 1080p or graph or golden or flow" > gpurun_out/dense_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/dense_tests.log
for m in dense dense_lowres dense_smooth; do for s in 1 2; do timeout 200 python tools/exp_streams.py $m $s 30; done; done 2>&1 | grep "streams\|Error"
