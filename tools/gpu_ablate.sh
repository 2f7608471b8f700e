timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_flowmodel_gpu.py tests/test_crop_gpu.py -x -q -m gpu -k "upsample or lowres or feature or flow or crop or predict" > gpurun_out/up_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/up_tests.log
timeout 300 python tools/upsample_bench.py 2>&1 | tail -8
for s in 1 2; do timeout 200 python tools/exp_streams.py dense_lowres $s 30 2>&1 | grep streams; done
