# developer loop for the dense strip kernel: parity tests, then us per interval with 1 and 2 streams
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/dense_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/dense_tests.log
for m in dense dense_smooth block; do for s in 1 2; do timeout 200 python tools/exp_streams.py $m $s 30; done; done 2>&1 | grep streams
