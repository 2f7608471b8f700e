set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputests.log
for s in 1 2; do timeout 200 python tools/exp_streams.py dense $s 30; FUVS_EXP_NOCOUNTS=1 timeout 200 python tools/exp_streams.py dense $s 30; done 2>&1 | grep streams
timeout 900 python bench.py > gpurun_out/bench_verify.json 2> gpurun_out/bench_verify.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_verify.json
