#!/bin/bash
# compute-sanitizer evidence pass: memcheck / racecheck / synccheck over smoke() and one 1080p
# interval per mode.  Logs -> gpurun_out/sanitize_*.log (summaries are copied to profiles/ by hand).
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 900 $CS --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_${tool}_smoke.log 2>&1
  echo "$tool smoke rc=$?"; tail -n 3 gpurun_out/sanitize_${tool}_smoke.log
  for m in dense linear block; do
    timeout 900 $CS --tool $tool --print-limit 20 python tools/profile_target.py --mode $m --clips 1 --reps 1 > gpurun_out/sanitize_${tool}_$m.log 2>&1
    echo "$tool $m rc=$?"; tail -n 3 gpurun_out/sanitize_${tool}_$m.log
  done
done
