set -x
bash tools/gpu_launch_lists.sh
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_strip -s 5 -c 3 -o gpurun_out/r02_dense_strip_final2 -f python tools/profile_target.py --mode dense --clips 1 --reps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:temporal_counts -s 1 -c 1 -o gpurun_out/r02_tcounts -f python tools/profile_target.py --mode dense --clips 1 --reps 1 > gpurun_out/ncu_full2.log 2>&1; echo "ncu counts rc=$?"
ls -la gpurun_out/*.ncu-rep
