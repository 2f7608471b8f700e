#!/bin/bash
# Evidence pass (one gpurun call): plain run, launch list with DRAM bytes, one full ncu capture per mode.
mkdir -p gpurun_out
R='regex:dense_step|dense_strip|temporal_counts|linear_blend|linear_lowres|block_|argmax'
for m in dense linear block linear_lowres; do
  python tools/profile_target.py --mode $m --clips 1 --reps 2 > gpurun_out/plain_$m.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$R" -s 0 -c 40 \
      --csv --log-file gpurun_out/launches_$m.csv python tools/profile_target.py --mode $m --clips 1 --reps 2 > gpurun_out/ncu_$m.log 2>&1
done
python tools/profile_target.py --mode dense --clips 1 --reps 1 > gpurun_out/plain_dense_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dense_strip -s 4 -c 4 -o gpurun_out/prof_dense_strip_final -f \
    python tools/profile_target.py --mode dense --clips 1 --reps 1 > gpurun_out/ncu_dense_full.log 2>&1
python tools/profile_target.py --mode dense_smooth --clips 1 --reps 1 > gpurun_out/plain_smooth_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dense_strip -s 4 -c 4 -o gpurun_out/prof_dense_strip_smooth_final -f \
    python tools/profile_target.py --mode dense_smooth --clips 1 --reps 1 > gpurun_out/ncu_smooth_full.log 2>&1
python tools/profile_target.py --mode linear --clips 1 --reps 1 > gpurun_out/plain_linear_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:linear_blend -s 1 -c 1 -o gpurun_out/prof_linear_final -f \
    python tools/profile_target.py --mode linear --clips 1 --reps 1 > gpurun_out/ncu_linear_full.log 2>&1
python tools/profile_target.py --mode block --clips 1 --reps 1 > gpurun_out/plain_block_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:block_rows -s 1 -c 1 -o gpurun_out/prof_block_rows_final -f \
    python tools/profile_target.py --mode block --clips 1 --reps 1 > gpurun_out/ncu_block_full.log 2>&1
tail -n 1 gpurun_out/ncu_dense_full.log gpurun_out/ncu_smooth_full.log gpurun_out/ncu_linear_full.log gpurun_out/ncu_block_full.log
