#!/bin/bash
# launch lists (duration + DRAM bytes per kernel) of one clip per mode -> gpurun_out/r02_launches_<mode>.csv
mkdir -p gpurun_out
R='regex:dense_step|dense_strip|temporal_counts|linear_blend|linear_lowres|block_|argmax|upsample|feature|warp_step'
for m in dense dense_smooth dense_lowres block block_clip block_lowres linear linear_lowres; do
  python tools/profile_target.py --mode $m --clips 1 --reps 2 > gpurun_out/plain_$m.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$R" -s 0 -c 60 \
      --csv --log-file gpurun_out/r02_launches_$m.csv python tools/profile_target.py --mode $m --clips 1 --reps 2 > gpurun_out/ncu_$m.log 2>&1
  echo "$m rc=$?"
done
