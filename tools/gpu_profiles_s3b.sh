#!/bin/bash
# Second evidence pass of the session: full ncu captures of the kernels rewritten after the first pass.
mkdir -p gpurun_out
python tools/confusion_target.py > gpurun_out/plain_conf.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:confusion_v16 -s 1 -c 1 -o gpurun_out/prof_confusion_i64_s3 -f python tools/confusion_target.py > gpurun_out/ncu_conf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:confusion_v16 -s 3 -c 1 -o gpurun_out/prof_confusion_u8_s3 -f python tools/confusion_target.py >> gpurun_out/ncu_conf.log 2>&1
python tools/feature_target.py > gpurun_out/plain_feat.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:feature_rows -s 1 -c 1 -o gpurun_out/prof_feature_rows_s3 -f python tools/feature_target.py > gpurun_out/ncu_feat.log 2>&1
python tools/profile_target.py --mode block --clips 1 --reps 1 > gpurun_out/plain_block_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:block_rows -s 1 -c 1 -o gpurun_out/prof_block_rows_s3 -f python tools/profile_target.py --mode block --clips 1 --reps 1 > gpurun_out/ncu_block_full.log 2>&1
tail -n 1 gpurun_out/ncu_conf.log gpurun_out/ncu_feat.log gpurun_out/ncu_block_full.log
