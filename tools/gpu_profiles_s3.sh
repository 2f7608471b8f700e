#!/bin/bash
# Evidence pass of the third session (one gpurun call): plain runs first, then launch lists with DRAM bytes per mode,
# full ncu captures of the kernels that changed (linear bulk, temporal counts), then the bench lines (never under ncu).
mkdir -p gpurun_out
R='regex:dense_step|dense_strip|temporal_counts|linear_blend|linear_lowres|block_|argmax'
for m in dense linear block linear_lowres; do
  python tools/profile_target.py --mode $m --clips 1 --reps 2 > gpurun_out/plain_$m.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$R" -s 0 -c 40 \
      --csv --log-file gpurun_out/s3_launches_$m.csv python tools/profile_target.py --mode $m --clips 1 --reps 2 > gpurun_out/ncu_$m.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:linear_blend -s 1 -c 1 -o gpurun_out/prof_linear_s3_final -f \
    python tools/profile_target.py --mode linear --clips 1 --reps 1 > gpurun_out/ncu_linear_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:temporal_counts -s 1 -c 1 -o gpurun_out/prof_tcounts_s3 -f \
    python tools/profile_target.py --mode dense --clips 1 --reps 1 > gpurun_out/ncu_tc_full.log 2>&1
tail -n 1 gpurun_out/ncu_linear_full.log gpurun_out/ncu_tc_full.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_s3.json 2> gpurun_out/bench_ref_s3.err
python bench.py > gpurun_out/bench_s3.json 2> gpurun_out/bench_s3.err
python tools/metric_bench.py > gpurun_out/metric_s3.json 2> gpurun_out/metric_s3.err
tail -c 1500 gpurun_out/bench_s3.json
