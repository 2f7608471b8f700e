FUVS_BLOCK_XW=64 python -m pytest tests -m gpu -x -q -k "block or full_size or saturate" 2>&1 | tail -2
for xw in 0 64 128 32 0 64; do
  FUVS_BLOCK_XW=$xw python bench.py --mode block --no-e2e --no-cpu --no-modes --steps 60 > gpurun_out/s3_b.json 2> gpurun_out/s3_b.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/s3_b.json'))
print('block xw=$xw', d['ms_per_step']*1e3/12, 'us/interval', d['miou_counts_checksum'])
"; done
