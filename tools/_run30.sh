for i in 1 2; do
python bench.py --mode linear --no-e2e --no-cpu --no-modes --steps 100 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); print('linear', d['ms_per_step']*1e3/12, d['roofline']['frac'], d['clocks'])"
done
