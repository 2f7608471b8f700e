python -m pytest tests -m gpu -x -q -k "confusion or temporal or metric or validation or test_step" 2>&1 | tail -3
python tools/metric_bench.py 2>&1 | grep -E "us_per_call|pred_|temporal|frac"
