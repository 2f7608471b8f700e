#!/bin/bash
# usage: tools/quick_bench.sh <mode> [ENV=VAL ...]  -> one line: mode, frames/s, us per interval, roofline frac
mode=$1; shift
env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --mode $mode --no-e2e --no-cpu --no-modes 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$mode $*', round(d['value'],1), 'fps', round(d['ms_per_step']/d['config']['intervals_per_step']*1000,2), 'us/interval', 'frac', round(d['roofline']['frac'],4))"
