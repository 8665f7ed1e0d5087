#!/bin/bash
# A/B of the stream priorities (prefetch = geometry side stream, main = the captured step): ms per step, 1 GPU.
O=gpurun_out; mkdir -p $O
F="--steps 40 --warmup 8 --no-cpu-baseline --no-roofline --no-scaling-baseline"
run() { echo -n "prefetch=$1 main=$2: " >> $O/j_priority_ab.txt
  FT3D_PREFETCH_PRIORITY=$1 FT3D_MAIN_PRIORITY=$2 timeout 100 python bench.py $F 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%.3f ms/step  %.1f scans/s  e2e %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value']))
" >> $O/j_priority_ab.txt; }
: > $O/j_priority_ab.txt
run -2 -1
run 0 -1
run -5 -1
run -1 -2
run -2 -1
cat $O/j_priority_ab.txt
