#!/bin/bash
for q in auto block warp; do
  timeout 600 python bench.py --no-cpu-baseline --no-scale --e2e-batches 4 --lp3-queue $q 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$q value %.1fM single %.1fM kernel %.2f us' % (d['value']/1e6, d['single_stream']['value']/1e6, d['roofline']['avg_launch_us']))"
done
for st in 12 16; do
  timeout 600 python bench.py --no-cpu-baseline --no-scale --e2e-batches 4 --streams $st 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams $st value %.1fM' % (d['value']/1e6))"
done
