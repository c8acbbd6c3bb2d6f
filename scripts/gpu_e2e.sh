#!/bin/bash
# tests touching HostStepper + e2e with different numbers of batches in flight
timeout 600 python -m pytest tests/test_cuda_1_rollout.py -x -q --timeout=180 2>&1 | tail -3
for eb in 8 4 16; do
  echo "== e2e batches $eb"
  timeout 600 python bench.py --no-cpu-baseline --no-scale --steps 6400 --warmup 640 --e2e-batches $eb 2>gpurun_out/e2e.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.1fM e2e %.1fM single %.1fM' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['single_batch_blocking']/1e6))" || tail -5 gpurun_out/e2e.err
done
