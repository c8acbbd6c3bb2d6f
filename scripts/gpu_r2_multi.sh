#!/bin/bash
# multi-GPU evidence (gpurun --gpus N): NCCL gather test (N >= 2), bench at N GPUs: BASELINE config 2 (4096 envs/GPU) and config 5 (16384 envs/GPU)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== NCCL gather test"; timeout 900 python -m pytest tests/test_cuda_3_multigpu.py -m gpu -q --timeout=600 2>&1 | tail -4 | tee gpurun_out/r2_multigpu_test_n$N.txt
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $N "$@"; }
echo "== bench config 2, $N GPUs (driver flags)"; run --steps 20 --warmup 5 --no-scale > gpurun_out/r2_bench_cfg2_n$N.json 2> gpurun_out/bench_n$N.err || tail -5 gpurun_out/bench_n$N.err
echo "== bench config 5 (16384 envs per GPU), $N GPUs"; run --steps 20 --warmup 5 --no-scale --envs 16384 > gpurun_out/r2_bench_cfg5_n$N.json 2> gpurun_out/bench5_n$N.err || tail -5 gpurun_out/bench5_n$N.err
echo "== reference arm under torchrun"; run --impl reference --steps 20 --warmup 5 --no-python-loop > gpurun_out/r2_bench_ref_n$N.json 2>/dev/null
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for f in ('r2_bench_cfg2_n%s' % n, 'r2_bench_cfg5_n%s' % n, 'r2_bench_ref_n%s' % n):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, 'n_gpus', d['n_gpus'], 'value %.1fM' % (d['value'] / 1e6), 'e2e %.1fM' % (d['e2e']['value'] / 1e6), d.get('config', {}).get('workload', '')[:60],
              'single %.1fM' % (d['single_batch']['value'] / 1e6) if d.get('single_batch') else '', 'cores %s' % d['cpu_baseline']['cores'] if d.get('impl') else '')
    except Exception as e:
        print(f, 'unreadable', e)
PY
