#!/bin/bash
# A/B of the small-crowd kernel's LP3 pass: block-compacted queue (default) vs warp-level queue (-DCS_FLAT_WARP_LP3=1).
# build_probe/probe and build_probe/probe_w are built here (no GPU needed) and travel with the snapshot.
mkdir -p gpurun_out
{
echo "== probe (block queue), 8 presteps";  timeout 300 build_probe/probe 8
echo "== probe (warp queue), 8 presteps";   timeout 300 build_probe/probe_w 8
echo "== probe (block queue), 18 presteps"; timeout 300 build_probe/probe 18 | grep -A2 "^B="
echo "== probe (warp queue), 18 presteps";  timeout 300 build_probe/probe_w 18 | grep -A2 "^B="
} > gpurun_out/ab_lp3.txt 2>&1
echo "== pytest gpu (default build)"; timeout 900 python -m pytest tests -m gpu -x -q --timeout=180 2>&1 | tail -4
echo "== rebuild with warp queue"; python - <<'PY'
from crowdnav_b200 import build
build.build(force=True, extra=['-DCS_FLAT_WARP_LP3=1'])
PY
echo "== pytest gpu (warp queue build)"; timeout 900 python -m pytest tests -m gpu -x -q --timeout=180 2>&1 | tail -4
cat gpurun_out/ab_lp3.txt
