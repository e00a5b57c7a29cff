#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-v28}
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1; grep -E "passed|failed" $OUT/${TAG}_pytest.log
for h in 0 1 2 3; do
for spec in "tiger 1024" "tiger 256" "glyphs 4096" "rand_bezier 8192" "tiger 8192"; do set -- $spec
  PM_DEBUG_HEAVY_CTAS=$h python bench.py --scene $1 --size $2 --steps 50 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('heavy_ctas=$h $1 $2: %.1f us/frame fine %.1f us heavy %.1f us bin %.1f us heavy_tiles %d' % (d['ms_per_step']*1e3, r['kernel_ms']*1e3, r['heavy_kernel_ms']*1e3, r['bin_kernel_ms']*1e3, d['frame_stats']['heavy_tiles']))" | tee -a $OUT/${TAG}_cfg.txt; done; done
