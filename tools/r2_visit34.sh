#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-v34}
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1; grep -E "passed|failed" $OUT/${TAG}_pytest.log | head -5
BENCH_ARGS="--steps 200" tools/ab_bench.sh 2>&1 | tee $OUT/${TAG}_ab.txt
echo "== strips 8192"; SPECS="8:0 8:3" tools/strip_study.sh 8192 2>&1 | tee $OUT/${TAG}_strips.txt
for spec in "tiger 1024" "glyphs 4096" "rand_bezier 8192"; do set -- $spec
  python bench.py --scene $1 --size $2 --steps 50 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2: %.1f us/frame fine %.1f us heavy %.1f us bin %.1f us heavy_tiles %d' % (d['ms_per_step']*1e3, r['kernel_ms']*1e3, r['heavy_kernel_ms']*1e3, r['bin_kernel_ms']*1e3, d['frame_stats']['heavy_tiles']))" | tee -a $OUT/${TAG}_cfg.txt; done
