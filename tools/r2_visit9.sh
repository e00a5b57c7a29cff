#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-v9}
for lib in "" $PWD/piet-metal_b200/variants/libpm_w7.so; do
  echo "== strips 8192 lib=${lib##*/}"; PM_LIB=$lib tools/strip_study.sh 8192 2>&1 | tee -a $OUT/${TAG}_strips.txt
done
echo "== strips 16384"; tools/strip_study.sh 16384 2>&1 | tee -a $OUT/${TAG}_strips.txt
for spec in "tiger 1024" "tiger 256" "tiger 16384" "glyphs 4096" "rand_bezier 8192"; do set -- $spec
  python bench.py --scene $1 --size $2 --steps 50 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2: %.1f us/frame fine %.1f us heavy %.1f us bin %.1f us heavy_tiles %d' % (d['ms_per_step']*1e3, r['kernel_ms']*1e3, r['heavy_kernel_ms']*1e3, r['bin_kernel_ms']*1e3, d['frame_stats']['heavy_tiles']))" | tee -a $OUT/${TAG}_cfg.txt; done
