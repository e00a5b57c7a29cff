#!/bin/bash
# A/B: bench every library under piet-metal_b200/variants/ plus the default one (kernel times only).
shopt -s nullglob
for lib in piet-metal_b200/libpiet_metal_b200.so piet-metal_b200/variants/*.so; do
  for rep in 1 2; do
    PM_LIB=$PWD/$lib python bench.py --no-cpu-baseline --e2e-steps 1 ${BENCH_ARGS:-} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$lib rep$rep frame %.1f us fine %.1f us heavy %.1f us bin %.1f us'%(d['ms_per_step']*1e3, r['kernel_ms']*1e3, r.get('heavy_kernel_ms',0)*1e3, r['bin_kernel_ms']*1e3))"
  done
done
