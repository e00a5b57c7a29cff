#!/bin/bash
# Fixed cost vs per-strip work on one GPU: render the strip rank g of N would own.
size=${1:-8192}
for spec in ${SPECS:-1:0 2:0 2:1 4:0 4:1 8:0 8:3 8:4}; do
  python bench.py --size $size --steps 100 --no-cpu-baseline --e2e-steps 1 --emulate-world $spec ${EXTRA:-} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; f=d['frame_stats']
print('size $size strip $spec rows %s: frame %.1f us | fine %.1f heavy %.1f bin %.1f | complex %d heavy %d tiles %d'%(d['config']['strip_tile_rows'], d['ms_per_step']*1e3, r['kernel_ms']*1e3, r['heavy_kernel_ms']*1e3, r['bin_kernel_ms']*1e3, f['complex_tiles'], f['heavy_tiles'], f['tiles']))"
done
