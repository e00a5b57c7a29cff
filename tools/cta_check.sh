#!/bin/bash
# Round-2 starting point: validate and measure the CTA-cooperative path for costly tiles (PM_CTA_TILES, written at
# the end of round 1 without GPU time left).  Run under gpurun from the repo root after `make -C piet-metal_b200 cta`
# (CTA_FLAGS="-DPM_FINE_BULK=1" or "-DPM_CTA_TILES=1 -DPM_FINE_BULK=1" for the TMA bulk-copy prefetch).
set -u
V=$PWD/piet-metal_b200/variants/libpiet_metal_b200_cta.so
[ -f "$V" ] || { echo "build it first: make -C piet-metal_b200 cta"; exit 1; }
echo "== parity tests with the variant library"; PM_LIB=$V timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== default library"; tools/strip_study.sh 8192 2>&1 | sed -n '1p;7,8p'
echo "== variant"; PM_LIB=$V tools/strip_study.sh 8192 2>&1 | sed -n '1p;7,8p'
for spec in "tiger 1024" "glyphs 4096"; do set -- $spec
  for lib in "" "$V"; do PM_LIB=$lib python bench.py --scene $1 --size $2 --steps 50 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 $2 lib=${lib##*/}: %.1f us/frame fine %.1f us' % (d['ms_per_step']*1e3, d['roofline']['kernel_ms']*1e3))"; done; done
