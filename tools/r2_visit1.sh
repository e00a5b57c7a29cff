#!/bin/bash
# round 2, first GPU visit: parity (default + variants), baseline bench, A/B of the CTA and TMA-bulk variants
set -u
OUT=gpurun_out; mkdir -p $OUT
V=$PWD/piet-metal_b200/variants
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/v1_pytest.log 2>&1; tail -3 $OUT/v1_pytest.log
python bench.py --no-cpu-baseline > $OUT/v1_bench.json 2> $OUT/v1_bench.err; cat $OUT/v1_bench.json
for lib in libpm_cta.so libpm_bulk.so; do
  echo "== parity with $lib"; PM_LIB=$V/$lib timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
done
BENCH_ARGS="--steps 200" tools/ab_bench.sh 2>&1 | tee $OUT/v1_ab.txt
for lib in "" $V/libpm_cta.so; do
  echo "== strips lib=${lib##*/}"; PM_LIB=$lib tools/strip_study.sh 8192 2>&1 | tee -a $OUT/v1_strips.txt
done
for spec in "tiger 1024" "tiger 256" "glyphs 4096" "rand_bezier 8192"; do set -- $spec
  for lib in "" "$V/libpm_cta.so"; do PM_LIB=$lib python bench.py --scene $1 --size $2 --steps 50 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 $2 lib=${lib##*/}: %.1f us/frame fine %.1f us bin %.1f us' % (d['ms_per_step']*1e3, d['roofline']['kernel_ms']*1e3, d['roofline']['bin_kernel_ms']*1e3))" | tee -a $OUT/v1_cfg.txt; done; done
