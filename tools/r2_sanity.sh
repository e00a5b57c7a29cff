#!/bin/bash
# last visit of a round: GPU tests, smoke(), the default bench line
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/final_pytest.log 2>&1; grep -E "passed|failed" $OUT/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > $OUT/final_bench.json 2> $OUT/final_bench.err; cat $OUT/final_bench.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('value %.0f Mpx/s  %.1f us/frame  frac %.4f  fine %.1f us bin %.1f us  e2e %.0f  plan %.3f ms  cpu %.1f' % (d['value'], d['ms_per_step']*1e3, r['frac'], r['kernel_ms']*1e3, r['bin_kernel_ms']*1e3, d['e2e']['value'], d['e2e']['plan_ms'], d['cpu_baseline']['value']))"
