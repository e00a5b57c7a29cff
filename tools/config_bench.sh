#!/bin/bash
# One bench line per BASELINE config that fits one GPU (besides the headline): tiger 1024^2 (config 2), tiger
# 16384^2 (strong-scaling base), rand_bezier 8192^2 (config 4), glyphs 4096^2 (config 5, on one GPU).
OUT=gpurun_out; mkdir -p $OUT
for spec in "tiger 1024" "tiger 16384" "rand_bezier 8192" "glyphs 4096"; do
  set -- $spec
  python bench.py --scene $1 --size $2 --steps 50 --no-cpu-baseline --e2e-steps 2 > $OUT/cfg_$1_$2.json 2> $OUT/cfg_err.log || tail -5 $OUT/cfg_err.log
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/cfg_$1_$2.json").read()); r=d["roofline"]; f=d["frame_stats"]
    print("$1 $2: %.1f us/frame %.0f Mpx/s | fine %.1f us (%.1f%% of HBM roofline) bin %.1f us | tiles %d with records %d overflow records %d | e2e %.0f Mpx/s" % (d["ms_per_step"]*1e3, d["value"], r["kernel_ms"]*1e3, 100*r["frac"], r["bin_kernel_ms"]*1e3, f["tiles"], f["complex_tiles"], f["overflow_records"], d["e2e"]["value"]))
except Exception as e: print("$1 $2 failed", e)
PY
done
