#!/bin/bash
# Round-2 evidence in one GPU visit: bench lines of every BASELINE config that fits one GPU, the ncu launch list and one
# full ncu capture of the headline frame's kernels, and SM-issue evidence for config 4 (10 k Bezier paths).
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r02}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json
: > $OUT/${TAG}_configs.jsonl
for spec in "tiger 1024" "tiger 8192" "tiger 16384" "rand_bezier 8192" "glyphs 4096"; do set -- $spec
  python bench.py --scene $1 --size $2 --steps 100 --e2e-steps 3 --no-cpu-baseline 2>>$OUT/${TAG}_bench.err >> $OUT/${TAG}_configs.jsonl
done
python - <<PY
import json
for l in open("$OUT/${TAG}_configs.jsonl"):
    d=json.loads(l); r=d["roofline"]; f=d["frame_stats"]
    print("%-28s %8.1f us/frame %9.0f Mpx/s | fine %.1f us (%.3f of HBM) heavy %.1f bin %.1f plan %.2f ms | complex %d heavy %d | e2e %.0f Mpx/s" % (d["config"]["workload"], d["ms_per_step"]*1e3, d["value"], r["kernel_ms"]*1e3, r["frac"], r["heavy_kernel_ms"]*1e3, r["bin_kernel_ms"]*1e3, r["plan_ms"], f["complex_tiles"], f["heavy_tiles"], d["e2e"]["value"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 75 -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/${TAG}_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fine|k_heavy|k_seg|k_row|k_list" -s 50 -c 5 -f -o $OUT/${TAG}_frame \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 >> $OUT/${TAG}_ncu.log 2>&1
ncu -i $OUT/${TAG}_frame.ncu-rep --page raw --csv > $OUT/${TAG}_frame_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_frame.ncu-rep --page source --csv > $OUT/${TAG}_frame_source.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:"k_fine|k_heavy|k_seg|k_row|k_list" -s 50 -c 5 -f -o $OUT/${TAG}_cfg4 \
    python bench.py --scene rand_bezier --size 8192 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 >> $OUT/${TAG}_ncu.log 2>&1
ncu -i $OUT/${TAG}_cfg4.ncu-rep --page raw --csv > $OUT/${TAG}_cfg4_raw.csv 2>/dev/null
ls -la $OUT/${TAG}_*
