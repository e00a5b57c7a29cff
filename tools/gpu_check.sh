#!/bin/bash
# One GPU-box visit: parity tests, bench, launch list, one full ncu capture of the fill kernel
# (exported as CSV so that it can be read without the .ncu-rep).  Run under gpurun from the repo root.
set -u
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-run}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
fi
timeout 600 python bench.py ${BENCH_ARGS:-} > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench.json
if [ "${BENCH16:-0}" = "1" ]; then
  timeout 600 python bench.py --size 16384 --no-cpu-baseline --steps 100 > $OUT/${TAG}_bench16k.json 2>> $OUT/${TAG}_bench.err
  cat $OUT/${TAG}_bench16k.json
fi
if [ "${PROFILE:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 --frame-events > $OUT/${TAG}_ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fine -s 5 -c 1 -f -o $OUT/${TAG}_fine \
      python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 --frame-events >> $OUT/${TAG}_ncu_bench.log 2>&1
  ncu -i $OUT/${TAG}_fine.ncu-rep --page raw --csv > $OUT/${TAG}_fine_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_fine.ncu-rep --page source --csv > $OUT/${TAG}_fine_source.csv 2>/dev/null
  if [ "${PROFILE_BIN:-0}" = "1" ]; then
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_seg|k_row" -s 10 -c 2 -f -o $OUT/${TAG}_bin \
        python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 --frame-events >> $OUT/${TAG}_ncu_bench.log 2>&1
    ncu -i $OUT/${TAG}_bin.ncu-rep --page raw --csv > $OUT/${TAG}_bin_raw.csv 2>/dev/null
    ncu -i $OUT/${TAG}_bin.ncu-rep --page source --csv > $OUT/${TAG}_bin_source.csv 2>/dev/null
  fi
  ls -la $OUT
fi
