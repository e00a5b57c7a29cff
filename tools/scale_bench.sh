#!/bin/bash
# Strong-scaling check on N GPUs of one box: tiger at 8192^2 and 16384^2, cost-balanced vs equal-height strips.
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
for size in 8192 16384; do
  for tag in ${MODES:-balanced equal}; do
    mode=""; [ "$tag" = "equal" ] && mode="--equal-strips"
    if [ "$N" = "1" ]; then
      python bench.py --gpus 1 --size $size --steps 100 --no-cpu-baseline --e2e-steps 2 $mode > $OUT/scale_n${N}_${size}_${tag}.json 2>$OUT/scale_err.log
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --size $size --steps 100 --no-cpu-baseline --e2e-steps 2 $mode > $OUT/scale_n${N}_${size}_${tag}.json 2>$OUT/scale_err.log
    fi
    python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/scale_n${N}_${size}_${tag}.json") if l.startswith("{")][-1])
    print("N=$N size=$size $tag: %.1f us/frame  %.0f Mpx/s  fine %.1f us bin %.1f us strips %s" % (d["ms_per_step"]*1e3, d["value"], d["roofline"]["kernel_ms"]*1e3, d["roofline"]["bin_kernel_ms"]*1e3, d["config"]["strip_tile_rows"]))
except Exception as e:
    print("N=$N size=$size $tag FAILED", e); print(open("$OUT/scale_err.log").read()[-1500:])
PY
    [ "$N" = "1" ] && break
  done
done
