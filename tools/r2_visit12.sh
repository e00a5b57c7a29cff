#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-v12}
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1; tail -4 $OUT/${TAG}_pytest.log
python bench.py --e2e-steps 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
echo "== strips 8192"; tools/strip_study.sh 8192 2>&1 | tee $OUT/${TAG}_strips.txt
