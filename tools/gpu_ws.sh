#!/bin/bash
# Kernel iteration session: parity tests on the warp-specialised kernel, kernel-only bench lines for both kernels
# (SNK_KERNEL=v1 = filter_kernel), optional ncu capture ($2 = ncu). Outputs in gpurun_out/ (tag = $1).
TAG=${1:-ws}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_text.py -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -8 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-text --no-file --steps 10 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    l=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("ws value", l["value"], "frac", l["roofline"]["frac"], "ms", l["roofline"]["launch_ms"], "parity", l["stats_parity"])
except Exception as e: print("bench parse failed", e); print(open("$OUT/${TAG}_bench.err").read()[-2000:])
PY
if [ "$3" == "v1" ]; then
SNK_KERNEL=v1 timeout 600 python bench.py --no-cpu-baseline --no-text --no-file --steps 10 > $OUT/${TAG}_bench_v1.json 2> $OUT/${TAG}_bench_v1.err
python - <<PY
import json
l=json.loads(open("$OUT/${TAG}_bench_v1.json").read().strip().splitlines()[-1])
print("v1 value", l["value"], "frac", l["roofline"]["frac"], "ms", l["roofline"]["launch_ms"])
PY
fi
for w in $WPGS; do
SNK_WS_WPG=$w timeout 600 python bench.py --no-cpu-baseline --no-text --no-file --steps 10 > $OUT/${TAG}_bench_wpg$w.json 2> $OUT/${TAG}_bench_wpg$w.err
python - <<PY
import json
l=json.loads(open("$OUT/${TAG}_bench_wpg$w.json").read().strip().splitlines()[-1])
print("wpg $w value", l["value"], "frac", l["roofline"]["frac"], "ms", l["roofline"]["launch_ms"])
PY
done
if [ "$2" == "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_ws_kernel -c 1 -f -o $OUT/prof_${TAG} \
  python bench.py --pairs 1048576 --steps 1 --warmup 1 --no-cpu-baseline --no-text --no-file > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/prof_${TAG}.ncu-rep
fi
