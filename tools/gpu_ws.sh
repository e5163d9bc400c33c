#!/bin/bash
# Kernel iteration session: smoke first (a hang costs 60 s, not the whole call), parity tests on the warp-specialised
# kernel, kernel-only bench lines (SNK_KERNEL=v1 = filter_kernel), optional ncu capture ($2 = ncu). Outputs in gpurun_out/ (tag = $1).
TAG=${1:-ws}
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; rc=$?; tail -2 $OUT/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed rc=$rc: stopping"; exit 1; fi
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_text.py -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -8 $OUT/${TAG}_pytest.log
bench_line() {  # $1 = label, rest = env assignments
    local label=$1; shift
    env "$@" timeout 240 python bench.py --no-cpu-baseline --no-text --no-file --steps 10 > $OUT/${TAG}_bench_$label.json 2> $OUT/${TAG}_bench_$label.err
    python - <<PY
import json
try:
    l=json.loads(open("$OUT/${TAG}_bench_$label.json").read().strip().splitlines()[-1])
    print("$label value", round(l["value"],1), "frac", round(l["roofline"]["frac"],4), "ms", round(l["roofline"]["launch_ms"],3), "parity", l.get("stats_parity"))
except Exception as e: print("$label bench parse failed", e); print(open("$OUT/${TAG}_bench_$label.err").read()[-1500:])
PY
}
bench_line ws SNK_KERNEL=ws
if [ "$3" == "v1" ]; then bench_line v1 SNK_KERNEL=v1; fi
for w in $WPGS; do bench_line wpg$w SNK_WS_WPG=$w; done
if [ "$2" == "ncu" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:filter_ws_kernel -c 1 -f -o $OUT/prof_${TAG} \
  python bench.py --pairs 1048576 --steps 1 --warmup 1 --no-cpu-baseline --no-text --no-file > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/prof_${TAG}.ncu-rep
fi
