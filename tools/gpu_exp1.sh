#!/bin/bash
# Experiment session 1 (round 2): ws kernel variants (group mapping, group size, J) + CLI host pipeline sweeps.
# (record of the experiment: the J=2 tuning build libsnk_engine_j2.so - SNK_CXXFLAGS=-DSNK_WS_J=2 SNK_ENGINE_LIB_NAME=... - aborted on the GPU and was dropped)
OUT=gpurun_out; TAG=exp1; mkdir -p $OUT
timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; rc=$?; tail -1 $OUT/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed rc=$rc: stopping"; exit 1; fi
bench_line() {
    local label=$1; shift
    env "$@" timeout 200 python bench.py --no-cpu-baseline --no-text --no-file --steps 10 > $OUT/${TAG}_bench_$label.json 2> $OUT/${TAG}_bench_$label.err
    python - <<PY
import json
try:
    l=json.loads(open("$OUT/${TAG}_bench_$label.json").read().strip().splitlines()[-1])
    print("$label value", round(l["value"],1), "frac", round(l["roofline"]["frac"],4), "ms", round(l["roofline"]["launch_ms"],3), "parity", l.get("stats_parity"))
except Exception as e: print("$label bench parse failed", e); print(open("$OUT/${TAG}_bench_$label.err").read()[-800:])
PY
}
bench_line ws_inter X=1
bench_line ws_block SNK_WS_MAP=block
bench_line ws_inter_wpg2 SNK_WS_WPG=2
bench_line ws_inter_wpg8 SNK_WS_WPG=8
J2=$PWD/soapnuke_b200/lib/libsnk_engine_j2.so
bench_line j2_inter SNK_ENGINE_LIB=$J2
bench_line j2_inter_wpg2 SNK_ENGINE_LIB=$J2 SNK_WS_WPG=2
bench_line j2_block SNK_ENGINE_LIB=$J2 SNK_WS_MAP=block
# CLI host pipeline: 8 M pairs plain, batch / thread sweeps; then multi-member .gz input + .gz output
timeout 600 python tools/cli_compare.py --pairs 8000000 --skip-reference \
  --env-sweep "SNK_BATCH_READS=32768;SNK_BATCH_READS=131072;SNK_BATCH_READS=262144;SNK_BATCH_READS=524288;SNK_BATCH_READS=262144,SNK_READ_THREADS=4,SNK_WRITE_THREADS=4;SNK_BATCH_READS=262144,SNK_READ_THREADS=12,SNK_WRITE_THREADS=12" \
  > $OUT/${TAG}_cli_plain.json 2> $OUT/${TAG}_cli_plain.err; tail -c 3000 $OUT/${TAG}_cli_plain.json
timeout 600 python tools/cli_compare.py --pairs 4000000 --gz --gz-members 16 --skip-reference \
  --env-sweep "SNK_GZ_SERIAL=1;SNK_BATCH_READS=131072" > $OUT/${TAG}_cli_gz.json 2> $OUT/${TAG}_cli_gz.err; tail -c 2000 $OUT/${TAG}_cli_gz.json
nproc
