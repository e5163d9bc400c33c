#!/bin/bash
# Session 6 (round 2), two GPUs: N>1 tests after the parallel engine creation, SNK_GPUS=1 vs 2 on 16 M pairs (steady state).
OUT=gpurun_out; TAG=exp6; mkdir -p $OUT
timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; rc=$?; tail -1 $OUT/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed rc=$rc: stopping"; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > $OUT/${TAG}_pytest_multi.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_multi.log; tail -4 $OUT/${TAG}_pytest_multi.log
timeout 240 python tools/cli_compare.py --pairs 16000000 --skip-reference --env-sweep "SNK_GPUS=1;SNK_GPUS=2" > $OUT/${TAG}_cli_16m.json 2> $OUT/${TAG}_cli_16m.err; tail -c 2500 $OUT/${TAG}_cli_16m.json; tail -2 $OUT/${TAG}_cli_16m.err
