#!/bin/bash
# Session 5 (round 2), one GPU: BASELINE configs 2 and 4 at stated scale, config 5's shape (reference sample, parity, kernel; 10 % of its reads).
OUT=gpurun_out; TAG=exp5; mkdir -p $OUT
timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; rc=$?; tail -1 $OUT/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed rc=$rc: stopping"; exit 1; fi
timeout 420 python tools/run_configs.py --configs 2,4 --gpus 1 --timeout 200 > $OUT/${TAG}_configs.jsonl 2> $OUT/${TAG}_configs.err
timeout 200 python tools/run_configs.py --configs 5 --gpus 1 --scale 0.1 --timeout 120 >> $OUT/${TAG}_configs.jsonl 2>> $OUT/${TAG}_configs.err
cut -c1-2200 $OUT/${TAG}_configs.jsonl; tail -3 $OUT/${TAG}_configs.err
