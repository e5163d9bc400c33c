#!/bin/bash
# Session 4 (round 2), 8 GPUs of one box: N>1 parity tests, then BASELINE configs 3, 5 and 4 at their stated scale through
# the CLI with SNK_GPUS=8 (streamed input, see tools/run_configs.py). Reference / kernel legs of these shapes run on 1-GPU boxes.
OUT=gpurun_out; TAG=exp4; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_box.txt; nproc >> $OUT/${TAG}_box.txt; free -g >> $OUT/${TAG}_box.txt; df -h /dev/shm >> $OUT/${TAG}_box.txt
timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; rc=$?; tail -1 $OUT/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed rc=$rc: stopping"; exit 1; fi
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q > $OUT/${TAG}_pytest_multi.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_multi.log; tail -4 $OUT/${TAG}_pytest_multi.log
timeout 900 python tools/run_configs.py --configs 3,5,4 --no-reference --no-kernel --timeout 600 > $OUT/${TAG}_configs.jsonl 2> $OUT/${TAG}_configs.err
cut -c1-1500 $OUT/${TAG}_configs.jsonl; tail -3 $OUT/${TAG}_configs.err
