#!/bin/bash
# Session 2 (round 2): whole GPU test tier on one GPU, CLI host pipeline sweeps (tmpfs mmap writes, fast deflate), bench line.
OUT=gpurun_out; TAG=exp2; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_box.txt; nproc >> $OUT/${TAG}_box.txt; free -g >> $OUT/${TAG}_box.txt
timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; rc=$?; tail -1 $OUT/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed rc=$rc: stopping"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -6 $OUT/${TAG}_pytest.log
timeout 400 python tools/cli_compare.py --pairs 8000000 --skip-reference \
  --env-sweep "SNK_WRITE_MMAP=0;SNK_WRITE_MMAP=1;SNK_WRITE_MMAP=1,SNK_BATCH_READS=65536;SNK_WRITE_MMAP=1,SNK_BATCH_READS=131072;SNK_WRITE_MMAP=1,SNK_READ_THREADS=4,SNK_WRITE_THREADS=4" \
  > $OUT/${TAG}_cli_plain.json 2> $OUT/${TAG}_cli_plain.err; tail -c 2500 $OUT/${TAG}_cli_plain.json
timeout 400 python tools/cli_compare.py --pairs 4000000 --gz --gz-members 16 \
  --env-sweep "SNK_GZ_CODEC=zlib;SNK_GZ_SERIAL=1,SNK_GZ_CODEC=zlib;SNK_BATCH_READS=131072" > $OUT/${TAG}_cli_gz.json 2> $OUT/${TAG}_cli_gz.err; tail -c 2500 $OUT/${TAG}_cli_gz.json
timeout 500 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 1500 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
