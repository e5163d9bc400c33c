#!/bin/bash
# Round-2 first GPU session (run with `gpurun --gpus 2`): all GPU tests incl. the N>1 parity tests, bench at N=1 and
# N=2 (NCCL), CLI timing breakdown. Outputs in gpurun_out/ (tag = $1).
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt; free -g >> $OUT/${TAG}_gpu.txt; df -h /dev/shm /tmp >> $OUT/${TAG}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q -x --durations=15 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -25 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
tail -c 3000 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
   > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err; echo "bench n2 exit $?"
tail -c 3000 $OUT/${TAG}_bench_n2.json; tail -5 $OUT/${TAG}_bench_n2.err
