#!/bin/bash
# One GPU-box session: parity tests, both bench arms, ncu launch list and one full capture of the
# dominant kernel. Outputs land in gpurun_out/ (tag = $1).
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
tail -1 $OUT/${TAG}_bench.json
if [ "$2" != "noref" ]; then
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref exit $?"
tail -1 $OUT/${TAG}_bench_ref.json
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-text > $OUT/${TAG}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_kernel -c 1 -f -o $OUT/prof_${TAG} \
  python bench.py --pairs 1048576 --steps 1 --warmup 1 --no-cpu-baseline --no-text > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -20
