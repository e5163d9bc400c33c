#!/bin/bash
# Short GPU session while iterating on the kernel: kernel parity tests, a kernel-only bench line and
# (optionally) one full ncu capture. Outputs in gpurun_out/ (tag = $1; $2 = "ncu" to profile).
TAG=${1:-quick}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-text > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
tail -1 $OUT/${TAG}_bench.json
if [ "$2" == "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_kernel -c 1 -f -o $OUT/prof_${TAG} \
  python bench.py --pairs 1048576 --steps 1 --warmup 1 --no-cpu-baseline --no-text > $OUT/${TAG}_ncu_full.log 2>&1
fi
