#!/bin/bash
# Final session of round 2 (one GPU): whole GPU test tier, both bench arms, ncu launch list + one full capture of the production
# kernel, compute-sanitizer memcheck over the checked-tile / contaminant / text-path tests. Outputs in gpurun_out/ (tag r2_final).
OUT=gpurun_out; TAG=r2_final; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_box.txt 2>&1; nproc >> $OUT/${TAG}_box.txt
timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; rc=$?; tail -1 $OUT/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed rc=$rc: stopping"; exit 1; fi
timeout 800 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
timeout 400 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; tail -c 600 $OUT/${TAG}_bench.json; tail -2 $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref exit $?"; tail -c 400 $OUT/${TAG}_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-text --no-file > $OUT/${TAG}_launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:filter_kernel -c 1 -f -o $OUT/prof_${TAG} \
  python bench.py --pairs 1048576 --steps 1 --warmup 1 --no-cpu-baseline --no-text --no-file > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/prof_${TAG}.ncu-rep
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -m gpu tests/test_gpu_text.py \
  "tests/test_gpu_parity.py::test_mixed_checked_and_unchecked_tiles" "tests/test_gpu_parity.py::test_engine_matches_oracle_contam" \
  "tests/test_gpu_parity.py::test_tile_fov_flags_in_len" "tests/test_gpu_parity.py::test_len_beyond_the_row_raises_the_length_flag" \
  > $OUT/${TAG}_sanitizer.txt 2>&1; echo "sanitizer exit $?" >> $OUT/${TAG}_sanitizer.txt; tail -5 $OUT/${TAG}_sanitizer.txt
