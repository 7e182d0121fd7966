#!/bin/bash
# One gpurun call: GPU parity tests, the bench at full size, an ncu launch list of one full-size step and full
# captures of the dominant kernels (one launch each, taken in the third window of the step).  Outputs under
# gpurun_out/ (scratch); summaries are copied to profiles/ by tools/collect_profiles.sh.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
python bench.py ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
ONE="--steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py $ONE > gpurun_out/${TAG}_launches_bench.log 2>&1
# count: per window k_enum_lin, 3 x onesweep (+ histogram), k_part_bounds, k_count_part, k_tab_apply_marked
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_enum_lin|DeviceRadixSortOnesweep|k_count_part|k_tab_apply_marked' -s 12 -c 6 -f \
    -o gpurun_out/${TAG}_count python bench.py $ONE > gpurun_out/${TAG}_ncu_count.log 2>&1
# correct: per window k_ec_lookup, k_ec_ext, k_ec_search
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_ec_lookup|k_ec_ext|k_ec_search' -s 6 -c 3 -f \
    -o gpurun_out/${TAG}_correct python bench.py $ONE > gpurun_out/${TAG}_ncu_correct.log 2>&1
ls -la gpurun_out
