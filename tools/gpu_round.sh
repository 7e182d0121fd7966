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
# Full captures (--set full) replay a kernel ~40 times and save / restore everything it writes (the 16 GiB filter, the
# table) between passes: ~6 minutes per kernel set at full size.  NCU_FULL=1 asks for them (profiles/r01g_* came from
# such a run); the default is the handful of metrics the roofline discussion uses (4 passes).
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,\
lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,\
smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,\
sm__throughput.avg.pct_of_peak_sustained_elapsed
if [ "${NCU_FULL:-0}" = 1 ]; then SET="--set full --import-source on"; else SET="--metrics $METRICS"; fi
# count: per window k_enum_count, k_enum_lin, onesweep launches, k_part_bounds, k_count_part, k_tab_apply_marked
timeout 900 ncu $SET --clock-control none -k 'regex:k_enum_count|k_enum_lin|DeviceRadixSortOnesweep|k_part_bounds|k_count_part|k_tab_apply_marked' -s 36 -c 18 -f \
    -o gpurun_out/${TAG}_count python bench.py $ONE > gpurun_out/${TAG}_ncu_count.log 2>&1
# correct: per window k_ec_lookup, k_ec_cov, k_ec_setup, k_ec_ext, k_ec_search, k_ec_merge
timeout 900 ncu $SET --clock-control none -k 'regex:k_ec_lookup|k_ec_cov|k_ec_setup|k_ec_ext|k_ec_search|k_ec_merge' -s 12 -c 6 -f \
    -o gpurun_out/${TAG}_correct python bench.py $ONE > gpurun_out/${TAG}_ncu_correct.log 2>&1
ls -la gpurun_out
