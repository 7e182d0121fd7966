#!/bin/bash
# Final collection of a round, one gpurun call: GPU tests, one bench line per workload at full size, an ncu launch list
# of one full-size step, and per-kernel ncu metrics (one launch of every kernel, a window in the middle of the step).
# Outputs under gpurun_out/ (scratch); the summaries are copied to profiles/ by hand (profiles/README.md says which).
set -u
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench.err
for w in count trim k55; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
  echo "bench $w rc=$?"; tail -2 gpurun_out/${TAG}_bench_$w.err
done
ONE="--steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-cli"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py $ONE > gpurun_out/${TAG}_launches_bench.log 2>&1
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,\
lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,\
smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,\
sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --metrics $METRICS --clock-control none -k 'regex:k_enum_count|k_seg_scan|k_enum_lin|k_rp_hist|k_rp_scan|k_rp_pass|k_part_bounds|k_count_part|k_tab_apply_marked' -s 26 -c 16 -f \
    -o gpurun_out/${TAG}_count python bench.py $ONE > gpurun_out/${TAG}_ncu_count.log 2>&1
timeout 900 ncu --metrics $METRICS --clock-control none -k 'regex:k_ec_lookup|k_ec_cov|k_ec_setup|k_ec_rescue|k_ec_ext|k_ec_search|k_ec_merge' -s 14 -c 8 -f \
    -o gpurun_out/${TAG}_correct python bench.py $ONE > gpurun_out/${TAG}_ncu_correct.log 2>&1
timeout 900 ncu --metrics $METRICS --clock-control none -k 'regex:k_trim' -s 0 -c 1 -f \
    -o gpurun_out/${TAG}_trim python bench.py --workload trim $ONE > gpurun_out/${TAG}_ncu_trim.log 2>&1
ls -la gpurun_out/${TAG}_*
