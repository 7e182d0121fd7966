#!/bin/bash
# One gpurun call: GPU parity tests, the bench at full size, an ncu launch list and full captures
# of the dominant kernels.  Outputs under gpurun_out/ (scratch); summaries are copied to profiles/.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
python bench.py ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
SMALL="--reads 4000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py $SMALL > gpurun_out/${TAG}_launches_bench.log 2>&1
for kern in ${NCU_KERNELS:-k_count_probe k_ec_read k_ec_lookup k_enum k_count_resolve k_count_replay}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -s 1 -c 1 -f -o gpurun_out/${TAG}_$kern \
      python bench.py $SMALL > gpurun_out/${TAG}_ncu_$kern.log 2>&1
done
ls -la gpurun_out
