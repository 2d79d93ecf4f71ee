#!/bin/bash
# The evidence kept under profiles/ for a build, in one go on one GPU:  gpurun --timeout 1500 -- bash tools/final_artifacts.sh r02d
# (bench line + reference arm, DRAM traffic by range replay, ncu launch list of a short bench run, one ncu --set full capture of
# the MPPI call's kernel and of the noise kernel, compute-sanitizer over small invocations).  Outputs in gpurun_out/<prefix>_*.
P=${1:-r02d}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/${P}_bench_n1.json 2> gpurun_out/${P}_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/${P}_bench_reference_arm.json 2> gpurun_out/${P}_bench_reference_arm.err
timeout 300 bash tools/measure_traffic.sh > gpurun_out/${P}_traffic.log 2>&1
cp gpurun_out/r02_traffic_range.csv gpurun_out/${P}_mppi_traffic_range_ncu.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_bench_launches_ncu.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-ramp --no-c4 --no-c5 --rbpf-scans 2 > gpurun_out/${P}_launches_run.log 2>&1
PROBE_SHAPES="4,16,10" timeout 600 ncu --set full --import-source on --clock-control none -k regex:mppi_rollout -s 30 -c 1 -f \
    -o gpurun_out/${P}_mppi_call python tools/mppi_probe.py > gpurun_out/${P}_ncu_call.log 2>&1
PROBE_SHAPES="4,16,10" timeout 600 ncu --set full --import-source on --clock-control none -k regex:mppi_noise -s 30 -c 1 -f \
    -o gpurun_out/${P}_mppi_noise python tools/mppi_probe.py > gpurun_out/${P}_ncu_noise.log 2>&1
for tool in memcheck synccheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool python tools/sanitize_small.py mppi > gpurun_out/${P}_sanitizer_${tool}_mppi.log 2>&1
done
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_small.py rbpf > gpurun_out/${P}_sanitizer_memcheck_rbpf.log 2>&1
echo done
