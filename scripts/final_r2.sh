#!/bin/bash
# Round-2 single-GPU measurement: bench, reference arm, ncu launch list, ncu full captures (each after a plain run of the same command)
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err || exit 1
tail -c 400 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-fit --no-secondary > gpurun_out/bench_short.json 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-fit --no-secondary > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-fit --no-secondary > gpurun_out/bench_short1.json 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:smo_persistent -s 1 -c 1 -f -o gpurun_out/prof_r2_smo \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-fit --no-secondary > gpurun_out/ncu_full.log 2>&1
python scripts/ncu_matvec_target.py || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:matvec_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_matvec \
    python scripts/ncu_matvec_target.py > gpurun_out/ncu_matvec.log 2>&1
ls -la gpurun_out | tail -8
