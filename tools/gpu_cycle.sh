#!/bin/bash
# One GPU measurement cycle (run under gpurun): parity tests, bench line, ncu launch list + full capture of the exact kernel.
# usage: tools/gpu_cycle.sh [noprof] [quick]   quick = only the parity tests that exercise the exact kernels
mkdir -p gpurun_out
SEL=""
if [[ "$*" == *quick* ]]; then SEL="quick"; fi
timeout 1200 python -m pytest tests -m gpu -x -q ${SEL:+-k "fixtures or synthetic or ragged or dense or sweep"} > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err; cat gpurun_out/bench1.json; tail -3 gpurun_out/bench1.err
if [[ "$*" != *noprof* ]]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --reads 4000000 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dcb_exact -s 3 -c 1 -f -o gpurun_out/prof_exact python bench.py --steps 2 --warmup 3 --reads 4000000 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
fi
ls -la gpurun_out | head -20
