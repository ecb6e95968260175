#!/bin/bash
# Round-end measurements on one B200 (run through gpurun; results land in gpurun_out/).
# No profiler here: ncu captures are separate calls (scripts/collect_ncu_*.sh).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r1_pytest_gpu.txt 2>&1; tail -2 gpurun_out/r1_pytest_gpu.txt
run() { out=$1; shift; python bench.py "$@" 2> gpurun_out/$out.err | tail -1 > gpurun_out/$out.json; python -c "
import json,sys
d=json.load(open('gpurun_out/$out.json'))
print('$out', d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('value'))"; }
run r1_bench_ba_100k_5k_1gpu
run r1_bench_ba_1k_200 --workload ba_1k_200
run r1_bench_ba_20k_2k_pcg --workload ba_20k_2k
run r1_bench_ba_20k_2k_dense --workload ba_20k_2k --linear-solver dense --steps 5 --warmup 3 --no-cpu-baseline
run r1_bench_ba_20k_2k_radial_dense --workload ba_20k_2k --num-intrinsics 3 --steps 5 --warmup 3 --no-cpu-baseline
run r1_bench_loc_1m_5k --workload loc_1m_5k --steps 5 --warmup 3
run r2_bench_detect_1020x768 --workload detect_1020x768 --steps 20 --warmup 3
python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/r1_bench_reference.err | tail -1 > gpurun_out/r1_bench_reference.json
