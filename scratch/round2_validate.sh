#!/bin/bash
# Full validation pass on one B200: GPU tests, both bench arms, chain cycle profile, DFMA operand-pattern probe,
# ncu launch list of the bench command.
mkdir -p gpurun_out
T=$1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300 > gpurun_out/${T}_gpu_tests.log
cat gpurun_out/${T}_gpu_tests.log
timeout 600 python bench.py 2> gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference 2> gpurun_out/${T}_bench_ref.err | tail -1 > gpurun_out/${T}_bench_ref.json
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'],'roof',d['roofline'].get('frac'))
r=json.load(open('gpurun_out/${T}_bench_ref.json')); print('ref',r['value'])
"
MRGP_CHAIN_PROF=1 python scratch/chain_prof.py > gpurun_out/${T}_chain_prof_warm.log 2>&1
FLUSH=1 MRGP_CHAIN_PROF=1 python scratch/chain_prof.py > gpurun_out/${T}_chain_prof_cold.log 2>&1
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/dfma2 scratch/dfma_bench2.cu && /tmp/dfma2 > gpurun_out/${T}_dfma2.log 2>&1
cat gpurun_out/${T}_dfma2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${T}_ncu_bench.log 2>&1
tail -2 gpurun_out/${T}_ncu_bench.log | cut -c1-300
