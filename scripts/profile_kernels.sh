#!/bin/bash
# ncu captures behind profiles/r01_*: run on the GPU box from the repo root (gpurun -- bash scripts/profile_kernels.sh)
set -x
mkdir -p gpurun_out
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,lts__t_sector_hit_rate.pct"
# 1. launch list of the contract bench (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
# 2. resident CG (config 3 at 592 systems = 4 per SM) and config 1
ncu $M --clock-control none -k regex:k_cg_resident -c 4 --csv --log-file gpurun_out/ncu_resident.csv python scripts/time_resident.py cfg1 cfg3 > gpurun_out/ncu_resident.log 2>&1
# 3. structured triangular solves 256^3
ncu $M --clock-control none -k regex:k_star_trsv -s 2 -c 2 --csv --log-file gpurun_out/ncu_trsv.csv python scripts/time_trsv.py 256 > gpurun_out/ncu_trsv.log 2>&1
# 4. BiCGstab(2) 256^3 advection-diffusion: launch list of one solve
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/ncu_bicg.csv python scripts/bench_configs.py cfg5 > gpurun_out/ncu_bicg.log 2>&1
