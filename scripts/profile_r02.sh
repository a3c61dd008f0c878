#!/bin/bash
# ncu captures behind profiles/r02_*: run on the GPU box from the repo root (gpurun -- bash scripts/profile_r02.sh)
set -x
mkdir -p gpurun_out
F="--set full --clock-control none --import-source on"
# 1. launch list of the contract bench (cold-cache, serialised: compare SHARES)
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --extras 0 > gpurun_out/r02_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --extras 0 > gpurun_out/r02_ncu_bench.log 2>&1
# 2. the fused CG pair
python scripts/time_cg.py 256 > gpurun_out/r02_plain_cg.log 2>&1 &&
ncu $F -k regex:'k_star_cg_tma|k_phase' -s 40 -c 6 -f -o gpurun_out/r02_cg python scripts/time_cg.py 256 > gpurun_out/r02_ncu_cg.log 2>&1
# 3. stand-alone products: TMA stencil product and bulk-copy CSR product
python scripts/time_spmv.py 256 5 csr > gpurun_out/r02_plain_spmv.log 2>&1 &&
ncu $F -k regex:'k_star_spmv_tma|k_spmv_csr_bulk' -s 6 -c 6 -f -o gpurun_out/r02_spmv python scripts/time_spmv.py 256 5 csr > gpurun_out/r02_ncu_spmv.log 2>&1
# 4. x-marching triangular solves 256^3
python scripts/time_trsv.py 256 > gpurun_out/r02_plain_trsv.log 2>&1 &&
ncu $F -k regex:k_star_trsv_pipe -s 2 -c 2 -f -o gpurun_out/r02_trsv python scripts/time_trsv.py 256 > gpurun_out/r02_ncu_trsv.log 2>&1
ls -la gpurun_out/*.ncu-rep
