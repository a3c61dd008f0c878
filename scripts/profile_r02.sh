#!/bin/bash
# ncu captures behind profiles/r02_*: run on the GPU box from the repo root (gpurun -- bash scripts/profile_r02.sh); summaries: scripts/ncu_summary.py
set -x
mkdir -p gpurun_out
F="--set full --clock-control none"
# 1. launch list of the contract bench (cold-cache, serialised: compare SHARES)
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --extras 0 > gpurun_out/r02_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --extras 0 > gpurun_out/r02_ncu_bench.log 2>&1
# 2. the fused CG pair
python scripts/time_cg.py 256 > gpurun_out/r02_plain_cg.log 2>&1 &&
ncu $F -k regex:'k_star_cg_tma|k_phase' -s 40 -c 2 -f -o gpurun_out/r02_cg python scripts/time_cg.py 256 > gpurun_out/r02_ncu_cg.log 2>&1
# 3. stand-alone products: TMA stencil product and bulk-copy CSR product
python scripts/time_spmv.py 256 5 csr > gpurun_out/r02_plain_spmv.log 2>&1 &&
ncu $F -k regex:k_star_spmv_tma -s 6 -c 1 -f -o gpurun_out/r02_spmv python scripts/time_spmv.py 256 5 > gpurun_out/r02_ncu_spmv.log 2>&1
ncu $F -k regex:k_spmv_csr_bulk -s 3 -c 1 -f -o gpurun_out/r02_csr python scripts/time_spmv.py 256 5 csr > gpurun_out/r02_ncu_csr.log 2>&1
# 4. x-marching triangular solves 256^3
python scripts/time_trsv.py 256 > gpurun_out/r02_plain_trsv.log 2>&1 &&
ncu $F -k regex:k_star_trsv_pipe -s 2 -c 2 -f -o gpurun_out/r02_trsv python scripts/time_trsv.py 256 > gpurun_out/r02_ncu_trsv.log 2>&1
# 5. pattern-coded stencils (256^3 Neumann): product and the fused CG kernel in its per-row-diagonal form
python scripts/time_coded.py 256 neumann > gpurun_out/r02_plain_coded.log 2>&1 &&
ncu $F -k regex:'k_coded_spmv|k_star_cg_tma' -s 4 -c 4 -f -o gpurun_out/r02_coded python scripts/time_coded.py 256 neumann > gpurun_out/r02_ncu_coded.log 2>&1
ls -la gpurun_out/*.ncu-rep
# the reports stay on the box (gpurun brings back at most 64 MiB): text summaries only
for n in cg spmv csr trsv coded; do python scripts/ncu_summary.py gpurun_out/r02_$n.ncu-rep > gpurun_out/r02_ncu_full_$n.txt 2>/dev/null; done
rm -f gpurun_out/*.ncu-rep
