#!/bin/bash
# Round-2 ncu evidence for the batched prover (run under gpurun, one GPU; each ncu pass only after the plain run exited 0):
#   1. launch list of two bench steps (gpu__time_duration per launch; cold-cache and serialised: compare SHARES)
#   2. ncu --set full of one launch each of the shipped G1 and G2 table-gather kernels at the bench's shape (4096 proofs, c = 17)
set -e
python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r2_plain_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_batch4096.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r2_ncu_launches.log 2>&1
python tools/ab_stage.py 4096 17 > gpurun_out/r2_plain_ab.json
ncu --set full --clock-control none --import-source on -k regex:'k_msm_batch' -s 6 -c 2 -o gpurun_out/r2_msm_batch \
    python tools/ab_stage.py 4096 17 > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log
