#!/bin/bash
# ncu --set full captures of the top kernels of the large transforms (run under gpurun, one GPU).
python tools/profile_transforms.py all > gpurun_out/ptf_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_bucket_accum|k_ntt_pass' -s 8 -c 6 \
    -o gpurun_out/r1_transforms python tools/profile_transforms.py all > gpurun_out/ptf_ncu.log 2>&1
tail -3 gpurun_out/ptf_ncu.log
