#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"decode|permute_rows" -f -o gpurun_out/r02_final python tools/ncu_final.py > gpurun_out/ncu_final.log 2>&1
tail -5 gpurun_out/ncu_final.log
ls -la gpurun_out/r02_final.ncu-rep
