#!/bin/bash
# Copy one gpu_round.sh run (gpurun_out/<tag>_*) into profiles/ as the round's evidence:  bash tools/refresh_profiles.sh r01h
TAG=$1
set -e
cd "$(dirname "$0")/.."
rm -f profiles/r01[a-z]_bench*.json profiles/r01[a-z]_launches.* profiles/r01[a-z]_per_launch_cuda_events.txt profiles/r01[a-z]_ncu_full_conv_halo_fwd.txt
python tools/ncu_summarize.py gpurun_out/${TAG}_launches.csv ${TAG} > /dev/null
cp gpurun_out/${TAG}_launches.csv profiles/
cp gpurun_out/${TAG}_bench.json profiles/${TAG}_bench.json
[ -f gpurun_out/${TAG}_bench_ref.json ] && cp gpurun_out/${TAG}_bench_ref.json profiles/${TAG}_bench_ref.json
cp gpurun_out/${TAG}_profile_plan.txt profiles/${TAG}_per_launch_cuda_events.txt
[ -f gpurun_out/${TAG}_top.ncu-rep ] && python tools/ncu_src_top.py gpurun_out/${TAG}_top.ncu-rep 12 > profiles/${TAG}_ncu_full_conv_halo_fwd.txt
ls profiles
