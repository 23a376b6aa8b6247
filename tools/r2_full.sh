#!/bin/bash
# usage: r2_full.sh <tag> <people> <kernel regex> [skip] [count] -- ncu --set full capture of the matching kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1; P=$2; K=$3; S=${4:-0}; C=${5:-4}
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip $S --launch-count $C \
  -o $O/${T}_full -f python bench.py --steps 2 --warmup 1 --cpu-seconds 0 --people $P > $O/${T}_ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 $O/${T}_ncu_full.log
