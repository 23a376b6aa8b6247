#!/bin/bash
# usage: r2_full5.sh <tag> <people> <regex> <skip> <count> -- ncu --set full in a --config4 run
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1; P=$2; K=$3; S=$4; C=$5
timeout 1400 ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip $S --launch-count $C \
  -o $O/${T}_full -f python bench.py --steps 4 --warmup 1 --cpu-seconds 0 --people $P --config4 > $O/${T}_ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 $O/${T}_ncu_full.log
