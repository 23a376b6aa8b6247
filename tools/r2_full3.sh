#!/bin/bash
# usage: r2_full3.sh <tag> <people> <regex> [count] -- ncu --set full of the kernels matching <regex> in a short bench run
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1; P=$2; K=$3; C=${4:-6}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip 4 --launch-count $C \
  -o $O/${T}_full -f python bench.py --steps 4 --warmup 1 --cpu-seconds 0 --people $P > $O/${T}_ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 $O/${T}_ncu_full.log
