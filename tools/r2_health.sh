#!/bin/bash
# usage: r2_health.sh <people> "E,B E,B ..." [primary] -- step timeline (tools/stamps.py) for JB_HEALTH_EARLY / JB_HEALTH_BLOCKS settings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
P=$1
for EB in $2; do
  E=${EB%,*}; B=${EB#*,}
  echo "== people $P health_early $E health_blocks $B $3"
  JB_HEALTH_EARLY=$E JB_HEALTH_BLOCKS=$B JB_STAMPS=1 timeout 600 python tools/stamps.py $P $3 2>&1 | tail -3
done
