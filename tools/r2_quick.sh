#!/bin/bash
# usage: r2_quick.sh <tag> "<pytest files>" "<people list>" [primary] -- selected gpu tests, then the step timeline at the given sizes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1
if [ -n "$2" ]; then
  timeout 1200 python -m pytest $2 -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/${T}_pytest.log; tail -4 $O/${T}_pytest.log
fi
for P in $3; do
  echo "== people $P $4"
  JB_STAMPS=1 timeout 600 python tools/stamps.py $P $4 2>&1 | tail -3
done
