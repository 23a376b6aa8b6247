#!/bin/bash
# usage: r2_ab.sh <tag> <people> [pytest] "ENV=VAL ..." "ENV=VAL ..."   -- the driver's bench command under alternative environment settings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1; P=$2; shift; shift
if [ "$1" = "pytest" ]; then
  shift
  timeout 1800 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${T}_pytest.log
  tail -15 $O/${T}_pytest.log
fi
i=0
for e in "$@"; do
  i=$((i+1))
  env $e timeout 900 python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --people $P > $O/${T}_ab$i.json 2> $O/${T}_ab$i.err; rc=$?
  echo "== [$e] exit $rc"; [ $rc -ne 0 ] && tail -5 $O/${T}_ab$i.err
  python tools/bench_brief.py $O/${T}_ab$i.json
done
