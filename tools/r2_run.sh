#!/bin/bash
# usage: r2_run.sh <tag> [pytest|nopytest] [bench args...]  -- gpu tests + the driver's bench command at 56 M and 2.7 M
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1; shift
if [ "$1" = "pytest" ]; then
  timeout 1800 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${T}_pytest.log
  tail -15 $O/${T}_pytest.log
fi
shift
timeout 900 python bench.py --steps 20 --warmup 5 --cpu-seconds 0 "$@" > $O/${T}_bench56.json 2> $O/${T}_bench56.err; echo "bench56 exit $?"; tail -3 $O/${T}_bench56.err
python tools/bench_brief.py $O/${T}_bench56.json
timeout 600 python bench.py --steps 20 --warmup 5 --people 7000000 --cpu-seconds 0 "$@" > $O/${T}_bench7.json 2> $O/${T}_bench7.err; echo "bench7 exit $?"; tail -3 $O/${T}_bench7.err
python tools/bench_brief.py $O/${T}_bench7.json
timeout 600 python bench.py --steps 20 --warmup 5 --people 2700000 --cpu-seconds 0 "$@" > $O/${T}_bench27.json 2> $O/${T}_bench27.err; echo "bench27 exit $?"; tail -3 $O/${T}_bench27.err
python tools/bench_brief.py $O/${T}_bench27.json
