#!/bin/bash
# usage: r2_iter.sh <tag> [pytest] -- one build->measure iteration: gpu tests, the bench at 56 / 7 / 2.7 M, launch lists at 56 M and 7 M
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1
if [ "$2" = "pytest" ]; then
  timeout 1800 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/${T}_pytest.log
  tail -4 $O/${T}_pytest.log
fi
for P in 56000000 7000000 2700000; do
  timeout 900 python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --people $P > $O/${T}_n1_$P.json 2> $O/${T}_n1_$P.err; rc=$?
  [ $rc -ne 0 ] && { echo "bench $P exit $rc"; tail -5 $O/${T}_n1_$P.err; }
  python tools/bench_brief.py $O/${T}_n1_$P.json
done
for P in 56000000 7000000; do
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none \
    -c 140 --csv --log-file $O/${T}_launches_$P.csv python bench.py --steps 6 --warmup 1 --cpu-seconds 0 --people $P > $O/${T}_ncu_launch_$P.log 2>&1
  echo "ncu launch $P exit $?"
  python tools/launch_summary.py $O/${T}_launches_$P.csv | cut -c1-130
done
