#!/bin/bash
# usage: r2_final.sh <tag> -- launch lists at 56 M and 7 M, ncu --set full of the household kernel and the health update at 56 M, bench lines at 7 M and 2.7 M
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum"
for P in 7000000 2700000; do
  timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --people $P > $O/${T}_n1_$P.json 2> $O/${T}_n1_$P.err
  python tools/bench_brief.py $O/${T}_n1_$P.json
done
for P in 56000000 7000000; do
  timeout 900 ncu --metrics $M --clock-control none -c 140 --csv --log-file $O/${T}_launches_$P.csv python bench.py --steps 6 --warmup 1 --cpu-seconds 0 --people $P > $O/${T}_ncu_launch_$P.log 2>&1
  echo "ncu launch $P exit $?"; python tools/launch_summary.py $O/${T}_launches_$P.csv | cut -c1-130
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tiles|k_health" --launch-skip 4 --launch-count 6 \
  -o $O/${T}_full -f python bench.py --steps 4 --warmup 1 --cpu-seconds 0 > $O/${T}_ncu_full.log 2>&1
echo "ncu full exit $?"
