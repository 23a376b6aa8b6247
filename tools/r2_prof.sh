#!/bin/bash
# usage: r2_prof.sh <tag> [people] [full]  -- ncu launch list of a short bench run (+ optional --set full capture of the interaction kernels)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1; P=${2:-56000000}
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none \
  -c 140 --csv --log-file $O/${T}_launches.csv python bench.py --steps 6 --warmup 1 --cpu-seconds 0 --people $P > $O/${T}_ncu_launch.log 2>&1
echo "ncu launch exit $?"
python tools/launch_summary.py $O/${T}_launches.csv
if [ "$3" = "full" ]; then
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_tiles|k_groups" --launch-skip 3 --launch-count 9 \
    -o $O/${T}_full -f python bench.py --steps 2 --warmup 1 --cpu-seconds 0 --people $P > $O/${T}_ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
