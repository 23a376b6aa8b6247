#!/bin/bash
# Round-2 baseline on one B200: gpu tests, the driver's bench command at 56 M, launch list + full captures of the
# primary-activity step's interaction kernels.  Outputs under gpurun_out/.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2a_pytest.log
tail -3 $O/r2a_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2a_bench56.json 2> $O/r2a_bench56.err; echo "bench exit $?"
cat $O/r2a_bench56.json
timeout 600 python bench.py --steps 20 --warmup 5 --people 2700000 --cpu-seconds 0 > $O/r2a_bench27.json 2> $O/r2a_bench27.err; echo "bench27 exit $?"
cat $O/r2a_bench27.json
# launch list of the first timed (primary-activity) step at 56 M
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none \
  -c 120 --csv --log-file $O/r2a_launches56.csv python bench.py --steps 6 --warmup 1 --cpu-seconds 0 > $O/r2a_ncu_launch.log 2>&1
echo "ncu launch exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_tiles|k_groups" --launch-skip 3 --launch-count 9 \
  -o $O/r2a_full56 -f python bench.py --steps 2 --warmup 1 --cpu-seconds 0 > $O/r2a_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la $O | tail -20
