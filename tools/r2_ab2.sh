#!/bin/bash
# usage: r2_ab2.sh <tag> -- per-warp vs per-block work lists in k_tiles (two builds), placement grid sizes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum"
timeout 600 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/${T}_pytest.log; tail -3 $O/${T}_pytest.log
cp june_b200/libjune_b200.so /tmp/main.so
for L in main blocklist; do
  [ $L = blocklist ] && cp june_b200/lib_blocklist.so june_b200/libjune_b200.so
  for P in 56000000 7000000 2700000; do
    timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --people $P > $O/${T}_${L}_$P.json 2> $O/${T}_${L}_$P.err
    echo -n "[$L] "; python tools/bench_brief.py $O/${T}_${L}_$P.json
  done
  for P in 56000000 7000000; do
    timeout 600 ncu --metrics $M --clock-control none -k regex:"k_tiles|k_place|k_newinf" -c 60 --csv --log-file $O/${T}_${L}_l$P.csv \
      python bench.py --steps 6 --warmup 1 --cpu-seconds 0 --people $P > /dev/null 2>&1
    echo "[$L] $P"; python tools/launch_summary.py $O/${T}_${L}_l$P.csv | cut -c1-130
  done
done
cp /tmp/main.so june_b200/libjune_b200.so
for C in 6 16 32; do
  JB_PLACE_CTAS=$C timeout 600 ncu --metrics $M --clock-control none -k regex:"k_place" -c 14 --csv --log-file $O/${T}_place$C.csv \
    python bench.py --steps 6 --warmup 1 --cpu-seconds 0 > /dev/null 2>&1
  echo "[place ctas $C]"; python tools/launch_summary.py $O/${T}_place$C.csv | cut -c1-130
done
