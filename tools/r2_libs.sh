#!/bin/bash
# usage: r2_libs.sh <tag> <people> lib1.so lib2.so ...  -- the driver's bench command with alternative builds of the library
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1; P=$2; shift; shift
cp june_b200/libjune_b200.so /tmp/orig.so
for l in "$@"; do
  cp june_b200/$l june_b200/libjune_b200.so
  timeout 900 python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --people $P > $O/${T}_$l.json 2> $O/${T}_$l.err; echo "== [$l] exit $?"
  python tools/bench_brief.py $O/${T}_$l.json
done
cp /tmp/orig.so june_b200/libjune_b200.so
