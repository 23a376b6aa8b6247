#!/bin/bash
# usage: r2_c4.sh <tag> -- placement-heavy parity tests, then the --config4 bench line at 56 M (and with the old one-block-per-256-people grid)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1
timeout 900 python -m pytest tests/test_leisure.py tests/test_leisure_policies.py tests/test_visits.py tests/test_commute.py tests/test_lockdown_tiers.py tests/test_golden_parity.py -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/${T}_pytest.log; tail -3 $O/${T}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --config4 > $O/${T}_c4.json 2> $O/${T}_c4.err; echo "bench exit $?"; python tools/bench_brief.py $O/${T}_c4.json
JB_PLACE_FULL_CTAS=100000 timeout 900 python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --config4 > $O/${T}_c4_old.json 2> $O/${T}_c4_old.err; echo "bench(old grid) exit $?"; python tools/bench_brief.py $O/${T}_c4_old.json
