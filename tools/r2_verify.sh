#!/bin/bash
# usage: r2_verify.sh <tag> -- what the driver runs at round end: gpu tests, smoke, both bench arms
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1
timeout 1800 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/${T}_pytest.log
tail -6 $O/${T}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke exit $?"; tail -3 $O/${T}_smoke.log
timeout 900 python bench.py --impl reference > $O/${T}_ref.json 2> $O/${T}_ref.err; echo "ref exit $?"; cut -c1-600 $O/${T}_ref.json
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench exit $?"; tail -3 $O/${T}_bench.err
python tools/bench_brief.py $O/${T}_bench.json
