# usage: bench_line.sh <people> [ENV=VAL ...]   -- one compact line of a bench.py run
n=$1; shift
env "$@" timeout 900 python bench.py $BENCH_ARGS --people $n --steps 45 --warmup 9 --cpu-seconds 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$n $*', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'roof', r['us_per_launch'], r['frac'], d['kernel_ms_by_supergroup'])"
