run() { echo "== $1"; env $1 timeout 200 python bench.py --steps 45 --warmup 9 --cpu-seconds 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), d['roofline']['us_per_launch'], d['kernel_ms_by_supergroup'], round(d['e2e']['ms_per_step'],4))"; }
for v in "$@"; do run "$v"; done
