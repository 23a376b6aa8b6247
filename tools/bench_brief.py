#!/usr/bin/env python
"""One compact line of a bench.py JSON line."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print(f"{d['config']['people_total'] / 1e6:.1f}M x{d['n_gpus']}: {d['ms_per_step']:.4f} ms/step ({d['value'] / 1e9:.1f} G/s), e2e {d['e2e']['ms_per_step']:.4f} ms, "
      f"launches {d['gpu_launches']}; " + ", ".join(f"{k} {v['us_per_launch']:.0f}us/{v['frac']:.3f}" for k, v in r["by_supergroup"].items()))
