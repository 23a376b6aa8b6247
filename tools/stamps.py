"""Debug timeline of the step (JB_STAMPS=1): runs the bench world for a number of queued, L2-flushed steps of
ONE shape (weekday residence-only steps by default) and lets the library print the mean start offset of every
kernel at destroy time.  usage: JB_STAMPS=1 python tools/stamps.py [people] [primary]"""
import os
import sys

os.environ.setdefault("JB_STAMPS", "1")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

from june_b200.device import DeviceWorld  # noqa: E402
from june_b200.disease import DiseaseConfig, outcome_probabilities, synthetic_rates  # noqa: E402
from june_b200.interaction import Interaction, load_defaults  # noqa: E402
from june_b200.synth import generate_world  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_700_000
primary = len(sys.argv) > 2
inter, dc = Interaction.from_file(), DiseaseConfig("covid19")
world = generate_world(n, seed=0, sector_betas=load_defaults()["company_sector_betas"])
dev = DeviceWorld(world, inter, dc, outcome_probabilities(dc, *synthetic_rates()), device=0)
rng = np.random.default_rng(0)
dev.infect(np.sort(rng.choice(n, int(0.03 * n), replace=False)).astype(np.uint32), time=-4.0, seed=1, step=10_000)
beta = inter.beta_table(world.beta_rules, [1.0] * world.n_regions, [None] * world.n_regions)
mask = 0b10101 if primary else 0b10001
now = 0.0
for i in range(60):
    if i == 20:  # CUDA-event brackets (what bench.py reports) over the last 40 steps, for comparison with the stamps
        dev.sync()
        dev.timing_enable(1)
        dev.timing_read(reset=True)
    dev.flush_l2()
    dev.step(dev.make_params(now, 1 / 24, i, mask, seed=3, flags=3, beta=beta), want_stats=False)
    now += 1 / 24
dev.sync()
tim = dev.timing_read()
print("event brackets: %.1f us per step over %d steps" % (1e3 * tim["step_ms"] / max(1, tim["n_steps"]), tim["n_steps"]), file=sys.stderr)
dev.close()
