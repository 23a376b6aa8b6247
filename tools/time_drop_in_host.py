#!/usr/bin/env python
"""Host work of the drop-in per step: B200Simulator.do_timestep on real reference objects, time spent OUTSIDE the engine
(policies, beta table, parameter block, upload of reference-made infections, write-back of what changed).  Build container
only (needs /root/reference); the engine is the CPU checker here, its own time is subtracted."""
import contextlib
import io
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")]
import ref_harness  # noqa: E402
import ref_world  # noqa: E402


def measure(n_people, days=6):
    from june_b200.june_plugin import make_b200_simulator_class
    from helpers import make_engine
    random.seed(2)
    np.random.seed(2)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        world, dc = ref_world.build_world(n_people, seed=2)
        proto = ref_world.make_simulator(world, dc, total_days=days, cases_per_capita=0.02)
    from june.simulator import Simulator
    B = make_b200_simulator_class(Simulator)
    in_engine = [0.0]

    class Timed:
        def __init__(self, eng):
            self._e = eng

        def __getattr__(self, name):
            f = getattr(self._e, name)
            if not callable(f):
                return f

            def g(*a, **k):
                t0 = time.perf_counter()
                try:
                    return f(*a, **k)
                finally:
                    in_engine[0] += time.perf_counter() - t0
            return g
    B.b200_engine = staticmethod(lambda soa, inter, dcfg, prob: Timed(make_engine("oracle", soa, prob, interaction=inter, disease=dcfg)))
    sim = B(world=world, interaction=proto.interaction, timer=proto.timer, activity_manager=proto.activity_manager,
            epidemiology=proto.epidemiology, tracker=None, record=None)
    per_step, changed = [], []
    orig = B.do_timestep

    def step(self):
        in_engine[0] = 0.0
        t0 = time.perf_counter()
        orig(self)
        per_step.append((time.perf_counter() - t0 - in_engine[0]) * 1e3)
        changed.append(self.b200_last_stats["n_new_infections"] + self.b200_last_stats["n_recovered_step"] + self.b200_last_stats["n_deaths_step"])
    B.do_timestep = step
    with contextlib.redirect_stdout(sink):
        sim.run()
    per_step, changed = per_step[1:], changed[1:]  # the first step compiles the world (one-off)
    return dict(people=n_people, steps=len(per_step), host_ms_per_step_median=round(float(np.median(per_step)), 3),
                host_ms_per_step_max=round(float(np.max(per_step)), 3),
                changed_people_per_step=round(float(np.mean(changed)), 1), infected_at_end=sim.b200_last_stats["n_infected"])


if __name__ == "__main__":
    out = [measure(n) for n in (5_000, 20_000, 60_000)]
    path = os.path.join(ROOT, "profiles", "r2_drop_in_host_time.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
