#!/usr/bin/env python
"""Time the UNMODIFIED Python reference at configs[0] (a programmatic 10 k-person world, 10 simulated days, CPU) beside
the C port of the same algorithm (oracle/june_oracle.c) on the same world -- the anchor for "port vs real reference".
Build container only (needs /root/reference).  Writes profiles/r2_reference_timing.json."""
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


def main(n_people=10_000, days=10):
    random.seed(1)
    np.random.seed(1)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        world, dc = ref_world.build_world(n_people, seed=1)
        sim = ref_world.make_simulator(world, dc, total_days=days, cases_per_capita=0.02)
    steps = [0]
    orig = type(sim).do_timestep

    def step(self):
        orig(self)
        steps[0] += 1
    type(sim).do_timestep = step
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sink):
        sim.run()
    t_ref = time.perf_counter() - t0
    type(sim).do_timestep = orig
    n_steps = steps[0]
    # the C port on the compiled copy of the same world, through the drop-in class (oracle engine, one thread)
    from june_b200.june_plugin import make_b200_simulator_class
    from helpers import make_engine
    from june.simulator import Simulator
    random.seed(1)
    np.random.seed(1)
    with contextlib.redirect_stdout(sink):
        world2, dc2 = ref_world.build_world(n_people, seed=1)
        proto = ref_world.make_simulator(world2, dc2, total_days=days, cases_per_capita=0.02)
    os.environ["OMP_NUM_THREADS"] = "1"
    B = make_b200_simulator_class(Simulator)
    B.b200_engine = staticmethod(lambda soa, inter, dcfg, prob: make_engine("oracle", soa, prob, interaction=inter, disease=dcfg))
    sb = B(world=world2, interaction=proto.interaction, timer=proto.timer, activity_manager=proto.activity_manager,
           epidemiology=proto.epidemiology, tracker=None, record=None)
    with contextlib.redirect_stdout(sink):
        sb.do_timestep()  # compiles the world (one-off)
        next(sb.timer)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sink):
        sb.run()
    t_port = time.perf_counter() - t0
    out = dict(config="configs[0]: programmatic world, households / schools / companies / care homes / hospital",
               people=n_people, days=days, steps=n_steps,
               reference_python=dict(seconds=round(t_ref, 2), person_steps_per_s=round(n_people * n_steps / t_ref), cores=1),
               c_port_through_drop_in=dict(seconds=round(t_port, 3), person_steps_per_s=round(n_people * (n_steps - 1) / t_port), cores=1,
                                           note="oracle engine + the drop-in's Python host work per step"),
               host=dict(cpu=open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t") if os.path.exists("/proc/cpuinfo") else ""))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "r2_reference_timing.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
