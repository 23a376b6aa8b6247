"""Golden daily curves from the UNMODIFIED reference: 32 simulation seeds on the same programmatic
world (tests/golden/ref_world.py), 25 days, 2 % seeded on day 0.  Per day: new infections, currently
infected, currently in hospital (ward + ICU), cumulative deaths -- the quantities of the reference's
``summary.csv`` (``records_writer.py:151-164``) computed from the world state.

    python tests/golden/make_curves.py      # writes tests/golden/curves_ref.npz  (build container only)
"""
import contextlib
import io
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, HERE)
import ref_world  # noqa: E402

N_SEEDS, N_PEOPLE, DAYS, CASES = 32, 3000, 25, 0.02


def one_run(seed):
    random.seed(seed)
    np.random.seed(seed)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        world, dc = ref_world.build_world(N_PEOPLE, seed=0)
        sim = ref_world.make_simulator(world, dc, total_days=DAYS, cases_per_capita=CASES)
    import june.policy.individual_policies as indiv_mod
    from june.epidemiology.infection import InfectionSelector
    from june.simulator import Simulator
    indiv_mod.mpi_size = 2  # see make_golden.py: the guardian rule is off in domain-decomposed runs
    people = list(world.people)
    out = np.zeros((DAYS, 4), np.int64)
    state = {"new": 0}
    orig_infect = InfectionSelector.infect_person_at_time
    orig_step = Simulator.do_timestep

    def infect(self, person, time):
        orig_infect(self, person, time)
        state["new"] += 1

    def step(self):
        day = int(self.timer.now)
        orig_step(self)
        out[day, 0] += state["new"]
        state["new"] = 0
        out[day, 1] = sum(p.infection is not None for p in people)
        out[day, 2] = sum(p.infection is not None and p.subgroups is not None and p.subgroups.medical_facility is not None for p in people)
        out[day, 3] = sum(p.dead for p in people)

    InfectionSelector.infect_person_at_time = infect
    Simulator.do_timestep = step
    try:
        with contextlib.redirect_stdout(sink):
            sim.run()
    finally:
        InfectionSelector.infect_person_at_time = orig_infect
        Simulator.do_timestep = orig_step
    return out, len(people)


if __name__ == "__main__":
    curves = []
    for s in range(N_SEEDS):
        c, n = one_run(s)
        curves.append(c)
        print(s, "total infected", int(c[:, 0].sum()), "peak hospital", int(c[:, 2].max()), "deaths", int(c[-1, 3]), flush=True)
    np.savez_compressed(os.path.join(HERE, "curves_ref.npz"), curves=np.array(curves), n_people=n, days=DAYS, cases=CASES)
