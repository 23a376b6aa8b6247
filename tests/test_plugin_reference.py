"""The drop-in binding on the REAL reference objects (only where /root/reference exists, i.e. in
the build container; skipped on the GPU box).  The reference's own ``Simulator`` subclass runs its
``run()`` loop, seeding through the reference ``InfectionSeed``; ``do_timestep`` goes through the C
ABI (CPU oracle engine here, the CUDA library on a GPU)."""
import contextlib
import io
import os
import random
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not present")


def test_reference_simulator_subclass_runs_on_the_abi():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_world
    from june_b200.june_plugin import make_b200_simulator_class
    from helpers import make_engine

    random.seed(3)
    np.random.seed(3)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        world, dc = ref_world.build_world(2500, seed=3)
        ref_sim = ref_world.make_simulator(world, dc, total_days=12, cases_per_capita=0.03)
    from june.simulator import Simulator
    B200Simulator = make_b200_simulator_class(Simulator)
    B200Simulator.b200_engine = staticmethod(lambda soa, inter, dcfg, prob: make_engine("oracle", soa, prob, interaction=inter, disease=dcfg))
    sim = B200Simulator(world=world, interaction=ref_sim.interaction, timer=ref_sim.timer,
                        activity_manager=ref_sim.activity_manager, epidemiology=ref_sim.epidemiology,
                        tracker=None, record=None)
    with contextlib.redirect_stdout(sink):
        sim.run()
    people = list(world.people)
    n_inf = sum(p.infection is not None for p in people)
    n_imm = sum(p.immunity.get_susceptibility(sim._selector.infection_class.infection_id()) == 0.0 for p in people)
    assert sim._step == 54
    assert n_imm > 75 + 30, "the epidemic must have spread beyond the 3 % seed"
    assert n_inf > 0
    st = sim._eng.fetch_person_state()
    assert n_inf == int(((st["pstate"] & 0x08) != 0).sum())
    # subgroup lists of the reference objects stay untouched (placement lives on the device)
    assert all(len(sg.people) == 0 for g in world.companies.members for sg in g.subgroups)


def test_reference_simulator_subclass_with_leisure_commute_and_lockdown():
    """The same binding with the reference's own Leisure / Travel objects and lockdown policies: venues,
    transports and the policy records are compiled from the reference objects."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_world
    from june_b200.june_plugin import make_b200_simulator_class
    from helpers import make_engine

    random.seed(4)
    np.random.seed(4)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        world, dc = ref_world.build_world(2500, seed=4, with_leisure=True, with_visits=True)
        ref_sim = ref_world.make_simulator(world, dc, total_days=8, cases_per_capita=0.04, with_leisure=True, with_visits=True, lockdown=True)
    from june.simulator import Simulator
    B200Simulator = make_b200_simulator_class(Simulator)
    B200Simulator.b200_engine = staticmethod(lambda soa, inter, dcfg, prob: make_engine("oracle", soa, prob, interaction=inter, disease=dcfg))
    sim = B200Simulator(world=world, interaction=ref_sim.interaction, timer=ref_sim.timer,
                        activity_manager=ref_sim.activity_manager, epidemiology=ref_sim.epidemiology,
                        tracker=None, record=None)
    seen = {"leisure": 0, "skipped_work": 0}
    orig = B200Simulator.do_timestep

    def step(self):
        orig(self)
        st = self.b200_last_stats
        seen["leisure"] += st["n_by_activity"][3]
        if "primary_activity" in self.timer.activities and self.timer.date.day >= 6:
            seen["skipped_work"] += 1 if st["n_by_activity"][2] < 0.8 * first_work[0] else 0
        elif "primary_activity" in self.timer.activities and not first_work:
            first_work.append(st["n_by_activity"][2])

    first_work = []
    B200Simulator.do_timestep = step
    with contextlib.redirect_stdout(sink):
        sim.run()
    assert sim._soa.has_leisure and sim._soa.leisure_activities == ["pub", "cinema", "grocery", "residence_visits"]
    assert sim._soa.has_visits and int(sim._soa.visit_household_off[-1]) > 1000 and int(sim._soa.visit_care_home_off[-1]) > 10
    assert sim._b200_visit_split == (0.66, 0.34)
    assert seen["leisure"] > 1000, "people must have gone to pubs / cinemas / groceries"
    assert seen["skipped_work"] >= 1, "CloseCompanies / CloseSchools must keep people away from work"


def test_reference_simulator_subclass_feeds_the_record():
    """With a ``record``, the subclass reports the step's events through ``Record.accumulate`` with the
    reference's own table names and keyword arguments (a duck-typed recorder: PyTables is absent here)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_world
    from june_b200.june_plugin import make_b200_simulator_class
    from helpers import make_engine

    class Recorder:
        record_static_data = False

        def __init__(self):
            self.rows, self.steps, self.summaries = {}, 0, 0

        def accumulate(self, table_name, **kw):
            self.rows.setdefault(table_name, []).append(kw)

        def summarise_time_step(self, timestamp, world):
            self.summaries += 1

        def time_step(self, timestamp):
            self.steps += 1

        def parameters(self, **kw):
            pass

        def combine_outputs(self):
            pass

    random.seed(6)
    np.random.seed(6)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        world, dc = ref_world.build_world(2000, seed=6)
        ref_sim = ref_world.make_simulator(world, dc, total_days=25, cases_per_capita=0.15, schedule="simple")
    from june.simulator import Simulator
    B200Simulator = make_b200_simulator_class(Simulator)
    B200Simulator.b200_engine = staticmethod(lambda soa, inter, dcfg, prob: make_engine("oracle", soa, prob, interaction=inter, disease=dcfg))
    rec = Recorder()
    sim = B200Simulator(world=world, interaction=ref_sim.interaction, timer=ref_sim.timer,
                        activity_manager=ref_sim.activity_manager, epidemiology=ref_sim.epidemiology,
                        tracker=None, record=rec)
    with contextlib.redirect_stdout(sink):
        sim.run()
    people = {p.id: p for p in world.people}
    assert rec.steps == rec.summaries == sim._step
    inf = [r for r in rec.rows["infections"] if r["location_spec"] != "infection_seed"]
    assert sum(len(r["infected_ids"]) for r in inf) > 300
    specs = {g.spec for sg in ("households", "companies", "schools", "care_homes", "hospitals") for g in getattr(world, sg).members}
    for r in inf:
        assert r["location_spec"] in specs and len(r["infected_ids"]) == len(r["infector_ids"]) == len(r["infection_ids"])
        assert all(i in people for i in r["infected_ids"] + r["infector_ids"])
    dead = {p.id for p in world.people if p.dead}
    assert {r["dead_person_id"] for r in rec.rows.get("deaths", [])} == dead and len(dead) > 0
    hospital_ids = {h.id for h in world.hospitals.members}
    assert all(r["hospital_id"] in hospital_ids for t in ("hospital_admissions", "icu_admissions", "discharges") for r in rec.rows.get(t, []))
    assert len(rec.rows["recoveries"]) > 100 and len(rec.rows["symptoms"]) > 500 and len(rec.rows["hospital_admissions"]) > 0


def test_reference_simulator_subclass_vaccinates_from_the_reference_campaigns():
    """The reference's own ``VaccinationCampaigns`` object is compiled into the device tables; people of the
    target groups get their doses at the configured coverage and their immunity dicts follow."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_world
    from june_b200.june_plugin import make_b200_simulator_class
    from helpers import make_engine

    random.seed(5)
    np.random.seed(5)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        world, dc = ref_world.build_world(2500, seed=5)
        ref_sim = ref_world.make_simulator(world, dc, total_days=10, cases_per_capita=0.02, vaccination=True)
    from june.simulator import Simulator
    B200Simulator = make_b200_simulator_class(Simulator)
    B200Simulator.b200_engine = staticmethod(lambda soa, inter, dcfg, prob: make_engine("oracle", soa, prob, interaction=inter, disease=dcfg))
    sim = B200Simulator(world=world, interaction=ref_sim.interaction, timer=ref_sim.timer,
                        activity_manager=ref_sim.activity_manager, epidemiology=ref_sim.epidemiology,
                        tracker=None, record=None)
    with contextlib.redirect_stdout(sink):
        sim.run()
    iid = sim._selector.infection_class.infection_id()
    people = [p for p in world.people if not p.dead]
    old = [p for p in people if 60 <= p.age < 100]
    mid = [p for p in people if 30 <= p.age < 45]
    rest = [p for p in people if p.age < 30 or 45 <= p.age < 60]
    vacc = lambda ps: [p for p in ps if p.vaccinated is not None]
    # ref_world: Pfizer for 60-100 at coverage 0.6 (doses 0, 1), AstraZeneca for 30-45 at coverage 0.5 (dose 0)
    f_old, f_mid = len(vacc(old)) / len(old), len(vacc(mid)) / len(mid)
    assert abs(f_old - 0.6) < 4 * np.sqrt(0.24 / len(old)) + 0.02, f_old
    assert abs(f_mid - 0.5) < 4 * np.sqrt(0.25 / len(mid)) + 0.02, f_mid
    assert not vacc(rest)
    assert {p.vaccine_type for p in vacc(old)} == {"Pfizer"} and {p.vaccine_type for p in vacc(mid)} == {"AstraZeneca"}
    assert max(p.vaccinated for p in vacc(old)) == 1 and max(p.vaccinated for p in vacc(mid)) == 0
    never_infected = [p for p in vacc(old) if p.infection is None and p.immunity.susceptibility_dict.get(iid, 1.0) > 0]
    assert len(never_infected) > 20 and all(p.immunity.susceptibility_dict[iid] < 1.0 for p in never_infected)


def _side_by_side(engine_kind, n_people=10_000, days=10):
    """configs[0]: a programmatic world (households, schools, companies, care homes, hospital), the reference's own
    CPU ``Simulator`` and the drop-in run side by side from the same seeds."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_world
    from june_b200.june_plugin import make_b200_simulator_class
    from helpers import make_engine

    def build():
        random.seed(11)
        np.random.seed(11)
        with contextlib.redirect_stdout(io.StringIO()):
            world, dc = ref_world.build_world(n_people, seed=11)
            sim = ref_world.make_simulator(world, dc, total_days=days, cases_per_capita=0.02)
        return world, sim

    def run(sim, world):
        trace = []
        orig = type(sim).do_timestep

        def step(self):
            orig(self)
            trace.append(frozenset(i for i, p in enumerate(world.people) if p.infected))
        type(sim).do_timestep = step
        try:
            random.seed(5)
            np.random.seed(5)
            with contextlib.redirect_stdout(io.StringIO()):
                sim.run()
        finally:
            type(sim).do_timestep = orig
        # (person ids are a process-wide counter in the reference: people are compared by their position in the world)
        dead = frozenset(i for i, p in enumerate(world.people) if p.dead)
        immune = frozenset(i for i, p in enumerate(world.people) if p.immunity.get_susceptibility(170852960) == 0.0)
        return trace, dead, immune

    world_a, ref_sim = build()
    ages_a = [(p.age, p.sex) for p in world_a.people]
    # The one rule of the path that the device does not evaluate: "a child under 15 who has to stay home fetches a random
    # adult housemate" (individual_policies.py:57-77).  The reference itself switches it off when it runs domain
    # decomposed (`if mpi_size == 1`): the comparison is made against that mode of the reference.
    import june.policy.individual_policies as ref_ip
    saved_mpi_size = ref_ip.mpi_size
    ref_ip.mpi_size = 2
    try:
        ref_trace, ref_dead, ref_immune = run(ref_sim, world_a)
    finally:
        ref_ip.mpi_size = saved_mpi_size

    world_b, proto = build()
    assert [(p.age, p.sex) for p in world_b.people] == ages_a, "the two worlds must be built identically"
    from june.simulator import Simulator
    B200Simulator = make_b200_simulator_class(Simulator)
    B200Simulator.b200_engine = staticmethod(lambda soa, inter, dcfg, prob: make_engine(engine_kind, soa, prob, interaction=inter, disease=dcfg))
    B200Simulator.b200_reference_rng = True
    sim = B200Simulator(world=world_b, interaction=proto.interaction, timer=proto.timer, activity_manager=proto.activity_manager,
                        epidemiology=proto.epidemiology, tracker=None, record=None)
    trace, dead, immune = run(sim, world_b)
    assert len(trace) == len(ref_trace) >= 4 * days
    for k, (a, b) in enumerate(zip(ref_trace, trace)):
        assert a == b, f"step {k}: infected people differ ({len(a)} in the reference, {len(b)} here, {len(a ^ b)} apart)"
    assert dead == ref_dead and immune == ref_immune
    assert len(ref_trace[-1] | ref_immune) > 2 * int(0.02 * n_people), "the epidemic must have spread"
    return len(ref_trace), len(ref_immune)


def test_drop_in_equals_the_reference_simulator_step_for_step():
    """``B200Simulator`` with the reference's random streams reproduces the reference ``Simulator``: identical sets of
    infected people after every one of the run's steps, identical deaths and immunities at the end."""
    _side_by_side("oracle")


def test_active_events_are_applied_through_their_device_form_or_refused():
    """``do_timestep`` applies ``self.events`` like the reference (simulator.py:368-376): an active event without a
    device form is an error, not a silent no-op."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_world
    from june_b200.june_plugin import make_b200_simulator_class
    from helpers import make_engine
    random.seed(6)
    np.random.seed(6)
    with contextlib.redirect_stdout(io.StringIO()):
        world, dc = ref_world.build_world(1500, seed=6)
        ref_sim = ref_world.make_simulator(world, dc, total_days=2, cases_per_capita=0.03)
    from june.event import Event, Events
    from june.simulator import Simulator

    class Plain(Event):
        def initialise(self, world=None):
            pass

        def apply(self, world, simulator, activities, day_type):
            raise AssertionError("the reference form must not run on the device path")

    class WithDeviceForm(Plain):
        calls = 0

        def b200_apply(self, simulator):
            type(self).calls += 1
            assert simulator._eng is not None

    B200Simulator = make_b200_simulator_class(Simulator)
    B200Simulator.b200_engine = staticmethod(lambda soa, inter, dcfg, prob: make_engine("oracle", soa, prob, interaction=inter, disease=dcfg))

    def make(events):
        return B200Simulator(world=world, interaction=ref_sim.interaction, timer=ref_sim.timer, activity_manager=ref_sim.activity_manager,
                             epidemiology=ref_sim.epidemiology, tracker=None, record=None, events=Events(events))
    sim = make([WithDeviceForm(start_time="2020-03-02", end_time="2020-03-03"), Plain(start_time="2021-01-01", end_time="2021-02-01")])
    with contextlib.redirect_stdout(io.StringIO()):
        sim.do_timestep()
    assert WithDeviceForm.calls == 1
    sim.events = Events([Plain(start_time="2020-03-01", end_time="2020-04-01")])
    with pytest.raises(NotImplementedError, match="Plain"):
        with contextlib.redirect_stdout(io.StringIO()):
            sim.do_timestep()
