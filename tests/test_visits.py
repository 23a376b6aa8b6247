"""Residence visits (groups/leisure/residence_visits.py:297-345, groups/household.py:110-149,202-205, policy/leisure_policies.py
:180-220): a household that receives a visitor keeps its residents home -- the ones already out doing leisure come back, the
ones whose turn has not come stay -- and the reference walks the people in world order.

* the sequential rules on a hand-made scenario (injected destinations), oracle and CUDA;
* CUDA (fixed-point resolution on the device) == oracle (the sequential walk) on a synthetic world with household and
  care-home visits, Philox mode, step by step;
* ``ChangeVisitsProbability`` moves the split between household and care-home visits.
"""
import datetime

import numpy as np
import pytest

from june_b200 import _lib
from june_b200.disease import DiseaseConfig, outcome_probabilities, synthetic_rates
from june_b200.interaction import Interaction, load_defaults
from june_b200.policy import ChangeVisitsProbability, Policies
from june_b200.simulator import ActivityManager, Epidemiology, InfectionSeed, Simulator
from june_b200.synth import generate_world
from june_b200.timer import Timer

from helpers import default_beta_table, have_gpu, make_engine, simulator_over

ACT_LEISURE, ACT_RESIDENCE = 3, 4
WD = [["medical_facility", "residence"], ["medical_facility", "primary_activity", "leisure", "residence"],
      ["medical_facility", "residence"], ["medical_facility", "leisure", "residence"], ["medical_facility", "residence"]]
WE = [["medical_facility", "leisure", "residence"]] * 3 + [["medical_facility", "residence"]]


def visits_world(n=20_000, seed=1):
    return generate_world(n, seed=seed, leisure=True, visits=True, sector_betas=load_defaults()["company_sector_betas"])


def make_sim(kind, world, policies=None, days=4, seed=3):
    inter, dc = Interaction.from_file(), DiseaseConfig("covid19")
    prob = outcome_probabilities(dc, *synthetic_rates())
    eng = make_engine(kind, world, prob, interaction=inter, disease=dc)
    config = {"time": {"initial_day": "2020-03-05 0:00", "total_days": days,
                       "step_duration": {"weekday": [1, 8, 1, 3, 11], "weekend": [4, 4, 4, 12]},
                       "step_activities": {"weekday": WD, "weekend": WE}},
              "activity_to_super_groups": {"medical_facility": ["hospitals"], "primary_activity": ["schools", "companies"],
                                           "leisure": ["pubs", "gyms", "cinemas", "groceries", "household_visits", "care_home_visits"],
                                           "residence": ["households", "care_homes"]}}
    timer = Timer.from_config(config)
    am = ActivityManager.from_config(config, world, policies or Policies.from_file(), timer)
    epi = Epidemiology(infection_seeds=[InfectionSeed(0.03, "2020-03-05", seed=seed)])
    return simulator_over(eng)(world, inter, timer, am, epi, disease_config=dc, outcome_prob=prob, seed=seed)


def household_of(world):
    """Residence group of every person, household ordinal ranges, the households' residents."""
    hh = next(sg for sg in world.supergroups if sg.name == "households")
    grp_of = np.repeat(np.arange(world.n_groups), np.diff(world.grp_offset.astype(np.int64)))
    act = (world.reg_info >> 8) & 7
    home = np.full(world.n_people, -1, np.int64)
    sel = (act == ACT_RESIDENCE) & (world.reg_member < world.n_people)
    home[world.reg_member[sel]] = grp_of[sel]
    return hh, home


def scenario(kind):
    """Five people of five households H2 < H3 < H4 < H5 < H6 (world order):
    V in H2 visits H5; X in H3 goes to a pub; A in H4 visits H3; B in H5 would go to a pub; D in H6 visits H4.
    X was out when A arrived at its household: fetched back.  B's household already had a visitor (V) at B's turn: B never
    draws.  A went visiting, then D came to A's household: A comes back home (and H3 stays "being visited" without a
    visitor).  V and D are out."""
    w = visits_world(6000, seed=5)
    hh, home = household_of(w)
    inter = Interaction.from_file()
    dc = DiseaseConfig("covid19")
    prob = outcome_probabilities(dc, *synthetic_rates())
    eng = make_engine(kind, w, prob, interaction=inter, disease=dc)
    adults = np.nonzero((w.age >= 20) & (w.age < 60) & (home >= hh.first_group) & (home < hh.first_group + hh.n_groups) &
                        (w.care_home_resident == 0))[0]
    first = {}
    for p in adults:
        first.setdefault(int(home[p]), int(p))
    order = sorted(first.values())       # one adult per household, households in world order
    v, x, a, b, d = (order[k] for k in (50, 300, 600, 900, 1200))
    pub = next(sg for sg in w.supergroups if sg.name == "pubs").first_group
    inj = np.full(w.n_people, -1, np.int32)
    inj[v], inj[x], inj[a], inj[b], inj[d] = home[b], pub, home[x], pub, home[a]
    eng.inject_leisure(inj)
    beta = default_beta_table(w, inter)
    mask = (1 << 0) | (1 << ACT_LEISURE) | (1 << ACT_RESIDENCE)
    st = eng.step(eng.make_params(now=0.0, delta_time=0.125, step=0, activity_mask=mask, seed=1,
                                  flags=_lib.STEP_INJECT_LEISURE, beta=beta))
    act = eng.fetch_person_state()["pstate"] & 7
    eng.close()
    return dict(v=int(act[v]), x=int(act[x]), a=int(act[a]), b=int(act[b]), d=int(act[d]), n_leisure=st["n_by_activity"][ACT_LEISURE])


def check_scenario(r):
    assert r["x"] == ACT_RESIDENCE, "X was at the pub when A arrived at its household: fetched back home"
    assert r["b"] == ACT_RESIDENCE, "B's household already had a visitor at B's turn: B stays home without drawing"
    assert r["a"] == ACT_RESIDENCE, "A went visiting, then D came to A's household: A comes back home"
    assert r["v"] == ACT_LEISURE and r["d"] == ACT_LEISURE and r["n_leisure"] == 2, "V and D are the ones who end up out"


def test_oracle_walks_the_people_in_world_order():
    check_scenario(scenario("oracle"))


@pytest.mark.gpu
def test_cuda_resolves_the_same_scenario():
    assert have_gpu()
    check_scenario(scenario("cuda"))


def run_steps(kind, world, n_steps, policies=None):
    sim = make_sim(kind, world, policies=policies)
    out = []
    while sim.timer.date < sim.timer.final_date and len(out) < n_steps:
        sim.epidemiology.infection_seeds_timestep(sim, sim.timer)
        sim.do_timestep()
        h = sim.history[-1]
        st = sim.device.fetch_person_state()
        out.append((h["n_new_infections"], h["n_infected"], h["n_draws"], tuple(h["n_by_activity"]), st["pstate"].copy(),
                    np.sort(sim.device.fetch_new_infections()[0])))
        next(sim.timer)
    return out


def test_visits_on_the_oracle_visitors_are_out_and_their_hosts_are_home():
    w = visits_world(12_000, seed=2)
    hh, home = household_of(w)
    base = run_steps("oracle", generate_world(12_000, seed=2, leisure=True, sector_betas=load_defaults()["company_sector_betas"]), 9)
    vis = run_steps("oracle", w, 9)
    # same venues, same tables otherwise: the visits activity sends more people out in the evening / weekend steps and
    # nobody during working hours (residence_visits.py:333-351)
    more = [vis[k][3][ACT_LEISURE] - base[k][3][ACT_LEISURE] for k in range(9)]
    assert all(m >= 0 for m in more) and sum(m > 50 for m in more) >= 2 and sum(m == 0 for m in more) >= 4, more
    assert sum(v[0] for v in vis) > 0


@pytest.mark.gpu
def test_cuda_matches_the_oracle_with_residence_visits():
    assert have_gpu()
    w = visits_world(40_000, seed=4)
    a, b = run_steps("cuda", w, 16), run_steps("oracle", w, 16)
    assert len(a) == len(b) == 16
    for k, (x, y) in enumerate(zip(a, b)):
        assert x[3] == y[3], f"step {k}: people per activity {x[3]} != {y[3]}"
        assert np.array_equal(x[4] & 7, y[4] & 7), f"step {k}: placement differs for {int(((x[4] ^ y[4]) & 7 != 0).sum())} people"
        assert x[:3] == y[:3] and np.array_equal(x[5], y[5]), f"step {k}: infections differ"
        assert np.array_equal(x[4], y[4]), f"step {k}: person state"
    assert sum(x[3][ACT_LEISURE] for x in a) > 10_000 and sum(x[0] for x in a) > 100


def test_change_visits_probability_moves_the_split():
    """(oracle) with care-home visits impossible, nobody is placed in a care home through leisure."""
    w = visits_world(15_000, seed=6)
    ch = next(sg for sg in w.supergroups if sg.name == "care_homes")

    def care_home_visitors(policies):
        sim = make_sim("oracle", w, policies=policies, days=3)
        n = 0
        while sim.timer.date < sim.timer.final_date:
            sim.do_timestep()
            n += sim.device.count_dynamic_members(ch.first_group, ch.n_groups) if hasattr(sim.device, "count_dynamic_members") else 0
            next(sim.timer)
        return sim
    pol = Policies([ChangeVisitsProbability("2020-01-01", "2021-01-01", new_residence_type_probabilities={"household": 1.0, "care_home": 0.0})])
    sim = care_home_visitors(pol)
    dist = sim.activity_manager.leisure.leisure_distributors["residence_visits"]
    assert dist.type_probabilities() == (1.0, 0.0)
    assert sim._visit_probabilities_on_device == (1.0, 0.0)
    sim2 = care_home_visitors(Policies([]))
    assert sim2.activity_manager.leisure.leisure_distributors["residence_visits"].type_probabilities() == (0.66, 0.34)
