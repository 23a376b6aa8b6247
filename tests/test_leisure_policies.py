"""Leisure policies: ``CloseLeisureVenue`` / ``ChangeLeisureProbability``
(policy/leisure_policies.py:69-180) acting on the leisure probability tables
(social_venue_distributor.py:176-205, leisure.py:205-228, activity_manager.py:200-209)."""
import datetime
import os
import sys

import numpy as np
import pytest

from june_b200 import _lib
from june_b200.interaction import load_defaults
from june_b200.leisure import generate_leisure_for_world
from june_b200.policy import ChangeLeisureProbability, CloseLeisureVenue, Policies
from june_b200.synth import generate_world

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def test_tables_match_the_reference_with_leisure_policies_active():
    """Fixture written by ``tests/golden/make_golden.py leisure_policies`` from the running reference."""
    from make_golden import LEISURE_POLICY_CASE as case
    z = np.load(os.path.join(HERE, "golden", "leisure_policy_tables.npz"))
    regions = [str(x) for x in z["regions"]]

    class W:  # the two things generate_leisure_for_world reads
        region_names = regions
        supergroups = [type("SG", (), {"name": n}) for n in ("pubs", "cinemas", "groceries")]

    leisure = generate_leisure_for_world(["pubs", "cinemas", "groceries"], W)
    assert leisure.activities == [str(x) for x in z["activities"]]
    pols = Policies([CloseLeisureVenue("2020-01-01", "2021-01-01", venues_to_close=case["venues_to_close"]),
                     ChangeLeisureProbability("2020-01-01", "2021-01-01", activity_reductions=case["activity_reductions"])])
    leisure.regional_compliance = {r: case["compliance"][k % len(case["compliance"])] for k, r in enumerate(regions)}
    for t, (when, dt, wh) in enumerate(case["slots"]):
        date = datetime.datetime.strptime(when, "%Y-%m-%d %H:%M")
        pols.apply_leisure_policies(date, leisure)
        does, cum = leisure.generate_leisure_probabilities_for_timestep(dt, wh, date)
        probs = np.diff(cum, axis=-1, prepend=0.0)
        assert np.allclose(does, z["does"][t], rtol=4e-16, atol=0), np.abs(does - z["does"][t]).max()
        assert np.allclose(probs, z["probs"][t], rtol=0, atol=3e-16)
        assert (probs[..., leisure.activities.index("cinema")] == 0).all(), "closed venues get no visitors"
    # the two regions differ only through their compliance
    assert not np.allclose(z["does"][0][0], z["does"][0][1])


def test_two_probability_changes_at_once_are_rejected_like_in_the_reference():
    class L:
        region_names = ["a"]
    pols = Policies([ChangeLeisureProbability("2020-01-01", "2021-01-01", {"pub": {"both_sexes": {"0-100": 0.5}}}),
                     ChangeLeisureProbability("2020-06-01", "2021-01-01", {"pub": {"both_sexes": {"0-100": 0.1}}})])
    pols.apply_leisure_policies(datetime.datetime(2020, 3, 1), L)
    with pytest.raises(ValueError):
        pols.apply_leisure_policies(datetime.datetime(2020, 7, 1), L)


def test_default_policy_file_carries_the_leisure_policies():
    pols = Policies.from_file()
    kinds = {type(p).__name__ for p in pols.leisure_policies}
    assert kinds == {"CloseLeisureVenue", "ChangeLeisureProbability", "ChangeVisitsProbability"}
    assert all(not p.is_active(datetime.datetime(2020, 6, 1)) for p in pols.leisure_policies)  # dated 3020 in the defaults


def make_sim(kind, world, policies, total_days, seed):
    from june_b200.disease import DiseaseConfig, outcome_probabilities, synthetic_rates
    from june_b200.interaction import Interaction
    from june_b200.simulator import ActivityManager, Epidemiology, InfectionSeed, Simulator
    from june_b200.timer import Timer
    from helpers import make_engine, simulator_over
    inter, dc = Interaction.from_file(), DiseaseConfig("covid19")
    prob = outcome_probabilities(dc, *synthetic_rates())
    eng = make_engine(kind, world, prob, interaction=inter, disease=dc)
    # config_simulation.yaml:15-41 without commute
    wd = [["medical_facility", "residence"], ["medical_facility", "primary_activity", "leisure", "residence"],
          ["medical_facility", "residence"], ["medical_facility", "leisure", "residence"], ["medical_facility", "residence"]]
    we = [["medical_facility", "leisure", "residence"]] * 3 + [["medical_facility", "residence"]]
    config = {"time": {"initial_day": "2020-03-02 0:00", "total_days": total_days,
                       "step_duration": {"weekday": [1, 8, 1, 3, 11], "weekend": [4, 4, 4, 12]},
                       "step_activities": {"weekday": wd, "weekend": we}},
              "activity_to_super_groups": {"medical_facility": ["hospitals"], "primary_activity": ["schools", "companies"],
                                           "leisure": ["pubs", "gyms", "cinemas", "groceries"],
                                           "residence": ["households", "care_homes"]}}
    timer = Timer.from_config(config)
    am = ActivityManager.from_config(config, world, policies, timer)
    epi = Epidemiology(infection_seeds=[InfectionSeed(0.01, "2020-03-02", seed=seed)])
    return simulator_over(eng)(world, inter, timer, am, epi, disease_config=dc, outcome_prob=prob, seed=seed)


def test_closed_venues_empty_out_and_reductions_reduce():
    """End to end through ``Simulator`` on the CPU oracle engine: with pubs/cinemas/gyms closed, leisure
    placement drops to the groceries' share; with every rate scaled by 0.05 it almost vanishes."""
    w = generate_world(30_000, seed=4, leisure=True, sector_betas=load_defaults()["company_sector_betas"])

    def leisure_count(policies):
        sim = make_sim("oracle", w, policies=policies, total_days=2, seed=9)
        n = 0
        while sim.timer.date < sim.timer.final_date:
            sim.do_timestep()
            n += sim.last_stats["n_by_activity"][3]
            next(sim.timer)
        return n

    base = leisure_count(Policies([]))
    closed = leisure_count(Policies([CloseLeisureVenue("2020-01-01", "2021-01-01", venues_to_close=["pub", "cinema", "gym"])]))
    reduced = leisure_count(Policies([ChangeLeisureProbability("2020-01-01", "2021-01-01", {
        a: {"both_sexes": {"0-100": 0.05}} for a in ("pub", "cinema", "grocery", "gym")})]))
    assert base > 5000
    assert 0 < closed < 0.8 * base
    assert reduced < 0.12 * base


@pytest.mark.gpu
def test_cuda_simulator_equals_oracle_simulator_under_leisure_policies():
    from helpers import have_gpu
    assert have_gpu()
    w = generate_world(30_000, seed=4, leisure=True, sector_betas=load_defaults()["company_sector_betas"])
    pol = lambda: Policies([CloseLeisureVenue("2020-03-03", "2021-01-01", venues_to_close=["pub", "gym"]),
                            ChangeLeisureProbability("2020-03-04", "2021-01-01", {"grocery": {"both_sexes": {"0-100": 0.3}}})])
    hist = []
    for kind in ("cuda", "oracle"):
        sim = make_sim(kind, w, pol(), total_days=4, seed=3)
        sim.run()
        hist.append([(h["n_new_infections"], h["n_infected"], tuple(h["n_by_activity"])) for h in sim.history])
    assert hist[0] == hist[1]
    assert sum(h[2][3] for h in hist[0]) > 1000
