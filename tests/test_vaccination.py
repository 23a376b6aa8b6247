"""Vaccination: VaccinationCampaigns.apply + update_vaccine_effect once per
simulated day (vaccination_campaign.py:340-385, vaccines.py:109-150,319-376) and the symptomatic
efficacy in the outcome table of new infections (health_index.py:183-204)."""
import datetime

import numpy as np
import pytest

from june_b200 import _lib
from june_b200.interaction import load_defaults
from june_b200.synth import generate_world
from june_b200.vaccination import Vaccine, VaccinationCampaign, VaccinationCampaigns

from helpers import have_gpu
from test_leisure import engines

CFG = dict(days_administered_to_effective=[3, 4, 5], days_effective_to_waning=[2, 2, 1], days_waning=[4, 3, 1], waning_factor=0.7,
           sterilisation_efficacies=[{"Covid19": {"0-100": 0.5}}, {"Covid19": {"0-100": 0.9}}, {"Covid19": {"0-100": 0.98}}],
           symptomatic_efficacies=[{"Covid19": {"0-100": 0.6}}, {"Covid19": {"0-100": 1.0}}, {"Covid19": {"0-100": 1.0}}])


def campaigns():
    v = Vaccine.from_config_dict("Test", CFG)
    return VaccinationCampaigns([
        VaccinationCampaign(vaccine=v, days_to_next_dose=[0, 5], dose_numbers=[0, 1], start_time="2020-03-02",
                            end_time="2020-03-07", group_by="age", group_type="50-100", group_coverage=0.8),
        VaccinationCampaign(vaccine=v, days_to_next_dose=[0], dose_numbers=[0], start_time="2020-03-03",
                            end_time="2020-03-09", group_by="primary_activity", group_type="school", group_coverage=0.5),
        VaccinationCampaign(vaccine=v, days_to_next_dose=[0], dose_numbers=[0], start_time="2020-03-04",
                            end_time="2020-03-10", group_by="age", group_type="5-60", group_coverage=0.3),
    ])


def test_daily_probability_and_efficacy_shape_on_the_oracle():
    w = generate_world(60_000, seed=1)
    (eng,), beta = engines(w, ["oracle"])
    camps = campaigns()
    eng.set_vaccination(*camps.device_tables(datetime.date(2020, 3, 2), ["Covid19"]))
    old = (w.age >= 65)  # only the first campaign targets them
    n_vacc = []
    for day in range(5):
        eng.vaccinate(day, seed=4)
        n_vacc.append(int((eng.fetch_vaccination()["dose"][old] >= 0).sum()))
    # c / (T - d c) per day makes the coverage grow linearly to c (vaccination_campaign.py:248-262)
    expect = 0.8 * old.sum()
    assert abs(n_vacc[-1] - expect) < 5 * np.sqrt(expect * 0.2)
    steps = np.diff([0] + n_vacc)
    assert np.all(np.abs(steps - expect / 5) < 6 * np.sqrt(expect / 5))
    # efficacy ramps linearly over days_administered_to_effective, plateaus, then wanes to 0.7 x
    eng2 = engines(w, ["oracle"])[0][0]
    eng2.set_vaccination(*VaccinationCampaigns([camps.vaccination_campaigns[0]]).device_tables(datetime.date(2020, 3, 2), ["Covid19"]))
    eng2.inject_vaccination_uniforms(np.zeros(w.n_people))  # everybody eligible is vaccinated on day 0
    who = int(np.nonzero(old)[0][0])
    susc = []
    for day in range(20):
        eng2.vaccinate(day, seed=1)
        susc.append(eng2.fetch_person_state()["susceptibility"][0, who])
    e = 1 - np.array(susc)
    assert np.allclose(e[:4], [0.0, 0.5 / 3, 1.0 / 3, 0.5]) and e[4] == 0.5   # dose 0: ramp over 3 days, plateau
    # dose 1 on day 5: ramp from 0.5 * 0.7 (previous efficacy x waning factor) to 0.9 over 4 days
    assert np.allclose(e[5:10], 0.35 + (0.9 - 0.35) * np.arange(5) / 4)
    assert np.allclose(e[10:12], 0.9) and np.allclose(e[12:15], 0.9 + (0.63 - 0.9) * np.arange(1, 4) / 3)
    assert np.all(e[15:] == e[14]), "after the trajectory ends the immunity stays where it was"
    assert eng2.fetch_vaccination()["dose"][who] == 1


@pytest.mark.gpu
def test_cuda_matches_oracle_with_vaccination():
    assert have_gpu()
    w = generate_world(60_000, seed=2, sector_betas=load_defaults()["company_sector_betas"])
    (a, b), beta = engines(w, ["cuda", "oracle"])
    tabs = campaigns().device_tables(datetime.date(2020, 3, 2), ["Covid19"])
    seeds = np.sort(np.random.default_rng(2).choice(w.n_people, 3000, replace=False)).astype(np.uint32)
    for e in (a, b):
        e.set_vaccination(*tabs)
        e.infect(seeds, time=-4.0, seed=3, step=999)
    flags = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION
    now, n_frac = 0.0, 0
    for day in range(9):
        for k, (m, dt) in enumerate([(0b10101, 8 / 24), (0b10001, 16 / 24)]):
            i = 2 * day + k
            sa = a.step(a.make_params(now, dt, i, m, seed=9, flags=flags, beta=beta))
            sb = b.step(b.make_params(now, dt, i, m, seed=9, flags=flags, beta=beta))
            for key in ("n_new_infections", "n_infected", "n_draws", "n_hospitalised", "n_dead_total", "n_recovered_step"):
                assert sa[key] == sb[key], (i, key, sa[key], sb[key])
            assert np.array_equal(a.fetch_new_infections()[0], b.fetch_new_infections()[0]), i
            if k == 0:
                for e in (a, b):
                    e.vaccinate(day, seed=9)
            now += dt
        pa, pb = a.fetch_person_state(), b.fetch_person_state()
        va, vb = a.fetch_vaccination(), b.fetch_vaccination()
        assert np.array_equal(pa["pstate"], pb["pstate"]) and np.array_equal(pa["tag"], pb["tag"]), day
        assert np.array_equal(pa["susceptibility"], pb["susceptibility"]), day
        for key in ("dose", "vaccine", "effective_multiplier"):
            assert np.array_equal(va[key], vb[key]), (day, key)
        s = pa["susceptibility"][0]
        n_frac = max(n_frac, int(((s > 0) & (s < 1)).sum()))
    assert n_frac > 10_000, "the fractional-susceptibility path of the interaction kernels must be exercised"
    ia, ib = a.fetch_infections(), b.fetch_infections()
    assert [(r["person"], r["max_tag"]) for r in ia] == [(r["person"], r["max_tag"]) for r in ib]


def test_simulator_records_the_vaccines_table():
    """``vaccines`` record rows (vaccines.py:309-313): one per administered dose, through ``Simulator``."""
    from test_leisure_policies import make_sim
    from june_b200.policy import Policies
    w = generate_world(12_000, seed=3, leisure=True, sector_betas=load_defaults()["company_sector_betas"])
    sim = make_sim("oracle", w, Policies.from_file(), total_days=12, seed=2)
    sim.epidemiology.vaccination_campaigns = campaigns()
    sim.record = object()
    sim.run()
    rows = sim.vaccine_rows()
    assert len(rows) > 3000
    assert {r["vaccine_names"] for r in rows} == {"Test"} and {r["dose_numbers"] for r in rows} == {0, 1}
    vs = sim.device.fetch_vaccination()
    last = {}
    for r in rows:
        assert r["dose_numbers"] == last.get(r["vaccinated_ids"], -1) + 1, "doses come in order, one row each"
        last[r["vaccinated_ids"]] = r["dose_numbers"]
    assert len(last) == int((vs["dose"] >= 0).sum())
    assert all(vs["dose"][p] == d for p, d in last.items())
    second = [r for r in rows if r["dose_numbers"] == 1]
    first_day = {r["vaccinated_ids"]: r["time_stamp"] for r in rows if r["dose_numbers"] == 0}
    assert second and all((r["time_stamp"].date() - first_day[r["vaccinated_ids"]].date()).days == 5 for r in second)
