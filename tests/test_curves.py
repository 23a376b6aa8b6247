"""Statistical parity of the daily curves (third parity requirement of BASELINE.json): 32 seeds of
the device path against 32 seeds of the UNMODIFIED reference on the same world
(tests/golden/make_curves.py -> curves_ref.npz).  Two-sample tests on summary statistics of the
daily infection / hospitalisation / death curves.  Tolerance: Kolmogorov-Smirnov p > 1e-3 and
means within 4 combined standard errors."""
import numpy as np
import pytest
from scipy import stats

from june_b200.disease import DiseaseConfig, outcome_probabilities, synthetic_rates
from june_b200.interaction import Interaction
from june_b200.policy import Hospitalisation, Policies, SevereSymptomsStayHome
from june_b200.simulator import ActivityManager, Epidemiology, InfectionSeed, Simulator
from june_b200.timer import Timer

from helpers import GOLDEN, Golden, have_gpu, make_engine, simulator_over

import os


def run_seeds(kind, n_seeds=32):
    ref = np.load(os.path.join(GOLDEN, "curves_ref.npz"))
    days, cases = int(ref["days"]), float(ref["cases"])
    world = Golden("w3k").world
    assert world.n_people == int(ref["n_people"])
    inter = Interaction.from_file()
    dc = DiseaseConfig("covid19")
    rates, bins = synthetic_rates()
    prob = outcome_probabilities(dc, rates, bins)
    out = np.zeros((n_seeds, days, 4), np.int64)
    for s in range(n_seeds):
        eng = make_engine(kind, world, prob, interaction=inter, disease=dc)
        # tests/golden/ref_world.py: make_simulator(schedule="default")
        timer = Timer(initial_day="2020-03-02 0:00", total_days=days, weekday_step_duration=[1, 8, 1, 3, 11],
                      weekend_step_duration=[12, 12],
                      weekday_activities=[["medical_facility", "residence"], ["medical_facility", "primary_activity", "residence"],
                                          ["medical_facility", "residence"], ["medical_facility", "residence"],
                                          ["medical_facility", "residence"]],
                      weekend_activities=[["medical_facility", "residence"], ["medical_facility", "residence"]])
        policies = Policies([Hospitalisation(dc), SevereSymptomsStayHome()])
        am = ActivityManager(world, policies, timer, ["medical_facility", "primary_activity", "residence"],
                             {"medical_facility": ["hospitals"], "primary_activity": ["schools", "companies"],
                              "residence": ["households", "care_homes"]})
        epi = Epidemiology(infection_seeds=[InfectionSeed(cases, "2020-03-02", seed=s)])
        sim = simulator_over(eng)(world, inter, timer, am, epi, disease_config=dc, outcome_prob=prob, seed=1000 + s)
        sim.run()
        seeded = int(round(cases * world.n_people))
        for h in sim.history:
            d = int(h["now"])
            out[s, d, 0] += h["n_new_infections"]
            out[s, d, 1] = h["n_infected"]
            out[s, d, 2] = h["n_hospitalised"]
            out[s, d, 3] = h["n_dead_total"]
        out[s, 0, 0] += seeded
        eng.close()
    return out, ref["curves"]


def summaries(c):
    return {
        "total infections": c[:, :, 0].sum(axis=1),
        "infections in days 0-9": c[:, :10, 0].sum(axis=1),
        "infections in days 10-24": c[:, 10:, 0].sum(axis=1),
        "currently infected, day 12": c[:, 12, 1],
        "currently infected, day 24": c[:, 24, 1],
        "peak in hospital": c[:, :, 2].max(axis=1),
        "hospital bed-days": c[:, :, 2].sum(axis=1),
        "deaths by day 24": c[:, 24, 3],
    }


def check(ours, ref):
    a, b = summaries(ours), summaries(ref)
    report = []
    for k in a:
        x, y = a[k].astype(float), b[k].astype(float)
        p = stats.ks_2samp(x, y).pvalue
        se = np.sqrt(x.var(ddof=1) / len(x) + y.var(ddof=1) / len(y))
        z = (x.mean() - y.mean()) / max(se, 1e-12)
        report.append((k, x.mean(), y.mean(), z, p))
    for k, mx, my, z, p in report:
        print(f"{k:32s} ours {mx:8.2f}  reference {my:8.2f}  z {z:+5.2f}  KS p {p:.3f}")
    for k, mx, my, z, p in report:
        assert p > 1e-3, f"{k}: KS p = {p:.2e}"
        assert abs(z) < 4.0, f"{k}: means differ by {z:.1f} standard errors"


def test_oracle_curves_match_reference_over_32_seeds():
    ours, ref = run_seeds("oracle")
    check(ours, ref)


@pytest.mark.gpu
def test_cuda_curves_match_reference_over_32_seeds():
    assert have_gpu()
    ours, ref = run_seeds("cuda")
    check(ours, ref)
