"""The counter-based RNG and the new-infection sampler of the oracle.

* Philox4x32-10 known-answer vectors (Random123 kat_vectors: zero counter/key, all-ones, pi digits).
* The completion-time / transmission-parameter samplers restate scipy.stats distributions
  (june/epidemiology/infection/trajectory_maker.py:95-132): Kolmogorov-Smirnov tests against
  scipy's CDFs, and a chi-square test of the outcome (max tag) frequencies against the
  health-index table.  Parity with the reference is STATISTICAL here by construction (the
  reference draws from numpy's global Mersenne Twister through scipy).
"""
import ctypes

import numpy as np
import pytest
from scipy import stats

from june_b200.disease import DiseaseConfig, outcome_probabilities, synthetic_rates
from june_b200.interaction import Interaction
from june_b200.synth import generate_world

from helpers import make_engine, oracle_lib


def philox(ctr, key):
    """Pure-python Philox4x32-10 (Salmon et al. SC'11) used to pin the C implementations."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c, k = list(ctr), list(key)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


def test_philox_known_answers():
    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


@pytest.fixture(scope="module")
def infected_world():
    w = generate_world(120_000, seed=5)
    dc = DiseaseConfig("covid19")
    rates, bins = synthetic_rates()
    prob = outcome_probabilities(dc, rates, bins)
    eng = make_engine("oracle", w, prob, disease=dc)
    eng.infect(np.arange(w.n_people, dtype=np.uint32), time=0.0, seed=99, step=7)
    rows = eng.fetch_infections()
    eng.close()
    return w, dc, prob, rows


def test_bernoulli_uniform_follows_the_spec(infected_world):
    # first uniform of stream 16 (max_severity) must be u52(x0, x1) of Philox(gid, step, 16, 0; seed)
    w, dc, prob, rows = infected_world
    cum = np.cumsum(prob, axis=-1)
    for r in rows[:200]:
        p = r["person"]
        x = philox([p, 7, 16, 0], [99, 0])
        u = (((x[0] << 32 | x[1]) >> 12) + 0.5) * 2.0 ** -52
        pop = 1 if (w.care_home_resident[p] and w.age[p] >= 50) else 0
        c = cum[pop, w.sex[p], min(int(w.age[p]), 99)]
        idx = min(int(np.searchsorted(c, u, side="left")), len(c) - 1)
        assert r["max_tag"] == idx


def ks(sample, cdf, alpha=1e-4):
    d, p = stats.kstest(sample, cdf)
    assert p > alpha, (d, p, len(sample))


def test_stage_times_follow_scipy_distributions(infected_world):
    w, dc, prob, rows = infected_world
    table = {mt: stages for mt, stages in dc.trajectory_table()}
    by_tag = {}
    for r in rows:
        by_tag.setdefault(r["max_tag"], []).append(r)
    checked = 0
    for tag, rs in by_tag.items():
        stages = table[tag]
        tt = np.array([r["traj_time"] for r in rs])
        assert tt.shape[1] == len(stages)
        assert [int(x) for x in rs[0]["traj_tag"]] == [s[0] for s in stages]
        dur = np.diff(tt, axis=1)  # duration of stage s = start(s+1) - start(s)
        for s in range(len(stages) - 1):
            _, code, p = stages[s]
            x = dur[:, s]
            if len(x) < 300:
                continue
            if code == 0:
                np.testing.assert_allclose(x, p[0], rtol=0, atol=1e-9)
            elif code == 2:
                ks(x, stats.beta(p[0], p[1], loc=p[2], scale=p[3]).cdf)
            elif code == 3:
                ks(x, stats.lognorm(p[0], loc=p[1], scale=p[2]).cdf)
            elif code == 5:
                ks(x, stats.exponweib(p[0], p[1], loc=p[2], scale=p[3]).cdf)
            checked += 1
    assert checked >= 12


def test_transmission_parameters_follow_config(infected_world):
    w, dc, prob, rows = infected_world
    par = np.array([r["tpar"] for r in rows])
    t_exposed = np.array([r["traj_time"][1] for r in rows])
    ks(par[:, 0], stats.norm(1.56, 0.08).cdf)                    # shape
    ks(1.0 / par[:, 2], stats.norm(0.53, 0.03).cdf)              # rate = 1/scale
    ks(par[:, 1] - t_exposed, stats.norm(-2.12, 0.1).cdf)        # shift - time_exposed
    ks(par[:, 3], stats.lognorm(0.5, loc=0.0, scale=1.0).cdf)    # max_infectiousness (factors never applied)


def test_outcome_frequencies_follow_health_index(infected_world):
    w, dc, prob, rows = infected_world
    person = np.array([r["person"] for r in rows])
    tag = np.array([r["max_tag"] for r in rows])
    sel = (w.age[person] >= 60) & (w.age[person] <= 69) & (w.sex[person] == 1) & (w.care_home_resident[person] == 0)
    obs = np.bincount(tag[sel], minlength=9).astype(float)
    exp = prob[0, 1, 65] * sel.sum()
    keep = exp > 5
    chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
    assert chi2 < stats.chi2.ppf(1 - 1e-4, keep.sum() - 1), (chi2, obs, exp)
