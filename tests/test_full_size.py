"""BASELINE.json configs[1] at its full size (2.7 M people) on the GPU: parity with the CPU oracle over a
simulated day and a half, run-to-run determinism (results must not depend on atomic ordering), and
the size-independent invariants of the path: everybody alive is placed exactly once, the per-region
summary adds up to the totals, infected people are immune, counters equal flag counts."""
import numpy as np
import pytest

from june_b200 import _lib
from june_b200.interaction import load_defaults
from june_b200.synth import generate_world
from june_b200.world import PS_DEAD, PS_INFECTED

from helpers import have_gpu
from test_leisure import engines

pytestmark = pytest.mark.gpu
N = 2_700_000
SCHED = [(0b10011, 1 / 24), (0b10101, 8 / 24), (0b10011, 1 / 24), (0b10001, 3 / 24), (0b10001, 11 / 24)]


def run(eng, w, beta, n_steps, events=False):
    seeds = np.sort(np.random.default_rng(5).choice(w.n_people, w.n_people // 40, replace=False)).astype(np.uint32)
    for k, part in enumerate(np.array_split(seeds, 6)):
        eng.infect(part, time=-6.0 + k, seed=7, step=5000 + k)
    flags = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION | (_lib.STEP_RECORD_EVENTS if events else 0)
    now, out = 0.0, []
    for i in range(n_steps):
        m, dt = SCHED[i % 5]
        st = eng.step(eng.make_params(now, dt, i, m, seed=123, flags=flags, beta=beta))
        out.append((st, eng.fetch_new_infections()[0], eng.fetch_summary()))
        now += dt
    return out


def test_full_size_parity_determinism_and_invariants():
    assert have_gpu()
    w = generate_world(N, seed=0, sector_betas=load_defaults()["company_sector_betas"], commute=True)
    (a, a2, b), beta = engines(w, ["cuda", "cuda", "oracle"])
    ra, ra2, rb = run(a, w, beta, 8), run(a2, w, beta, 8), run(b, w, beta, 8)
    total_new = dead_before = 0
    for i, ((sa, ma, ga), (sa2, ma2, ga2), (sb, mb, gb)) in enumerate(zip(ra, ra2, rb)):
        # determinism: two CUDA runs
        assert sa == sa2 and np.array_equal(ma, ma2) and np.array_equal(ga, ga2), f"step {i}: run-to-run difference"
        # parity with the oracle
        for key in ("n_new_infections", "n_infected", "n_dead_total", "n_recovered_step", "n_hospitalised", "n_icu",
                    "n_draws", "n_by_activity", "n_no_activity"):
            assert sa[key] == sb[key], (i, key, sa[key], sb[key])
        assert np.array_equal(ma, mb), f"step {i}: infected sets differ"
        assert np.array_equal(ga, gb), f"step {i}: per-region summary differs"
        # invariants
        assert sum(sa["n_by_activity"]) + dead_before == w.n_people, "everybody alive is placed exactly once"
        dead_before = sa["n_dead_total"]  # placement precedes the step's health update
        assert sa["n_no_activity"] == 0
        # (the summary's ward / intensive-care columns count the people PLACED there by this step, like the reference's
        # len(hospital.ward) / len(hospital.icu): records_writer.py:290-294)
        assert ga[:, 1].sum() == sa["n_infected"] and ga[:, 2].sum() + ga[:, 3].sum() == sa["n_by_activity"][0]
        assert ga[:, 0].sum() == sa["n_new_infections"]
        total_new += len(ma)
    assert total_new > 20_000
    pa, pb = a.fetch_person_state(), b.fetch_person_state()
    assert np.array_equal(pa["pstate"], pb["pstate"]) and np.array_equal(pa["tag"], pb["tag"])
    assert np.array_equal(pa["medical"], pb["medical"])
    np.testing.assert_allclose(pa["trans_prob"], pb["trans_prob"], rtol=1e-9, atol=0)
    infected = (pa["pstate"] & PS_INFECTED) != 0
    assert int(infected.sum()) == ra[-1][0]["n_infected"]
    assert int(((pa["pstate"] & PS_DEAD) != 0).sum()) == ra[-1][0]["n_dead_total"]
    assert np.all(pa["susceptibility"][0][infected] == 0.0), "infected people are immune (infection_selector.py:160-162)"
    assert np.all(pa["trans_prob"][~infected] == 0.0)
