"""Infection events and infector attribution:
``Interaction._blame_subgroup`` / ``_blame_individuals`` (interaction.py:643-662) and the rows of the
``infections`` record table (interaction.py:664-683)."""
import numpy as np
import pytest

from june_b200 import _lib
from june_b200.interaction import load_defaults
from june_b200.synth import generate_world

from helpers import have_gpu
from test_leisure import MASK_WORK, engines

FLAGS = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION | _lib.STEP_RECORD_EVENTS


def run(eng, w, beta, n_steps, seed=5):
    seeds = np.sort(np.random.default_rng(1).choice(w.n_people, w.n_people // 12, replace=False)).astype(np.uint32)
    eng.infect(seeds, time=-5.0, seed=11, step=1000)
    out, now = [], 0.0
    for i in range(n_steps):
        m, dt = (MASK_WORK & ~0b1000, 8 / 24) if i % 2 == 0 else (0b10001, 16 / 24)
        before = eng.fetch_person_state()
        st = eng.step(eng.make_params(now, dt, i, m, seed=seed, flags=FLAGS, beta=beta))
        ev = eng.fetch_infection_events()
        assert len(ev["infected"]) == st["n_new_infections"]
        out.append((before, ev))
        now += dt
    return out


def test_oracle_blames_present_infectors_in_proportion_to_their_transmission():
    w = generate_world(40_000, seed=2, sector_betas=load_defaults()["company_sector_betas"])
    (eng,), beta = engines(w, ["oracle"])
    grp_of = np.repeat(np.arange(w.n_groups), np.diff(w.grp_offset.astype(np.int64)))
    members = {}
    hh = w.supergroup("households")
    hits = expect = var = 0.0
    n_events = 0
    for before, ev in run(eng, w, beta, 8):
        infected_before = (before["pstate"] & 0x08) != 0
        trans = before["trans_prob"]
        assert np.array_equal(ev["infected"], np.sort(ev["infected"]))
        assert not infected_before[ev["infected"]].any()
        assert infected_before[ev["infector"]].all(), "the infector must carry an infection"
        assert (trans[ev["infector"]] > 0).all()
        n_events += len(ev["infected"])
        for a, b, g in zip(ev["infected"], ev["infector"], ev["group"]):
            lo, hi = w.grp_offset[g], w.grp_offset[g + 1]
            mem = w.reg_member[lo:hi]
            for q in (a, b):  # both sit in the group of the event: registered there, or its patients
                assert q in mem or (before["medical"][q] >= 0 and before["medical"][q] >> 8 == g), (q, g)
            if hh.first_group <= g < hh.first_group + hh.n_groups:
                cand = [q for q in mem if infected_before[q] and trans[q] > 0]
                if len(cand) == 2 and np.ptp(w.reg_info[lo:hi][np.isin(mem, cand)] & 0xFF) == 0:
                    # two infectors in the same household subgroup: P(first) = t0 / (t0 + t1)
                    p0 = trans[cand[0]] / (trans[cand[0]] + trans[cand[1]])
                    hits += b == cand[0]
                    expect += p0
                    var += p0 * (1 - p0)
    assert n_events > 500
    assert var > 3 and abs(hits - expect) < 4.5 * np.sqrt(var), (hits, expect, var)


@pytest.mark.gpu
def test_cuda_events_equal_the_oracles():
    assert have_gpu()
    w = generate_world(60_000, seed=3, sector_betas=load_defaults()["company_sector_betas"])
    (a, b), beta = engines(w, ["cuda", "oracle"])
    total = 0
    for (_, ea), (_, eb) in zip(run(a, w, beta, 8), run(b, w, beta, 8)):
        for key in ("infected", "infector", "group", "variant"):
            assert np.array_equal(ea[key], eb[key]), key
        total += len(ea["infected"])
    assert total > 1000
