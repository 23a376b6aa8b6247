"""``CloseCompaniesLockdownTiers`` (individual_policies.py:551-601): company workers with lockdown status
"random" who work outside their home super area skip work with probability ``regional_compliance`` when the
work or the home region is in lockdown tier 3 / 4."""
import datetime

import numpy as np
import pytest

from june_b200 import _lib
from june_b200.interaction import load_defaults
from june_b200.policy import CloseCompaniesLockdownTiers, Policies, TieredLockdown
from june_b200.synth import generate_world

from helpers import have_gpu
from test_leisure import MASK_WORK, engines

FLAGS = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION


def world_with_out_of_area_workers(n=40_000):
    w = generate_world(n, seed=11, sector_betas=load_defaults()["company_sector_betas"])
    rng = np.random.default_rng(4)
    kind, lock = (w.person_attr >> 2) & 7, w.person_attr & 3
    cand = np.nonzero((kind == 2) & (lock == 2))[0]  # company workers with lockdown status "random"
    chosen = rng.choice(cand, len(cand) // 2, replace=False)
    wr = np.full(w.n_people, 0xFF, np.uint8)
    wr[chosen] = rng.integers(0, w.n_regions, len(chosen))
    # some people who do NOT qualify carry a work region too: key workers, pupils
    others = rng.choice(np.nonzero((kind != 2) | (lock != 2))[0], 2000, replace=False)
    wr[others] = rng.integers(0, w.n_regions, len(others))
    w.person_work_region = wr
    return w, chosen


def run(kinds):
    w, chosen = world_with_out_of_area_workers()
    engs, beta = engines(w, kinds)
    tier_regions = [1, 4]
    rec = [dict(type=_lib.POL_CLOSE_COMPANIES_TIERS, mask=sum(1 << r for r in tier_regions))]
    rc = np.linspace(0.3, 0.9, w.n_regions)
    out = []
    for e in engs:
        e.set_individual_policies(rec, rc)
        states = []
        now = 0.0
        for i in range(6):
            e.step(e.make_params(now, 8 / 24, i, MASK_WORK & ~0b1000, seed=8, flags=FLAGS, beta=beta))
            states.append(e.fetch_person_state()["pstate"].copy())
            now += 8 / 24
        out.append(states)
    return w, chosen, tier_regions, rc, out


def check(w, chosen, tier_regions, rc, states):
    home = w.sa_region[w.super_area]
    wr = w.person_work_region
    kind, lock = (w.person_attr >> 2) & 7, w.person_attr & 3
    subject = (kind == 2) & (lock == 2) & (wr != 0xFF)
    in_tier = np.isin(home, tier_regions) | (np.isin(wr, tier_regions) & (wr != 0xFF))
    at_work = np.stack([(s & 7) == 2 for s in states])             # [step][person]
    workers = (w.has_activity & 0b100) != 0
    untouched = workers & ~(subject & in_tier)
    assert at_work[:, untouched].all(), "nobody outside the policy's subject group stays away from work"
    for r in range(w.n_regions):
        sel = subject & in_tier & (home == r)
        if sel.sum() < 150:
            continue
        skipped = 1.0 - at_work[:, sel].mean()
        n = sel.sum() * len(states)
        assert abs(skipped - rc[r]) < 4.5 * np.sqrt(rc[r] * (1 - rc[r]) / n), (r, skipped, rc[r])
    assert (subject & in_tier).sum() > 1500


def test_oracle_tiered_company_closure():
    w, chosen, tiers, rc, out = run(["oracle"])
    check(w, chosen, tiers, rc, out[0])


def test_policy_class_builds_the_region_mask_from_the_tiers():
    p = CloseCompaniesLockdownTiers("2020-01-01", "2021-01-01")
    rec = p.device_record(dict(lockdown_tiers=[None, 3, 2, 4, 1]))
    assert rec["type"] == _lib.POL_CLOSE_COMPANIES_TIERS and rec["mask"] == (1 << 1) | (1 << 3)
    pols = Policies([p, TieredLockdown("2020-01-01", "2021-01-01", tiers_per_region={"a": 4, "b": 1})])
    # default: what the running reference does -- its do_timestep never applies the tiers (simulator.py:359-366)
    assert pols.regional_scalars(datetime.datetime(2020, 6, 1), ["a", "b"], {})[1] == [None, None]
    pols.apply_tiered_lockdown = True
    comp, tiers = pols.regional_scalars(datetime.datetime(2020, 6, 1), ["a", "b"], {})
    assert tiers == [4, 1]
    recs = pols.device_policies(datetime.datetime(2020, 6, 1), dict(stay_at_home_mask=0, lockdown_tiers=tiers))
    assert recs == [dict(type=_lib.POL_CLOSE_COMPANIES_TIERS, mask=1)]


@pytest.mark.gpu
def test_cuda_tiered_company_closure_equals_the_oracle():
    assert have_gpu()
    w, chosen, tiers, rc, out = run(["cuda", "oracle"])
    for sa, sb in zip(*out):
        assert np.array_equal(sa, sb)
    check(w, chosen, tiers, rc, out[0])
