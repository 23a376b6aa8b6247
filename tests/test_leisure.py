"""Leisure: probability tables against the reference's own formulae,
the placement sampler against those tables (oracle, CPU), and CUDA == oracle on leisure worlds.

Reference: june/groups/leisure/leisure.py:205-372, social_venue_distributor.py:143-205,257-278.
"""
import datetime

import numpy as np
import pytest

from june_b200 import _lib
from june_b200.disease import DiseaseConfig, outcome_probabilities, synthetic_rates
from june_b200.interaction import Interaction, load_defaults
from june_b200.leisure import Leisure, PubDistributor, generate_leisure_for_world, parse_age_probabilities
from june_b200.synth import generate_world
from june_b200.world import ACT_LEISURE, ACT_RESIDENCE

from helpers import default_beta_table, have_gpu, make_engine

MASK_LEISURE = 0b11001   # medical | leisure | residence
MASK_WORK = 0b11101      # medical | primary | leisure | residence


def engines(world, kinds):
    inter = Interaction.from_file()
    dc = DiseaseConfig("covid19")
    rates, bins = synthetic_rates()
    prob = outcome_probabilities(dc, rates, bins)
    return [make_engine(k, world, prob, interaction=inter, disease=dc) for k in kinds], default_beta_table(world, inter)


def tables_for(world, when=datetime.datetime(2020, 3, 7, 12), dt=4 / 24, working_hours=False):
    leisure = generate_leisure_for_world(["pubs", "cinemas", "groceries", "gyms"], world)
    assert leisure.activities == world.leisure_activities
    return leisure.generate_leisure_probabilities_for_timestep(dt, working_hours, when)


def test_age_bins_are_half_open():
    v = parse_age_probabilities({"0-9": 1.0, "9-15": 2.0, "19-31": 3.0})
    assert v[0] == 1.0 and v[8] == 1.0 and v[9] == 2.0 and v[14] == 2.0 and v[15] == 0.0 and v[19] == 3.0 and v[31] == 0.0


def test_poisson_parameters_and_probabilities_follow_the_reference_formulae():
    d = PubDistributor.from_config()
    cfg = load_defaults()["leisure"]["pubs"]
    # social_venue_distributor.py:143-149: times_per_week / n_days * 24 / hours_per_day
    assert d.poisson_parameters["weekday"]["m"][25] == pytest.approx(cfg["times_per_week"]["weekday"]["male"]["19-31"] / 5 * 24 / 3)
    assert d.poisson_parameters["weekend"]["f"][70] == pytest.approx(cfg["times_per_week"]["weekend"]["female"]["66-86"] / 2 * 24 / 12)
    leisure = Leisure({"pub": d}, ["A", "B"])
    leisure.regional_compliance = {"A": 1.0, "B": 0.5}
    leisure.policy_reductions = {"pub": {dt: {s: [0.2] * 100 for s in "mf"} for dt in ("weekday", "weekend")}}
    does, cum = leisure.generate_leisure_probabilities_for_timestep(3 / 24, False, datetime.datetime(2020, 3, 2, 18))
    lam = d.poisson_parameters["weekday"]["m"][25]
    # leisure.py:338 with social_venue_distributor.py:199-204: lambda * (1 + compliance * (reduction - 1))
    assert does[0, 1, 25] == pytest.approx(1 - np.exp(-3 / 24 * lam * 0.2))
    assert does[1, 1, 25] == pytest.approx(1 - np.exp(-3 / 24 * lam * (1 + 0.5 * (0.2 - 1))))
    assert np.all(cum[..., -1] == 1.0)


def test_oracle_sampler_follows_the_tables():
    """Attendance produced by the placement stage vs the table it samples from: overall rate, the
    split over activities (chi-square), venues only from the person's area list."""
    w = generate_world(120_000, seed=5, leisure=True)
    (eng,), beta = engines(w, ["oracle"])
    does, cum = tables_for(w)
    eng.set_leisure_tables(does[None], cum[None])
    st = eng.step(eng.make_params(0.0, 4 / 24, 0, MASK_LEISURE, seed=9, flags=0, beta=beta, leisure_table=1))
    act = eng.fetch_person_state()["pstate"] & 7
    elig = w.care_home_resident == 0
    region = w.sa_region[w.super_area]
    p = does[region, w.sex, np.minimum(w.age, 99)]
    went = act == ACT_LEISURE
    assert not went[~elig].any(), "care-home residents never do leisure (leisure.py:246-248)"
    n_exp, n_var = p[elig].sum(), (p[elig] * (1 - p[elig])).sum()
    assert abs(went.sum() - n_exp) < 5 * np.sqrt(n_var), (went.sum(), n_exp)
    assert st["n_by_activity"][ACT_LEISURE] == went.sum()
    assert st["n_by_activity"][ACT_RESIDENCE] + went.sum() + st["n_by_activity"][0] == w.n_people
    # with primary activity in the step, people who have one do not draw leisure at all
    st2 = eng.step(eng.make_params(1.0, 8 / 24, 1, MASK_WORK, seed=9, flags=0, beta=beta, leisure_table=1))
    act2 = eng.fetch_person_state()["pstate"] & 7
    has_primary = (w.has_activity & 0b100) != 0
    assert not (act2[has_primary] == ACT_LEISURE).any()
    assert (act2[~has_primary] == ACT_LEISURE).sum() == st2["n_by_activity"][ACT_LEISURE] > 0


@pytest.mark.gpu
def test_cuda_matches_oracle_on_a_leisure_world():
    assert have_gpu()
    w = generate_world(60_000, seed=4, sector_betas=load_defaults()["company_sector_betas"], leisure=True)
    (a, b), beta = engines(w, ["cuda", "oracle"])
    slots = [tables_for(w, datetime.datetime(2020, 3, 7, 12), 4 / 24), tables_for(w, datetime.datetime(2020, 3, 2, 9), 8 / 24, True),
             tables_for(w, datetime.datetime(2020, 3, 2, 18), 3 / 24)]
    for e in (a, b):
        e.set_leisure_tables(np.stack([s[0] for s in slots]), np.stack([s[1] for s in slots]))
    seeds = np.sort(np.random.default_rng(2).choice(w.n_people, 2500, replace=False)).astype(np.uint32)
    for e in (a, b):
        e.infect(seeds, time=-4.0, seed=3, step=999)
    sched = [(MASK_LEISURE, 4 / 24, 1), (MASK_WORK, 8 / 24, 2), (0b10001, 1 / 24, 0), (MASK_LEISURE, 3 / 24, 3), (0b10001, 11 / 24, 0)]
    now, went, new_in_venues = 0.0, 0, 0
    flags = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION
    for i in range(20):
        m, dt, tab = sched[i % 5]
        record = _lib.STEP_RECORD_LAMBDA if i % 5 == 0 else 0  # plain launches and graph replays both covered
        sa = a.step(a.make_params(now, dt, i, m, seed=21, flags=flags | record, beta=beta, leisure_table=tab))
        sb = b.step(b.make_params(now, dt, i, m, seed=21, flags=flags | record, beta=beta, leisure_table=tab))
        for key in ("n_new_infections", "n_infected", "n_draws", "n_by_activity", "n_hospitalised", "n_dead_total"):
            assert sa[key] == sb[key], (i, key, sa[key], sb[key])
        ma, mb = a.fetch_new_infections()[0], b.fetch_new_infections()[0]
        assert np.array_equal(ma, mb), f"step {i}: infected sets differ"
        pa, pb = a.fetch_person_state(), b.fetch_person_state()
        assert np.array_equal(pa["pstate"], pb["pstate"]), f"step {i}: placement / flags differ"
        if record:
            la, lb = a.fetch_lambda(), b.fetch_lambda()
            assert np.array_equal(np.isnan(la), np.isnan(lb))
            ok = ~np.isnan(la)
            np.testing.assert_allclose(la[ok], lb[ok], rtol=1e-9, atol=0)
        at_leisure = (pa["pstate"] & 7) == ACT_LEISURE
        went += int(at_leisure.sum())
        new_in_venues += int(at_leisure[ma].sum())
        now += dt
    assert went > 10_000 and new_in_venues > 0


@pytest.mark.gpu
def test_injected_leisure_placement_is_honoured_by_both_sides():
    """Test mode: venues handed in per person (how a trace of the reference's own placement is
    replayed) -- the venue's interaction must then agree, draw for draw."""
    assert have_gpu()
    w = generate_world(30_000, seed=6, leisure=True)
    (a, b), beta = engines(w, ["cuda", "oracle"])
    rng = np.random.default_rng(3)
    pubs = w.supergroup("pubs")
    gyms = w.supergroup("gyms")
    venue = np.full(w.n_people, -1, np.int32)
    goers = rng.choice(w.n_people, 9000, replace=False)
    venue[goers[:6000]] = pubs.first_group + rng.integers(0, min(4, pubs.n_groups), 6000)   # crowded venues
    venue[goers[6000:]] = gyms.first_group + rng.integers(0, gyms.n_groups, 3000)
    seeds = np.sort(rng.choice(w.n_people, 1500, replace=False)).astype(np.uint32)
    for e in (a, b):
        e.infect(seeds, time=-4.0, seed=3, step=999)
        e.inject_leisure(venue)
    flags = _lib.STEP_INJECT_LEISURE | _lib.STEP_RECORD_LAMBDA
    sa = a.step(a.make_params(0.0, 4 / 24, 0, MASK_LEISURE, seed=1, flags=flags, beta=beta))
    sb = b.step(b.make_params(0.0, 4 / 24, 0, MASK_LEISURE, seed=1, flags=flags, beta=beta))
    for key in ("n_new_infections", "n_infected", "n_draws", "n_by_activity"):
        assert sa[key] == sb[key], (key, sa[key], sb[key])
    act = a.fetch_person_state()["pstate"] & 7
    expect = (venue >= 0) & (w.care_home_resident == 0)
    assert np.array_equal(act == ACT_LEISURE, expect)
    la, lb = a.fetch_lambda(), b.fetch_lambda()
    assert np.array_equal(np.isnan(la), np.isnan(lb))
    ok = ~np.isnan(la)
    np.testing.assert_allclose(la[ok], lb[ok], rtol=1e-9, atol=0)
    assert np.array_equal(a.fetch_new_infections()[0], b.fetch_new_infections()[0])
    assert (~np.isnan(la[expect])).sum() > 1000, "venue visitors must have drawn"


@pytest.mark.gpu
def test_cuda_matches_oracle_with_lockdown_policies_and_leisure():
    """Individual policies drawing from their Philox streams: both sides must
    place every person identically, step after step, with leisure on top."""
    assert have_gpu()
    from june_b200.policy import CloseCompanies
    w = generate_world(50_000, seed=9, sector_betas=load_defaults()["company_sector_betas"], leisure=True)
    (a, b), beta = engines(w, ["cuda", "oracle"])
    CloseCompanies.initialize(w)
    f, k, r = CloseCompanies.furlough_ratio, CloseCompanies.key_ratio, CloseCompanies.random_ratio
    assert f is not None
    recs = [dict(type=_lib.POL_QUARANTINE, mask=(1 << 4) | (1 << 5), p=[7, 0.8]),
            dict(type=_lib.POL_SHIELDING, p=[70, 0.6]),
            dict(type=_lib.POL_CLOSE_SCHOOLS, mask=0b11100000, p=[0.7]),
            dict(type=_lib.POL_CLOSE_COMPANIES, p=[0.3, 2 * f, 0.5 * k, f, k, r])]
    rc = np.linspace(0.8, 1.1, w.n_regions)
    does, cum = tables_for(w)
    seeds = np.sort(np.random.default_rng(2).choice(w.n_people, 2500, replace=False)).astype(np.uint32)
    for e in (a, b):
        e.set_leisure_tables(does[None], cum[None])
        e.set_individual_policies(recs, rc)
        e.infect(seeds, time=-3.0, seed=3, step=999)
    now = 0.5
    stayed = 0
    for i in range(10):
        m, dt = (MASK_WORK, 8 / 24) if i % 2 == 0 else (MASK_LEISURE, 3 / 24)
        flags = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION
        sa = a.step(a.make_params(now, dt, i, m, seed=4, flags=flags, beta=beta, leisure_table=1))
        sb = b.step(b.make_params(now, dt, i, m, seed=4, flags=flags, beta=beta, leisure_table=1))
        for key in ("n_new_infections", "n_infected", "n_draws", "n_by_activity", "n_hospitalised", "n_dead_total",
                    "n_recovered_step", "n_no_activity"):
            assert sa[key] == sb[key], (i, key, sa[key], sb[key])
        pa, pb = a.fetch_person_state(), b.fetch_person_state()
        assert np.array_equal(pa["pstate"], pb["pstate"]), f"step {i}"
        if i % 2 == 0:
            has_primary = (w.has_activity & 0b100) != 0
            stayed += int((((pa["pstate"] & 7) != 2) & has_primary & ((pa["pstate"] & 0x10) == 0)).sum())
        now += dt
    assert stayed > 20_000, "closures and shielding must keep people away from their primary activity"


@pytest.mark.gpu
def test_limit_long_commute_alone_uses_the_subject_list_and_equals_the_oracle():
    """The default policy file keeps ``limit_long_commute`` always on (policy_covid19.yaml:129-135): the device
    places everyone with the table-free fast path and re-evaluates only the long-distance commuters."""
    assert have_gpu()
    from june_b200.world import PA_LONG_COMMUTER
    w = generate_world(50_000, seed=9, sector_betas=load_defaults()["company_sector_betas"])
    rng = np.random.default_rng(3)
    workers = np.nonzero((w.has_activity & 0b100) != 0)[0]
    long_c = rng.choice(workers, 4000, replace=False)
    w.person_attr = w.person_attr.copy() if w.person_attr is not None else np.zeros(w.n_people, np.uint8)
    w.person_attr[long_c] |= PA_LONG_COMMUTER
    (a, b), beta = engines(w, ["cuda", "oracle"])
    recs = [dict(type=_lib.POL_LIMIT_LONG_COMMUTE, p=[0.3])]
    seeds = np.sort(rng.choice(w.n_people, 2500, replace=False)).astype(np.uint32)
    for e in (a, b):
        e.set_individual_policies(recs, np.ones(w.n_regions))
        e.infect(seeds, time=-6.0, seed=3, step=999)
    now, skipped = 0.5, 0
    launches0 = a.launch_count() if hasattr(a, "launch_count") else None
    for i in range(8):
        m, dt = (MASK_WORK & ~0b1000, 8 / 24) if i % 2 == 0 else (0b10001, 16 / 24)
        flags = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION
        sa = a.step(a.make_params(now, dt, i, m, seed=4, flags=flags, beta=beta))
        sb = b.step(b.make_params(now, dt, i, m, seed=4, flags=flags, beta=beta))
        for key in ("n_new_infections", "n_infected", "n_draws", "n_by_activity", "n_no_activity"):
            assert sa[key] == sb[key], (i, key, sa[key], sb[key])
        pa, pb = a.fetch_person_state(), b.fetch_person_state()
        assert np.array_equal(pa["pstate"], pb["pstate"]), f"step {i}"
        if i % 2 == 0:
            act = pa["pstate"][long_c] & 7
            alive = (pa["pstate"][long_c] & 0x10) == 0
            skipped += int(((act != 2) & alive).sum())
            others = np.setdiff1d(workers, long_c)
            healthy = (pa["pstate"][others] & 0x78) == 0
            assert ((pa["pstate"][others] & 7)[healthy] == 2).all(), "nobody else is touched by the policy"
        now += dt
    # as in the reference, the activity is skipped when random() < going_to_work_probability (0.3)
    assert 0.25 * 4 * 4000 < skipped < 0.36 * 4 * 4000, skipped
