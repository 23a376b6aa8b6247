"""Commute: Travel.get_commute_subgroup (travel.py:125-131) ->
City.get_commute_subgroup (city.py:83-102) -> Station.get_commute_subgroup (station.py:52-76)."""
import numpy as np
import pytest

from june_b200 import _lib
from june_b200.interaction import load_defaults
from june_b200.synth import generate_world
from june_b200.world import ACT_COMMUTE

from helpers import have_gpu
from test_leisure import engines

MASK_COMMUTE = 0b10011  # medical | commute | residence


def test_oracle_places_commuters_on_their_stations_transports():
    w = generate_world(150_000, seed=7, commute=True)
    (eng,), beta = engines(w, ["oracle"])
    flags = _lib.STEP_RECORD_EVENTS
    seeds = np.sort(np.random.default_rng(0).choice(w.n_people, 8000, replace=False)).astype(np.uint32)
    eng.infect(seeds, time=-4.0, seed=1, step=99)
    eng.step(eng.make_params(-0.5, 12 / 24, 7, 0b10001, seed=3, flags=0, beta=beta))  # first health update: probabilities > 0
    st = eng.step(eng.make_params(0.0, 12 / 24, 0, MASK_COMMUTE, seed=3, flags=flags, beta=beta))
    act = eng.fetch_person_state()["pstate"] & 7
    commuter = w.person_commute >= 0
    assert np.array_equal(act == ACT_COMMUTE, commuter), "exactly the public-transport commuters travel"
    assert st["n_by_activity"][ACT_COMMUTE] == commuter.sum() > 10_000
    # infections on transports happen between people who can share one: same station (inter-city) or same city
    ev = eng.fetch_infection_events()
    ct, it = w.supergroup("city_transports"), w.supergroup("inter_city_transports")
    on_transport = (ev["group"] >= ct.first_group) & (ev["group"] < it.first_group + it.n_groups)
    assert on_transport.sum() > 20
    tr_station = np.repeat(np.arange(len(w.station_transport_off) - 1), np.diff(w.station_transport_off.astype(np.int64)))
    station_of_group = dict(zip(w.station_transport.tolist(), tr_station.tolist()))
    st_city = np.repeat(np.arange(len(w.city_station_off) - 1), np.diff(w.city_station_off.astype(np.int64)))
    city_of_station = dict(zip(w.city_station.tolist(), st_city.tolist()))
    for a, b, g in zip(ev["infected"][on_transport], ev["infector"][on_transport], ev["group"][on_transport]):
        s = station_of_group[int(g)]
        for q in (a, b):
            code = int(w.person_commute[q])
            assert code >= 0
            if code & 1:
                assert code >> 1 == s
            else:
                assert city_of_station[s] == code >> 1
    # a step without commute leaves everybody off the transports
    eng.step(eng.make_params(1.0, 8 / 24, 1, 0b10101, seed=3, flags=0, beta=beta))
    assert not ((eng.fetch_person_state()["pstate"] & 7) == ACT_COMMUTE).any()


@pytest.mark.gpu
def test_cuda_matches_oracle_with_commute_and_leisure():
    assert have_gpu()
    w = generate_world(80_000, seed=8, sector_betas=load_defaults()["company_sector_betas"], commute=True, leisure=True)
    (a, b), beta = engines(w, ["cuda", "oracle"])
    from test_leisure import MASK_LEISURE, MASK_WORK, tables_for
    does, cum = tables_for(w)
    seeds = np.sort(np.random.default_rng(2).choice(w.n_people, 4000, replace=False)).astype(np.uint32)
    for e in (a, b):
        e.set_leisure_tables(does[None], cum[None])
        e.infect(seeds, time=-4.0, seed=3, step=999)
    sched = [(MASK_COMMUTE, 1 / 24, 0), (MASK_WORK, 8 / 24, 1), (MASK_COMMUTE, 1 / 24, 0), (MASK_LEISURE, 3 / 24, 1), (0b10001, 11 / 24, 0)]
    now, on_transport = 0.0, 0
    flags = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION | _lib.STEP_RECORD_EVENTS
    for i in range(15):
        m, dt, tab = sched[i % 5]
        sa = a.step(a.make_params(now, dt, i, m, seed=21, flags=flags, beta=beta, leisure_table=tab))
        sb = b.step(b.make_params(now, dt, i, m, seed=21, flags=flags, beta=beta, leisure_table=tab))
        for key in ("n_new_infections", "n_infected", "n_draws", "n_by_activity", "n_hospitalised", "n_dead_total"):
            assert sa[key] == sb[key], (i, key, sa[key], sb[key])
        ea, eb = a.fetch_infection_events(), b.fetch_infection_events()
        for key in ("infected", "infector", "group"):
            assert np.array_equal(ea[key], eb[key]), (i, key)
        pa, pb = a.fetch_person_state(), b.fetch_person_state()
        assert np.array_equal(pa["pstate"], pb["pstate"]), f"step {i}: placement / flags differ"
        sg = w.supergroup("city_transports")
        on_transport += int(((ea["group"] >= sg.first_group) & (ea["group"] < w.supergroup("schools").first_group)).sum())
        now += dt
    assert on_transport > 0


@pytest.mark.gpu
def test_injected_commute_placement():
    assert have_gpu()
    w = generate_world(30_000, seed=6, commute=True)
    (a, b), beta = engines(w, ["cuda", "oracle"])
    rng = np.random.default_rng(3)
    ct = w.supergroup("city_transports")
    tr = np.full(w.n_people, -1, np.int32)
    goers = rng.choice(w.n_people, 5000, replace=False)
    tr[goers] = ct.first_group + rng.integers(0, min(20, ct.n_groups), 5000)
    seeds = np.sort(rng.choice(w.n_people, 1500, replace=False)).astype(np.uint32)
    for e in (a, b):
        e.infect(seeds, time=-4.0, seed=3, step=999)
        e.step(e.make_params(-0.5, 12 / 24, 7, 0b10001, seed=3, flags=0, beta=beta))
        e.inject_commute(tr)
    flags = _lib.STEP_INJECT_COMMUTE | _lib.STEP_RECORD_LAMBDA
    sa = a.step(a.make_params(0.0, 1 / 24, 0, MASK_COMMUTE, seed=1, flags=flags, beta=beta))
    sb = b.step(b.make_params(0.0, 1 / 24, 0, MASK_COMMUTE, seed=1, flags=flags, beta=beta))
    for key in ("n_new_infections", "n_infected", "n_draws", "n_by_activity"):
        assert sa[key] == sb[key], (key, sa[key], sb[key])
    assert np.array_equal((a.fetch_person_state()["pstate"] & 7) == ACT_COMMUTE, tr >= 0)
    la, lb = a.fetch_lambda(), b.fetch_lambda()
    assert np.array_equal(np.isnan(la), np.isnan(lb))
    ok = ~np.isnan(la)
    np.testing.assert_allclose(la[ok], lb[ok], rtol=1e-9, atol=0)
    assert ok[tr >= 0].sum() > 1000
