"""CUDA path vs the CPU oracle on the same seeded inputs, through the C ABI.

* Philox mode, full do_timestep incl. device-side infection creation: infection events, person
  state and symptom tags must be identical step by step; transmission probabilities and
  infection exponents within 1e-9 relative (BASELINE.json tolerance; the two sides use different
  libm implementations).
* Injected-uniform mode on ragged hand-made worlds: empty groups, single-member groups,
  households larger than a 32-slot tile, companies larger than a tile, tiny schools -- draw
  ordinals must agree across every kernel flavour.
* Two variants (the multi-variant kernel instantiation).
"""
import numpy as np
import pytest

from june_b200 import _lib
from june_b200.disease import DiseaseConfig, outcome_probabilities, synthetic_rates
from june_b200.interaction import Interaction, load_defaults
from june_b200.synth import generate_world
from june_b200.world import (ACT_MEDICAL, ACT_PRIMARY, ACT_RESIDENCE, KIND_MATRIX, KIND_SCHOOL, BetaRule, Supergroup,
                             WorldSoA, make_ginfo, make_info)

from helpers import default_beta_table, have_gpu, make_engine

pytestmark = pytest.mark.gpu
REL = 1e-9
MASKS = [0b10001, 0b10101, 0b10001, 0b10001, 0b10101]  # (medical|residence), +primary
DTS = [1 / 24, 8 / 24, 1 / 24, 3 / 24, 11 / 24]


def setup(world, n_variants=1):
    inter = Interaction.from_file()
    dc = DiseaseConfig("covid19")
    rates, bins = synthetic_rates()
    prob = outcome_probabilities(dc, rates, bins)
    a = make_engine("cuda", world, prob, interaction=inter, disease=dc, n_variants=n_variants)
    b = make_engine("oracle", world, prob, interaction=inter, disease=dc, n_variants=n_variants)
    return a, b, default_beta_table(world, inter)


def compare_state(a, b, tag=""):
    sa, sb = a.fetch_person_state(), b.fetch_person_state()
    assert np.array_equal(sa["pstate"], sb["pstate"]), f"{tag}: pstate"
    assert np.array_equal(sa["tag"], sb["tag"]), f"{tag}: tags"
    assert np.array_equal(sa["medical"], sb["medical"]), f"{tag}: wards"
    assert np.array_equal(sa["susceptibility"], sb["susceptibility"]), f"{tag}: susceptibility"
    np.testing.assert_allclose(sa["trans_prob"], sb["trans_prob"], rtol=REL, atol=0, err_msg=f"{tag}: trans")


def test_graph_replay_matches_plain_launches_and_oracle():
    """Without the record/inject test flags the step runs as a replayed CUDA graph: state must
    still equal the oracle's, step after step."""
    assert have_gpu()
    w = generate_world(40_000, seed=8, sector_betas=load_defaults()["company_sector_betas"])
    a, b, beta = setup(w)
    seeds = np.sort(np.random.default_rng(1).choice(w.n_people, 1200, replace=False)).astype(np.uint32)
    for e in (a, b):
        e.infect(seeds, time=-5.0, seed=11, step=1000)
    flags = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION
    now = 0.0
    for i in range(30):
        m, dt = MASKS[i % 5], DTS[i % 5]
        sa = a.step(a.make_params(now, dt, i, m, seed=5, flags=flags, beta=beta))
        sb = b.step(b.make_params(now, dt, i, m, seed=5, flags=flags, beta=beta))
        for key in ("n_new_infections", "n_infected", "n_dead_total", "n_recovered_step", "n_hospitalised", "n_draws"):
            assert sa[key] == sb[key], (i, key, sa[key], sb[key])
        assert np.array_equal(a.fetch_new_infections()[0], b.fetch_new_infections()[0]), f"step {i}"
        if i % 5 == 4:
            compare_state(a, b, f"step {i}")
        now += dt
    compare_state(a, b, "end")


def test_philox_mode_full_steps_match_oracle():
    assert have_gpu()
    w = generate_world(60_000, seed=3, sector_betas=load_defaults()["company_sector_betas"])
    a, b, beta = setup(w)
    seeds = np.random.default_rng(0).choice(w.n_people, 1500, replace=False).astype(np.uint32)
    for k, part in enumerate(np.array_split(np.sort(seeds), 5)):
        for e in (a, b):
            e.infect(part, time=-2.0 * k, seed=11, step=1000 + k)
    compare_state(a, b, "after seeding")
    flags = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION | _lib.STEP_RECORD_LAMBDA
    now, total_new = 0.0, 0
    for i in range(40):
        m, dt = MASKS[i % 5], DTS[i % 5]
        pa = a.make_params(now, dt, i, m, seed=77, flags=flags, beta=beta)
        pb = b.make_params(now, dt, i, m, seed=77, flags=flags, beta=beta)
        sa, sb = a.step(pa), b.step(pb)
        ma, va = a.fetch_new_infections()
        mb, vb = b.fetch_new_infections()
        assert np.array_equal(ma, mb), f"step {i}: infected sets differ"
        la, lb = a.fetch_lambda(), b.fetch_lambda()
        assert np.array_equal(np.isnan(la), np.isnan(lb)), f"step {i}: people who drew differ"
        ok = ~np.isnan(la)
        np.testing.assert_allclose(la[ok], lb[ok], rtol=REL, atol=0, err_msg=f"step {i}: lambda")
        for key in ("n_new_infections", "n_infected", "n_dead_total", "n_recovered_step", "n_hospitalised", "n_icu",
                    "n_draws", "n_by_activity"):
            assert sa[key] == sb[key], (i, key, sa[key], sb[key])
        compare_state(a, b, f"step {i}")
        # per-region summary (records_writer.py:151-164): identical, and consistent with the totals
        ra, rb = a.fetch_summary(), b.fetch_summary()
        assert np.array_equal(ra, rb), f"step {i}: per-region summary"
        # (ward + intensive-care patients of the summary are the people PLACED in their ward by this step, as in the reference)
        assert ra[:, 1].sum() == sa["n_infected"] and ra[:, 2].sum() + ra[:, 3].sum() == sa["n_by_activity"][ACT_MEDICAL]
        assert ra[:, 4].sum() == sa["n_deaths_step"] and ra[:, 6].sum() == sa["n_recovered_step"]
        assert ra[:, 0].sum() == sa["n_new_infections"] - sa["n_halo_infected"]
        total_new += len(ma)
        now += dt
    assert total_new > 300
    ia, ib = a.fetch_infections(), b.fetch_infections()
    assert [r["person"] for r in ia] == [r["person"] for r in ib]
    for ra, rb in zip(ia, ib):
        assert ra["traj_tag"] == rb["traj_tag"] and ra["stage"] == rb["stage"] and ra["max_tag"] == rb["max_tag"]
        np.testing.assert_allclose(ra["traj_time"], rb["traj_time"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(ra["tpar"], rb["tpar"], rtol=1e-12, atol=1e-12)


def ragged_world(rng, big_household=True):
    """Hand-made world exercising the awkward shapes."""
    hh_sizes = [1, 2, 0, 5, 40 if big_household else 6, 1, 1, 3, 0, 0, 7, 33, 2, 32, 31, 1] + list(rng.integers(0, 9, 60))
    comp_sizes = [0, 1, 70, 2, 33, 32, 31, 5, 0, 12, 200, 3]
    n = int(sum(hh_sizes))
    age = rng.integers(0, 100, n).astype(np.uint8)
    sex = rng.integers(0, 2, n).astype(np.uint8)
    rows = []
    p = 0
    n_hosp, n_school = 2, 2
    g_school0, g_comp0 = n_hosp, n_hosp + n_school
    g_hh0 = g_comp0 + len(comp_sizes)
    G = g_hh0 + len(hh_sizes)
    has = np.zeros(n, np.uint8)
    for h, sz in enumerate(hh_sizes):
        for _ in range(sz):
            a_ = age[p]
            lsg = 0 if a_ < 18 else 1 if a_ <= 25 else 2 if a_ < 65 else 3
            rows.append((g_hh0 + h, lsg, p, ACT_RESIDENCE))
            has[p] |= 1 << ACT_RESIDENCE
            p += 1
    people = rng.permutation(n)
    pos = 0
    for c, sz in enumerate(comp_sizes):
        for q in people[pos:pos + sz]:
            rows.append((g_comp0 + c, 0, int(q), ACT_PRIMARY)); has[q] |= 1 << ACT_PRIMARY
        pos += sz
    years = []
    aux = np.zeros(G, np.uint32)
    for s_, ny in enumerate((3, 1)):
        aux[g_school0 + s_] = (len(years) << 8) | ny
        years += list(range(5, 5 + ny))
        for q in people[pos:pos + 3]:
            rows.append((g_school0 + s_, 0, int(q), ACT_PRIMARY)); has[q] |= 1 << ACT_PRIMARY
        pos += 3
        for y in range(ny):
            k = 45 if (s_ == 0 and y == 1) else 4
            for q in people[pos:pos + k]:
                rows.append((g_school0 + s_, 1 + y, int(q), ACT_PRIMARY)); has[q] |= 1 << ACT_PRIMARY
            pos += k
    for q in people[pos:pos + 9]:
        rows.append((int(q) % n_hosp, 0, int(q), ACT_PRIMARY)); has[q] |= 1 << ACT_PRIMARY
    rows.sort()
    g_arr = np.array([r[0] for r in rows])
    off = np.zeros(G + 1, np.uint32)
    off[1:] = np.cumsum(np.bincount(g_arr, minlength=G))
    beta_idx = np.zeros(G, np.uint32)
    beta_idx[g_school0:g_comp0] = 1
    beta_idx[g_comp0:g_hh0] = 2
    beta_idx[g_hh0:] = 3
    scale = np.zeros(G, np.uint32)
    scale[g_comp0:g_hh0] = rng.integers(0, 3, len(comp_sizes))
    region = rng.integers(0, 2, G).astype(np.uint32)
    w = WorldSoA(
        age=age, sex=sex, care_home_resident=np.zeros(n, np.uint8), has_activity=has,
        super_area=rng.integers(0, 3, n).astype(np.uint32), sa_hospital=np.array([0, 1, 1], np.int32),
        sa_region=np.array([0, 1, 0], np.uint16), region_names=["North East", "London"],
        grp_offset=off, grp_info=make_ginfo(beta_idx, scale, region).astype(np.uint32), grp_aux=aux,
        reg_member=np.array([r[2] for r in rows], np.uint32),
        reg_info=make_info(np.array([r[1] for r in rows]), np.array([r[3] for r in rows])).astype(np.uint16),
        school_years=np.array(years, np.uint8),
        supergroups=[Supergroup("hospitals", "hospital", ACT_MEDICAL, 0, n_hosp, KIND_MATRIX, 1, 1),
                     Supergroup("schools", "school", ACT_PRIMARY, g_school0, n_school, KIND_SCHOOL, 1, 0),
                     Supergroup("companies", "company", ACT_PRIMARY, g_comp0, len(comp_sizes), KIND_MATRIX, 0, 0),
                     Supergroup("households", "household", ACT_RESIDENCE, g_hh0, len(hh_sizes), KIND_MATRIX, 0, 0)],
        beta_rules=[BetaRule("hospital"), BetaRule("school"), BetaRule("company"), BetaRule("household", None, False),
                    BetaRule("household_visits", None, False), BetaRule("care_visits", None, False)],
        beta_scale=np.array([1.0, 0.5, 0.3]))
    w.validate()
    return w


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("modes", [(0, 0), (1, 1), (0, 1), (1, 0)])
def test_injected_uniforms_on_ragged_worlds(seed, modes):
    assert have_gpu()
    rng = np.random.default_rng(seed)
    w = ragged_world(rng)
    w.supergroup("companies").launch_mode, w.supergroup("households").launch_mode = modes
    a, b, beta = setup(w)
    n = w.n_people
    inf = rng.choice(n, n // 5, replace=False).astype(np.uint32)
    for e in (a, b):
        e.infect(np.sort(inf), time=-4.0, seed=5, step=900)
    flags = (_lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION | _lib.STEP_RECORD_LAMBDA | _lib.STEP_INJECT_UNIFORMS)
    now = 0.0
    drawn = 0
    for i in range(12):
        u = rng.random(4 * n)
        for e in (a, b):
            e.inject_uniforms(u)
        m, dt = MASKS[i % 5], DTS[i % 5] * 3
        sa = a.step(a.make_params(now, dt, i, m, seed=1, flags=flags, beta=beta))
        sb = b.step(b.make_params(now, dt, i, m, seed=1, flags=flags, beta=beta))
        assert sa["n_draws"] == sb["n_draws"], f"step {i}"
        la, lb = a.fetch_lambda(), b.fetch_lambda()
        assert np.array_equal(np.isnan(la), np.isnan(lb)), f"step {i}"
        ok = ~np.isnan(la)
        np.testing.assert_allclose(la[ok], lb[ok], rtol=REL, atol=0)
        assert np.array_equal(a.fetch_new_infections()[0], b.fetch_new_infections()[0]), f"step {i}: infected sets"
        compare_state(a, b, f"step {i}")
        drawn += sa["n_draws"]
        now += dt
    assert drawn > 200


def test_two_variants_match_oracle():
    assert have_gpu()
    w = generate_world(20_000, seed=9)
    a, b, beta = setup(w, n_variants=2)
    rng = np.random.default_rng(4)
    seeds = np.sort(rng.choice(w.n_people, 800, replace=False)).astype(np.uint32)
    variants = rng.integers(0, 2, len(seeds)).astype(np.uint32)
    susc = np.ones((2, w.n_people))
    susc[1, rng.random(w.n_people) < 0.3] = 0.4  # partial cross-immunity against variant 1
    for e in (a, b):
        e.set_person_state(susceptibility=susc)
        e.infect(seeds, time=-5.0, seed=3, step=800, variants=variants)
    flags = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION | _lib.STEP_RECORD_LAMBDA
    now, seen = 0.0, np.zeros(2, int)
    for i in range(15):
        m, dt = MASKS[i % 5], DTS[i % 5]
        a.step(a.make_params(now, dt, i, m, seed=21, flags=flags, beta=beta))
        b.step(b.make_params(now, dt, i, m, seed=21, flags=flags, beta=beta))
        (ma, va), (mb, vb) = a.fetch_new_infections(), b.fetch_new_infections()
        assert np.array_equal(ma, mb) and np.array_equal(va, vb), f"step {i}"
        seen += np.bincount(va, minlength=2)
        compare_state(a, b, f"step {i}")
        now += dt
    assert seen.min() > 5


@pytest.mark.gpu
def test_a_rejected_step_leaves_the_world_usable():
    """Error behaviour at the boundary: a step the library refuses (here: injected leisure placement that was
    never supplied) raises, changes nothing, and the following steps run exactly as if it had not happened."""
    assert have_gpu()
    w = generate_world(20_000, seed=3, sector_betas=load_defaults()["company_sector_betas"])
    a, b, beta = setup(w)
    b.close()
    inter, dc = Interaction.from_file(), DiseaseConfig("covid19")
    b = make_engine("cuda", w, outcome_probabilities(dc, *synthetic_rates()), interaction=inter, disease=dc)
    seeds = np.arange(0, w.n_people, 9, dtype=np.uint32)
    flags = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION
    for e in (a, b):
        e.infect(seeds, time=-4.0, seed=5, step=900)
    now = 0.0
    for i in range(6):
        m, dt = MASKS[i % 5], DTS[i % 5]
        if i in (0, 3):  # first use of this step shape, and a shape whose graph already exists
            with pytest.raises(Exception) as err:
                a.step(a.make_params(now, dt, i, m, seed=1, flags=flags | _lib.STEP_INJECT_LEISURE, beta=beta))
            assert "jb_inject_leisure" in str(err.value)
        sa = a.step(a.make_params(now, dt, i, m, seed=1, flags=flags, beta=beta))
        sb = b.step(b.make_params(now, dt, i, m, seed=1, flags=flags, beta=beta))
        assert sa == sb, i
        compare_state(a, b, f"step {i}")
        now += dt
    assert sa["n_infected"] > 1000
