"""Variants: per-variant transmission distributions, cross immunity
(``Infection.immunity_ids``, infection.py:136-191) and the ``Mutation`` event (event/mutation.py:9-66)."""
import numpy as np
import pytest

from june_b200 import _lib
from june_b200.disease import DiseaseConfig, outcome_probabilities, synthetic_rates
from june_b200.interaction import Interaction, load_defaults
from june_b200.synth import generate_world

from helpers import default_beta_table, have_gpu, make_engine

FLAGS = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION


def variant_tables(dc):
    """covid19.yaml and b117.yaml differ in the scale of max_infectiousness (1.0 vs 1.7)."""
    code, base = dc.transmission_table()
    b117 = [(c, list(p)) for c, p in base]
    assert base[0][0] == 3  # lognormal(s, loc, scale)
    b117[0] = (3, [base[0][1][0], base[0][1][1], 1.7, 0.0])
    return [base, b117]


def setup(kinds, n=30_000, cross=(0b11, 0b11)):
    w = generate_world(n, seed=7, sector_betas=load_defaults()["company_sector_betas"])
    inter, dc = Interaction.from_file(), DiseaseConfig("covid19")
    prob = outcome_probabilities(dc, *synthetic_rates())
    engs = [make_engine(k, w, prob, interaction=inter, disease=dc, n_variants=2) for k in kinds]
    for e in engs:
        e.set_variants(variant_tables(dc), cross)
    return w, engs, default_beta_table(w, inter)


def run_steps(e, beta, n_steps, start=0, now=0.0):
    for i in range(start, start + n_steps):
        m, dt = (0b10101, 8 / 24) if i % 2 == 0 else (0b10001, 16 / 24)
        e.step(e.make_params(now, dt, i, m, seed=12, flags=FLAGS, beta=beta))
        now += dt
    return now


def mutation_case(kinds):
    w, engs, beta = setup(kinds)
    seeds = np.sort(np.random.default_rng(0).choice(w.n_people, 3000, replace=False)).astype(np.uint32)
    probs = np.array([0.0, 1.0, 0.5, 0.25] + [0.1] * (w.n_regions - 4))[: w.n_regions]
    out = []
    for e in engs:
        e.infect(seeds, time=-3.0, seed=5, step=700)
        now = run_steps(e, beta, 4)
        before = {r["person"]: r for r in e.fetch_infections()}
        st_before = e.fetch_person_state()
        e.mutate(1, probs, seed=99, step=4)
        after = {r["person"]: r for r in e.fetch_infections()}
        st_after = e.fetch_person_state()
        run_steps(e, beta, 4, start=4, now=now)
        out.append((before, after, st_before, st_after, e.fetch_person_state(), {r["person"]: r for r in e.fetch_infections()}))
    return w, probs, out


def check_mutation(w, probs, before, after, st_before, st_after):
    region = w.sa_region[w.super_area]
    assert before.keys() == after.keys()
    mutated = np.array([p for p in after if after[p]["variant"] == 1], np.int64)
    for r in range(w.n_regions):
        infected_r = [p for p in before if region[p] == r]
        k = int((region[mutated] == r).sum())
        if probs[r] == 0.0:
            assert k == 0
        elif probs[r] == 1.0:
            assert k == len(infected_r)
        elif len(infected_r) > 100:
            assert abs(k - probs[r] * len(infected_r)) < 4.5 * np.sqrt(len(infected_r) * probs[r] * (1 - probs[r]))
    n_scale = []
    for p in before:
        b, a = before[p], after[p]
        # symptoms and start time are untouched (mutation.py:62-64)
        for key in ("start_time", "stage", "tag", "max_tag", "traj_time", "traj_tag"):
            assert a[key] == b[key], key
        if a["variant"] == 1:
            assert a["tpar"] != b["tpar"] and st_after["trans_prob"][p] == 0.0  # fresh transmission object
            n_scale.append(a["tpar"][3] / 1.7)
        else:
            assert a["tpar"] == b["tpar"] and st_after["trans_prob"][p] == st_before["trans_prob"][p]
    # lognormal(s, loc=0, scale) has median `scale`: the new norms carry the variant's 1.7
    code, base = DiseaseConfig("covid19").transmission_table()
    assert abs(np.median(n_scale) - base[0][1][2]) < 0.05 * base[0][1][2]
    assert np.array_equal(st_before["pstate"] & 0x78, st_after["pstate"] & 0x78)


def test_oracle_mutation_converts_the_regional_fraction_and_keeps_symptoms():
    w, probs, out = mutation_case(["oracle"])
    before, after, sb, sa, final, final_rows = out[0]
    check_mutation(w, probs, before, after, sb, sa)
    # the new variant spreads afterwards
    assert sum(1 for p, r in final_rows.items() if r["variant"] == 1 and p not in after) > 0


def test_cross_immunity_masks():
    """No cross protection: somebody infected by variant 0 stays fully susceptible to variant 1."""
    w, (e,), beta = setup(["oracle"], n=5_000, cross=(0b01, 0b10))
    e.infect(np.arange(0, 5000, 10, dtype=np.uint32), time=-1.0, seed=5, step=700, variants=np.zeros(500, np.uint32))
    e.infect(np.arange(5, 5000, 10, dtype=np.uint32), time=-1.0, seed=5, step=701, variants=np.ones(500, np.uint32))
    su = e.fetch_person_state()["susceptibility"]
    assert (su[0, 0::10] == 0).all() and (su[1, 0::10] == 1).all()
    assert (su[1, 5::10] == 0).all() and (su[0, 5::10] == 1).all()


@pytest.mark.gpu
def test_cuda_mutation_and_variants_equal_the_oracle():
    assert have_gpu()
    w, probs, out = mutation_case(["cuda", "oracle"])
    (ba, aa, sba, saa, fa, ra), (bb, ab, sbb, sab, fb, rb) = out
    check_mutation(w, probs, ba, aa, sba, saa)
    assert aa.keys() == ab.keys() and ra.keys() == rb.keys()
    for rows_a, rows_b in ((aa, ab), (ra, rb)):
        for p in rows_a:
            x, y = rows_a[p], rows_b[p]
            assert x["variant"] == y["variant"] and x["stage"] == y["stage"] and x["traj_tag"] == y["traj_tag"]
            np.testing.assert_allclose(x["tpar"], y["tpar"], rtol=1e-12, atol=1e-12)
    assert np.array_equal(fa["pstate"], fb["pstate"]) and np.array_equal(fa["susceptibility"], fb["susceptibility"])


@pytest.mark.gpu
def test_cuda_cross_immunity_equals_the_oracle():
    assert have_gpu()
    w, (a, b), beta = setup(["cuda", "oracle"], n=20_000, cross=(0b01, 0b11))
    rng = np.random.default_rng(2)
    seeds = np.sort(rng.choice(w.n_people, 2000, replace=False)).astype(np.uint32)
    var = rng.integers(0, 2, 2000).astype(np.uint32)
    for e in (a, b):
        e.infect(seeds, time=-4.0, seed=5, step=700, variants=var)
        run_steps(e, beta, 8)
    sa, sb = a.fetch_person_state(), b.fetch_person_state()
    assert np.array_equal(sa["pstate"], sb["pstate"]) and np.array_equal(sa["susceptibility"], sb["susceptibility"])
    inf0 = seeds[var == 0]
    assert (sa["susceptibility"][1][inf0] == 1).all(), "variant 0 gives no protection against variant 1 here"
