"""Health events: the rows of the reference's ``symptoms``
(epidemiology.py:256-266), ``deaths`` (:176-186), ``recoveries`` (:209-214), ``hospital_admissions`` /
``icu_admissions`` / ``discharges`` (medical_care_policies.py:463-515) record tables."""
import numpy as np
import pytest

from june_b200 import _lib
from june_b200.interaction import load_defaults
from june_b200.synth import generate_world
from june_b200.world import ACT_RESIDENCE

from helpers import have_gpu
from test_leisure import MASK_WORK, engines

FLAGS = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION | _lib.STEP_RECORD_EVENTS
DEATH, RECOVERY, WARD, ICU, DISCHARGE, SYMPTOMS = 1, 2, 3, 4, 5, 6


def run(eng, w, beta, n_steps, seed=5):
    seeds = np.sort(np.random.default_rng(1).choice(w.n_people, w.n_people // 6, replace=False)).astype(np.uint32)
    eng.infect(seeds, time=-9.0, seed=11, step=1000)
    out, now = [], 0.0
    for i in range(n_steps):
        m, dt = (MASK_WORK & ~0b1000, 12 / 24) if i % 2 == 0 else (0b10001, 12 / 24)
        before = eng.fetch_person_state()
        st = eng.step(eng.make_params(now, dt, i, m, seed=seed, flags=FLAGS, beta=beta))
        out.append((before, eng.fetch_person_state(), st, eng.fetch_health_events()))
        now += dt
    return out


def test_oracle_health_events_follow_the_state_transitions():
    w = generate_world(40_000, seed=2, sector_betas=load_defaults()["company_sector_betas"])
    (eng,), beta = engines(w, ["oracle"])
    hosp = w.supergroup("hospitals")
    res_group = np.full(w.n_people, -1, np.int64)
    grp_of = np.repeat(np.arange(w.n_groups), np.diff(w.grp_offset.astype(np.int64)))
    is_res = ((w.reg_info >> 8) & 7) == ACT_RESIDENCE
    res_group[w.reg_member[is_res]] = grp_of[is_res]
    seen = {k: 0 for k in range(1, 7)}
    for before, after, st, ev in run(eng, w, beta, 40):
        p, k, g = ev["person"].astype(np.int64), ev["kind"], ev["group"]
        assert np.array_equal(np.lexsort((k, p)), np.arange(len(p))), "rows come sorted by (person, kind)"
        for kind in seen:
            seen[kind] += int((k == kind).sum())
        assert int((k == DEATH).sum()) == st["n_deaths_step"] and int((k == RECOVERY).sum()) == st["n_recovered_step"]
        inf_b, inf_a = (before["pstate"] & 0x08) != 0, (after["pstate"] & 0x08) != 0
        dead_a = (after["pstate"] & 0x10) != 0
        med_b, med_a = before["medical"], after["medical"]
        assert (inf_b[p] | inf_a[p]).all(), "only infected people (before, or newly this step) have health events"
        # symptoms: exactly the people whose tag changed while they stayed infected or ended the infection
        sym = k == SYMPTOMS
        still = inf_a[p[sym]]
        assert np.array_equal(g[sym][still], after["tag"][p[sym]][still].astype(np.int32))
        assert (before["tag"][p[sym]] != g[sym]).all()
        # deaths: in hospital -> the hospital, else the residence
        d = k == DEATH
        assert dead_a[p[d]].all() and not inf_a[p[d]].any()
        exp = np.where(med_b[p[d]] >= 0, med_b[p[d]] >> 8, res_group[p[d]])
        assert np.array_equal(g[d], exp.astype(np.int32))
        # recoveries carry no location
        r = k == RECOVERY
        assert (g[r] == -1).all() and not inf_a[p[r]].any() and not dead_a[p[r]].any()
        # ward admissions / ICU transfers / discharges against the medical-facility word before and after
        wa = k == WARD
        assert (med_b[p[wa]] < 0).all() and ((med_a[p[wa]] & 0xFF) == 1).all() and np.array_equal(med_a[p[wa]] >> 8, g[wa])
        ic = k == ICU
        assert ((med_b[p[ic]] & 0xFF) == 1).all() and ((med_a[p[ic]] & 0xFF) == 2).all() and np.array_equal(med_a[p[ic]] >> 8, g[ic])
        di = k == DISCHARGE
        assert (med_b[p[di]] >= 0).all() and (med_a[p[di]] < 0).all() and np.array_equal(med_b[p[di]] >> 8, g[di])
        for sel in (wa, ic, di):
            assert ((g[sel] >= hosp.first_group) & (g[sel] < hosp.first_group + hosp.n_groups)).all()
        # completeness: every ward admission / discharge visible in the state has its row
        newly_ward = (med_b < 0) & (med_a >= 0) & ((med_a & 0xFF) == 1)
        assert int(newly_ward.sum()) == int(wa.sum())
        released = (med_b >= 0) & (med_a < 0) & ~dead_a
        assert int(released.sum()) == int(di.sum())
    assert all(seen[kind] > 0 for kind in seen), seen


def test_no_events_without_the_record_flag():
    w = generate_world(8_000, seed=2, sector_betas=load_defaults()["company_sector_betas"])
    (eng,), beta = engines(w, ["oracle"])
    eng.infect(np.arange(0, 8000, 7, dtype=np.uint32), time=-9.0, seed=11, step=1000)
    eng.step(eng.make_params(0.0, 0.5, 0, 0b10001, seed=1, flags=FLAGS & ~_lib.STEP_RECORD_EVENTS, beta=beta))
    assert len(eng.fetch_health_events()["person"]) == 0


@pytest.mark.gpu
def test_cuda_health_events_equal_the_oracles():
    assert have_gpu()
    w = generate_world(60_000, seed=3, sector_betas=load_defaults()["company_sector_betas"])
    (a, b), beta = engines(w, ["cuda", "oracle"])
    total = {k: 0 for k in range(1, 7)}
    for (_, _, sa, ea), (_, _, sb, eb) in zip(run(a, w, beta, 40), run(b, w, beta, 40)):
        for key in ("person", "kind", "group"):
            assert np.array_equal(ea[key], eb[key]), key
        for k in total:
            total[k] += int((ea["kind"] == k).sum())
    assert all(v > 0 for v in total.values()), total
