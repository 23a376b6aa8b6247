"""Parity against golden traces of the UNMODIFIED reference (tests/golden/make_golden.py).

The same checks run against
  * the CPU oracle (``oracle/june_oracle.c``)            -- here, without a GPU;
  * the CUDA path through the C ABI (``libjune_b200.so``) -- ``-m gpu`` on the B200 box.

Per reference ``do_timestep`` we check, with the reference's own uniform stream injected:
  placement (activity chosen per person; in the leisure trace the reference's own venue choices
  are replayed and the visitors' draws checked), the infection exponent Lambda of every Bernoulli draw
  (<= 1e-9 relative, BASELINE.json), the set of newly infected people (bit-exact), and the person
  state after the health update (infected/dead flags, symptom tag and hospital ward exact;
  transmission probability <= 1e-9 relative).
"""
import numpy as np
import pytest

from june_b200 import _lib
from june_b200.interaction import Interaction
from june_b200.world import ACT_COMMUTE, ACT_LEISURE, PS_ACT_MASK, PS_DEAD, PS_INFECTED

from helpers import Golden, default_beta_table, have_gpu, make_engine, rows_to_dicts

REL = 1e-9  # tolerance stated in BASELINE.json:north_star
FLAGS = (_lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION | _lib.STEP_INJECT_UNIFORMS
         | _lib.STEP_NO_NEW_INFECTIONS | _lib.STEP_RECORD_LAMBDA)


# parameters of the lockdown policies of the w2k_policies trace (tests/golden/ref_world.py:make_simulator)
def policy_records(names, ratios):
    f, k, r = (None if np.isnan(x) else float(x) for x in ratios)
    table = {
        "Quarantine": dict(type=_lib.POL_QUARANTINE, mask=(1 << (2 + 2)) | (1 << (3 + 2)), p=[7, 0.8]),  # mild, severe
        "Shielding": dict(type=_lib.POL_SHIELDING, p=[70, 0.6]),
        "CloseSchools": dict(type=_lib.POL_CLOSE_SCHOOLS, mask=sum(1 << y for y in (5, 6, 7, 12, 13, 17)), p=[0.7]),
        "CloseCompanies": dict(type=_lib.POL_CLOSE_COMPANIES, p=[0.3, 0.2, 0.2, f, k, r]),
        "LimitLongCommute": dict(type=_lib.POL_LIMIT_LONG_COMMUTE, p=[0.4]),
    }
    return [(n, table[n]) for n in names if n in table]


# the campaigns of the w2k_vaccination trace (tests/golden/ref_world.py: VACCINES, make_simulator)
def vaccination_tables():
    import datetime

    from june_b200.vaccination import Vaccine, VaccinationCampaign, VaccinationCampaigns
    cfg = {
        "Pfizer": dict(days_administered_to_effective=[2, 2, 5], days_effective_to_waning=[1, 2, 1], days_waning=[2, 2, 1],
                       waning_factor=0.8,
                       sterilisation_efficacies=[{"Covid19": {"0-100": 0.52}}, {"Covid19": {"0-70": 0.95, "70-100": 0.85}}, {"Covid19": {"0-100": 0.98}}],
                       symptomatic_efficacies=[{"Covid19": {"0-100": 0.6}}, {"Covid19": {"0-100": 0.9}}, {"Covid19": {"0-100": 0.98}}]),
        "AstraZeneca": dict(days_administered_to_effective=[3, 7, 5], days_effective_to_waning=[2, 1, 1], days_waning=[3, 1, 1],
                            waning_factor=1.0,
                            sterilisation_efficacies=[{"Covid19": {"0-100": 0.32}}, {"Covid19": {"0-100": 0.75}}, {"Covid19": {"0-100": 0.88}}],
                            symptomatic_efficacies=[{"Covid19": {"0-100": 0.4}}, {"Covid19": {"0-100": 0.75}}, {"Covid19": {"0-100": 0.88}}]),
    }
    pf, az = (Vaccine.from_config_dict(k, cfg[k]) for k in ("Pfizer", "AstraZeneca"))
    camps = VaccinationCampaigns([
        VaccinationCampaign(vaccine=pf, days_to_next_dose=[0, 3], dose_numbers=[0, 1], start_time="2020-03-03",
                            end_time="2020-03-08", group_by="age", group_type="60-100", group_coverage=0.6),
        VaccinationCampaign(vaccine=az, days_to_next_dose=[0], dose_numbers=[0], start_time="2020-03-04",
                            end_time="2020-03-07", group_by="age", group_type="30-45", group_coverage=0.5)])
    return camps.device_tables(datetime.date(2020, 3, 2), ["Covid19"])


def run_trace(kind: str, name: str, max_steps=None, launch_modes=None):
    g = Golden(name)
    if launch_modes is not None:  # exercise the other kernel flavour on the same trace
        for sg in g.world.supergroups:
            if sg.spec in launch_modes and not sg.has_dynamic:  # (care homes that take visitors stay on the per-group kernels)
                sg.launch_mode = launch_modes[sg.spec]
    inter = Interaction.from_file()
    # JB_MIRROR=0: companies / schools gather their members' bytes through the person index instead of
    # streaming the slot-ordered state mirror (read by the CUDA library when the world is created)
    import os
    os.environ["JB_MIRROR"] = str((launch_modes or {}).get("_mirror", 1))
    try:
        eng = make_engine(kind, g.world, g.outcome_prob, interaction=inter)
    finally:
        os.environ.pop("JB_MIRROR", None)
    beta = default_beta_table(g.world, inter)
    n = g.world.n_people
    grp_sg = g.world.group_supergroup()
    sg_act = np.array([sg.activity for sg in g.world.supergroups])
    total_draws = total_inf = 0
    worst = 0.0
    if "s0_vac_today" in g.z.files:
        eng.set_vaccination(*vaccination_tables())
    for i in range(g.n_steps if max_steps is None else min(max_steps, g.n_steps)):
        s = g.step(i)
        if len(s["pre_rows"]):
            eng.upload_infections(rows_to_dicts(s["pre_rows"]))
        draw_u = s["draw_u"]
        if "lz_choice" in s and len(draw_u):
            # Residence visits.  The reference appends people to a subgroup as they are placed and re-appends the residents
            # of a household every time a visitor arrives (household.py:124-149), so inside a visited household (or care
            # home) its draws come in "last appended" order; the engines keep (subgroup, members before visitors, person).
            # Same groups, same people, same uniforms per person: the uniforms are handed over in the engines' order.
            dp_, pg_, pl_ = s["draw_person"], s["place_group"], s["place_lsg"]
            grp = pg_[dp_]
            block = np.cumsum(np.r_[0, grp[1:] != grp[:-1]])
            visitor = (s["lz_choice"][dp_] == grp).astype(np.int64)
            order = np.lexsort((dp_, visitor, pl_[dp_], block))
            assert np.array_equal(block[order], block), "a group's draws are contiguous in both orders"
            draw_u = draw_u[order]
            run_trace.visit_reorders = getattr(run_trace, "visit_reorders", 0) + int((order != np.arange(len(order))).sum())
        eng.inject_uniforms(draw_u)
        flags = FLAGS
        if g.world.has_leisure and "lz_choice" in s:
            # the destination every person drew at its turn (a venue, a household, a care home), -1 = none: who stays or
            # comes back home because its household receives a visitor is decided by the engine, as in the reference
            eng.inject_leisure(s["lz_choice"].astype(np.int32))
            flags |= _lib.STEP_INJECT_LEISURE
            run_trace.visits = getattr(run_trace, "visits", 0) + int(((s["lz_choice"] >= 0) & (sg_act[grp_sg[np.maximum(s["lz_choice"], 0)]] == 4)).sum())
            run_trace.fetched_home = getattr(run_trace, "fetched_home", 0) + int(((s["lz_choice"] >= 0) & (s["place_group"] != s["lz_choice"])).sum())
        elif g.world.has_leisure:
            # the reference's own leisure choices (leisure.py:253-263 draws from Python's / numba's
            # global generators) are replayed: venue per person, -1 = did not go out
            pg = s["place_group"]
            at_venue = (pg >= 0) & (sg_act[grp_sg[np.maximum(pg, 0)]] == ACT_LEISURE)
            eng.inject_leisure(np.where(at_venue, pg, -1).astype(np.int32))
            flags |= _lib.STEP_INJECT_LEISURE
        if g.world.has_commute:
            # the reference's own station / transport choices (randint in city.py:95-97, station.py:52-76)
            pg = s["place_group"]
            on_transport = (pg >= 0) & (sg_act[grp_sg[np.maximum(pg, 0)]] == ACT_COMMUTE)
            eng.inject_commute(np.where(on_transport, pg, -1).astype(np.int32))
            flags |= _lib.STEP_INJECT_COMMUTE
            run_trace.commuters = getattr(run_trace, "commuters", 0) + int(on_transport.sum())
            if int(s["activity_mask"]) & (1 << ACT_COMMUTE):
                # the compiled commute tables describe exactly what the reference did: every traveller is a
                # commuter of the tables and sits on a transport of one of "its" stations; commuters who
                # do not travel are in hospital / at home with severe symptoms / dead
                w_ = g.world
                assert (w_.person_commute[on_transport] >= 0).all()
                tr_station = np.repeat(np.arange(len(w_.station_transport_off) - 1), np.diff(w_.station_transport_off.astype(np.int64)))
                station_of = dict(zip(w_.station_transport.tolist(), tr_station.tolist()))
                for q in np.nonzero(on_transport)[0]:
                    code, st_ = int(w_.person_commute[q]), station_of[int(pg[q])]
                    if code & 1:
                        assert code >> 1 == st_
                    else:
                        c_ = code >> 1
                        assert st_ in w_.city_station[w_.city_station_off[c_]: w_.city_station_off[c_ + 1]]
                stayed = (w_.person_commute >= 0) & ~on_transport
                assert stayed.sum() <= 0.1 * max(1, on_transport.sum())
        if "pol_active" in s:
            # individual policies: the reference's active list (in order) and every random() it drew
            recs = policy_records([str(x) for x in s["pol_active"]], g.z["close_companies_ratios"])
            eng.set_individual_policies([r for _, r in recs], np.ones(g.world.n_regions))
            u = np.full((_lib.MAX_POLICIES, _lib.POLICY_DRAWS, n), np.nan)
            index = {nm: k for k, (nm, _) in enumerate(recs)}
            names = [str(x) for x in s["pol_names"]]
            for row in s["pol_draws"]:
                u[index[names[int(row[0])]], int(row[1]), int(row[2])] = row[3]
            eng.inject_policy_uniforms(u)
            flags |= _lib.STEP_INJECT_POLICY
            n_policy_draws = getattr(run_trace, "policy_draws", 0) + len(s["pol_draws"])
            run_trace.policy_draws = n_policy_draws
        if "health_events" in s:
            flags |= _lib.STEP_RECORD_EVENTS
        p = eng.make_params(now=float(s["now"]), delta_time=float(s["dt"]), step=i,
                            activity_mask=int(s["activity_mask"]), seed=0, flags=flags, beta=beta)
        eng.step_begin(p)
        eng.step_interact()
        # -- infection exponents, in draw order ------------------------------------------------
        lam = eng.fetch_lambda()[:n]
        dp = s["draw_person"]
        assert len(dp) == len(s["draw_u"]) == len(s["draw_lambda"])
        drawn = np.zeros(n, bool)
        drawn[dp] = True
        assert np.array_equal(~np.isnan(lam), drawn), f"step {i}: set of people who consumed a draw differs"
        if len(dp):
            ref = s["draw_lambda"]
            got = lam[dp]
            err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
            err[ref == 0] = np.abs(got[ref == 0])
            worst = max(worst, float(err.max()))
            assert err.max() <= REL, f"step {i}: Lambda differs by {err.max():.3e}"
        # -- infection events: bit-exact -----------------------------------------------------------
        member, variant = eng.fetch_new_infections()
        ref_new = np.sort(s["new_rows"][:, 0].astype(np.int64)) if len(s["new_rows"]) else np.zeros(0, np.int64)
        assert np.array_equal(member.astype(np.int64), ref_new), f"step {i}: infected set differs"
        if "infection_events" in s:
            # rows of the reference's `infections` table (interaction.py:664-683): who was infected where.  The
            # infector is np.random.choice in the reference (interaction.py:643-662), so only its validity is checked
            ev = eng.fetch_infection_events()
            ref_ev = s["infection_events"]
            assert np.array_equal(ev["infected"].astype(np.int64), ref_ev[:, 0]), f"step {i}: infections table, infected ids"
            assert np.array_equal(ev["group"].astype(np.int64), ref_ev[:, 2]), f"step {i}: infections table, locations"
            before = eng.fetch_person_state()
            assert ((before["pstate"][ev["infector"]] & PS_INFECTED) != 0).all() and (before["trans_prob"][ev["infector"]] > 0).all()
            assert np.array_equal(s["place_group"][ev["infector"]], s["place_group"][ref_ev[:, 1]]), f"step {i}: infector not in the reference's group"
            run_trace.infection_rows = getattr(run_trace, "infection_rows", 0) + len(ref_ev)
        if len(s["new_rows"]):
            eng.upload_infections(rows_to_dicts(s["new_rows"]))
        stats = eng.step_finish()
        assert stats["n_draws"] == len(dp)
        if "health_events" in s:
            # rows of the symptoms / deaths / recoveries / hospital_admissions / icu_admissions / discharges tables
            hev = eng.fetch_health_events()
            got = np.stack([hev["person"].astype(np.int64), hev["kind"].astype(np.int64), hev["group"].astype(np.int64)], axis=1)
            assert np.array_equal(got, s["health_events"]), f"step {i}: health event tables differ"
            for k_ in got[:, 1]:
                run_trace.health_rows[int(k_)] = run_trace.health_rows.get(int(k_), 0) + 1
        if "summary" in s:
            # the reference's own summarise_infections / _hospitalisations / _deaths on this step's events: the eight
            # numbers of a summary.csv row per region (records_writer.py:764-773)
            ref_sum = s["summary"]
            got_sum = eng.fetch_summary().astype(np.int64)
            # current_infected, current_hospitalised, daily_hospitalised, current_intensive_care, daily_intensive_care,
            # daily_hospital_deaths, daily_deaths
            assert np.array_equal(got_sum[:, [1, 2, 7, 3, 8, 5, 4]], ref_sum[:, [0, 2, 3, 4, 5, 6, 7]]), \
                f"step {i}: summary row differs\n{got_sum}\n{ref_sum}"
            if i > 0:  # daily_infected: rows of the infections table by the region of the group (step 0 holds the seeds)
                ev = eng.fetch_infection_events()
                by_group = np.bincount((g.world.grp_info[ev["group"]] >> 16).astype(np.int64), minlength=g.world.n_regions)
                assert np.array_equal(by_group, ref_sum[:, 1]), f"step {i}: daily_infected"
            run_trace.summary_rows = getattr(run_trace, "summary_rows", 0) + int((ref_sum.sum(axis=1) > 0).sum())
        if "vac_today" in s:
            # the day's vaccinations follow the health update (epidemiology.py:295-303); the reference's
            # own random() per person is replayed
            eng.inject_vaccination_uniforms(np.where(np.isnan(s["vac_u"]), 2.0, s["vac_u"]))
            eng.vaccinate(int(s["vac_today"]))
            run_trace.vaccine_draws = getattr(run_trace, "vaccine_draws", 0) + int((~np.isnan(s["vac_u"])).sum())
        # -- placement -------------------------------------------------------------------------------
        st = eng.fetch_person_state()
        act = st["pstate"] & PS_ACT_MASK
        pg = s["place_group"]
        exp_act = np.where(pg >= 0, sg_act[grp_sg[np.maximum(pg, 0)]], 7)
        # hospital workers sit in a hospital group through their primary activity
        mism = (act != exp_act) & (pg >= 0)
        for q in np.nonzero(mism)[0]:
            if "lz_choice" in s and s["lz_choice"][q] == pg[q] and act[q] == ACT_LEISURE:
                continue  # a visitor: in a household or care home through leisure
            lo, hi = g.world.grp_offset[pg[q]], g.world.grp_offset[pg[q] + 1]
            slot = lo + np.nonzero(g.world.reg_member[lo:hi] == q)[0]
            assert len(slot) and (g.world.reg_info[slot[0]] >> 8) == act[q], f"step {i}: person {q} misplaced"
        assert np.array_equal(act == 7, pg < 0), f"step {i}: set of placed people differs"
        # -- health update ---------------------------------------------------------------------------
        assert np.array_equal((st["pstate"] & PS_INFECTED) != 0, s["post_infected"] != 0), f"step {i}: infected flags"
        assert np.array_equal((st["pstate"] & PS_DEAD) != 0, s["post_dead"] != 0), f"step {i}: dead flags"
        assert np.array_equal(st["tag"], s["post_tag"]), f"step {i}: symptom tags"
        assert np.array_equal(st["medical"], s["post_med"]), f"step {i}: hospital wards"
        inf = s["post_infected"] != 0
        ref_t, got_t = s["post_trans"][inf], st["trans_prob"][inf]
        terr = np.abs(got_t - ref_t) / np.maximum(np.abs(ref_t), 1e-300)
        terr[ref_t == 0] = np.abs(got_t[ref_t == 0])
        if inf.any():
            worst = max(worst, float(terr.max()))
            assert terr.max() <= REL, f"step {i}: transmission probability differs by {terr.max():.3e}"
        assert np.array_equal(st["susceptibility"][0], s["post_susc"]), f"step {i}: susceptibilities"
        if "post_vac_dose" in s:
            vs = eng.fetch_vaccination()
            assert np.array_equal(vs["dose"], s["post_vac_dose"]), f"step {i}: person.vaccinated"
            assert np.array_equal(vs["effective_multiplier"][0], s["post_effmult"]), f"step {i}: effective multipliers"
        assert stats["n_infected"] == int(inf.sum())
        total_draws += len(dp)
        total_inf += len(ref_new)
    eng.close()
    return total_draws, total_inf, worst


FIXTURES = ["w3k", "w1k_long", "w2k_leisure", "w2k_policies", "w2k_commute", "w2k_vaccination", "w2k_records", "w2k_visits"]
run_trace.health_rows = {}


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_matches_reference_trace(name):
    run_trace.policy_draws = run_trace.commuters = run_trace.vaccine_draws = run_trace.infection_rows = run_trace.summary_rows = 0
    run_trace.visits = run_trace.fetched_home = run_trace.visit_reorders = 0
    run_trace.health_rows = {}
    draws, infections, worst = run_trace("oracle", name)
    if name == "w2k_records":
        assert run_trace.infection_rows > 500
        assert run_trace.summary_rows > 80, "summary.csv rows of the reference's own summarise_time_step must have been compared"
        assert all(run_trace.health_rows.get(k, 0) > 0 for k in range(1, 7)), run_trace.health_rows
    if name == "w2k_visits":
        assert run_trace.visits > 1000, "people must have visited households and care homes"
        assert run_trace.fetched_home > 50, "residents out doing leisure must have been fetched home by a visitor"
    if name == "w2k_policies":
        assert run_trace.policy_draws > 5000, "the lockdown policies must have drawn"
    if name == "w2k_commute":
        assert run_trace.commuters > 2000, "commuters must have travelled"
    if name == "w2k_vaccination":
        assert run_trace.vaccine_draws > 1500, "eligible people must have drawn for their jab"
    assert draws > 10000 and infections > 100
    assert worst <= REL


# kernel flavours per spec: 0 = 32-slot tiles (stream for the residence supergroup, gather
# otherwise), 1 = one warp per group
LAUNCH_MODES = {
    "tiles": {"household": 0, "company": 0, "care_home": 0},
    "warps": {"household": 1, "company": 1, "care_home": 1},
    "mixed": {"household": 1, "company": 0, "care_home": 0},
    "tiles_gather": {"household": 0, "company": 0, "care_home": 0, "_mirror": 0},
    "warps_gather": {"household": 1, "company": 1, "care_home": 1, "_mirror": 0},
}


@pytest.mark.gpu
@pytest.mark.parametrize("modes", ["tiles", "warps", "mixed", "tiles_gather", "warps_gather"])
@pytest.mark.parametrize("name", FIXTURES)
def test_cuda_matches_reference_trace(name, modes):
    assert have_gpu(), "CUDA library or device missing: the product path has no fallback"
    draws, infections, worst = run_trace("cuda", name, launch_modes=LAUNCH_MODES[modes])
    assert draws > 10000 and infections > 100
    assert worst <= REL
