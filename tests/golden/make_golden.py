"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(IDAS-Durham/JUNE, /root/reference) on a small programmatic world and tracing its hot path.

Run by hand in the build container (the reference cannot travel to the GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

What is traced, per ``Simulator.do_timestep`` call (june/simulator.py:346-626):
  * the step inputs: ``timer.now``, ``timer.duration``, ``timer.activities``, processed betas;
  * infections created *before* the step (seeding, infection_seed.py:209-278) as parameter rows;
  * the placement the reference's ActivityManager produced (which subgroup holds whom);
  * every Bernoulli draw of ``Interaction._gets_infected`` (interaction.py:634-641): the person,
    the exponent ``Lambda`` and the uniform consumed (the module-level ``random`` is replaced by an
    iterator over a fixed stream -- SURVEY.md section 4.4);
  * the infections created in the step (parameter rows sampled by the reference's scipy path);
  * the person state after ``Epidemiology.update_health_status``.
Also dumped: processed contact matrices / betas, the health-index table, gamma-pdf spot values.
"""
from __future__ import annotations

import contextlib
import io
import os
import random as pyrandom
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_world  # noqa: E402

from june_b200.adapter import compile_world  # noqa: E402
from june_b200.world import ACTIVITY_CODES, MAX_STAGES  # noqa: E402


def infection_row(pidx, person, time):
    inf = person.infection
    tr = inf.transmission
    name = type(tr).__name__
    if name == "TransmissionGamma":
        tpar = [tr.shape, tr.shift, tr.scale, tr.norm, 0.0]
    elif name == "TransmissionXNExp":
        tpar = [tr.n, tr.alpha, tr.time_first_infectious, tr.norm, tr.norm_time]
    else:
        tpar = [tr.probability, 0.0, 0.0, 0.0, 0.0]
    traj = inf.symptoms.trajectory
    tt = [float(t) for t, _ in traj] + [0.0] * (MAX_STAGES - len(traj))
    tg = [int(g) for _, g in traj] + [0] * (MAX_STAGES - len(traj))
    return [pidx[person.id], float(time), len(traj), int(inf.symptoms.max_tag), float(tr.probability)] + tpar + tt + tg


ROW_LEN = 5 + 5 + 2 * MAX_STAGES


def run_trace(n_people, total_days, seed, out_name, cases_per_capita=0.02, schedule="default", n_uniform=400000,
              with_leisure=False, lockdown=False, with_commute=False, vaccination=False, records=False, with_visits=False):
    pyrandom.seed(seed)
    np.random.seed(seed)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        world, dc = ref_world.build_world(n_people, seed=seed, with_leisure=with_leisure, with_commute=with_commute,
                                          with_visits=with_visits)
        sim = ref_world.make_simulator(world, dc, total_days=total_days, cases_per_capita=cases_per_capita,
                                       schedule=schedule, with_leisure=with_leisure, lockdown=lockdown,
                                       with_commute=with_commute, vaccination=vaccination, with_visits=with_visits)
    import june.interaction.interaction as inter_mod
    import june.policy.individual_policies as indiv_mod
    from june.epidemiology.infection import InfectionSelector
    from june.groups.company import InteractiveCompany
    from june.interaction import Interaction
    from june.simulator import Simulator

    # The "child cannot stay home alone -> fetch a guardian" rule is order dependent, draws from
    # Python's global `random.choice`, and is switched off by the reference itself whenever it runs
    # domain-decomposed (`if mpi_size == 1`, individual_policies.py:57-77).  The device path
    # replaces the domain-decomposed mode, so the traces are taken in that branch.
    indiv_mod.mpi_size = 2

    a2sg = {"medical_facility": ["hospitals"], "primary_activity": ["schools", "companies"],
            "residence": ["households", "care_homes"]}
    if with_leisure:
        a2sg["leisure"] = ["pubs", "cinemas", "groceries"] + (["household_visits", "care_home_visits"] if with_visits else [])
    if with_commute:
        a2sg["commute"] = ["city_transports", "inter_city_transports"]
    soa = compile_world(world, a2sg, InteractiveCompany.sector_betas, leisure=sim.activity_manager.leisure,
                        travel=sim.activity_manager.travel, long_commute_distance=150 if lockdown else None)
    people = list(world.people)
    pidx = {p.id: i for i, p in enumerate(people)}
    gidx = {}
    for sg in soa.supergroups:
        for k, g in enumerate(getattr(world, sg.name).members):
            gidx[id(g)] = sg.first_group + k

    # ---- instrumentation ---------------------------------------------------------------------
    rng = np.random.default_rng(1234 + seed)
    uniforms = rng.random(n_uniform)
    state = {"upos": 0, "draw_person": [], "draw_lambda": [], "draw_u": [], "rows": [], "pending": []}

    def fake_random():
        u = uniforms[state["upos"]]
        state["upos"] += 1
        state["draw_u"].append(u)
        return float(u)

    inter_mod.random = fake_random
    orig_gets = Interaction._gets_infected
    orig_sub = Interaction._time_step_for_subgroup
    orig_infect = InfectionSelector.infect_person_at_time

    def gets(self, params, infection_ids):
        state["draw_lambda"].append(float(params.sum()))
        return orig_gets(self, params, infection_ids)

    def sub(self, infector_tensor, susceptible_subgroup_id, subgroup_susceptibles):
        state["draw_person"].extend(pidx[i] for i in subgroup_susceptibles.keys())
        return orig_sub(self, infector_tensor=infector_tensor, susceptible_subgroup_id=susceptible_subgroup_id,
                        subgroup_susceptibles=subgroup_susceptibles)

    def infect(self, person, time):
        orig_infect(self, person, time)
        state["rows"].append(infection_row(pidx, person, time))

    Interaction._gets_infected = gets
    Interaction._time_step_for_subgroup = sub
    InfectionSelector.infect_person_at_time = infect

    # ---- individual policies: every random() they consume, tagged (policy class, draw, person) -----
    pol_state = {"name": None, "person": None, "j": 0, "draws": []}
    prng = np.random.default_rng(4321 + seed)

    def policy_random():
        u = float(prng.random())
        pol_state["draws"].append((pol_state["name"], pol_state["j"], pidx[pol_state["person"].id], u))
        pol_state["j"] += 1
        return u

    if lockdown:
        indiv_mod.random = policy_random
        for pol in sim.activity_manager.policies.individual_policies.policies:
            for meth in ("check_stay_home_condition", "check_skips_activity"):
                if hasattr(pol, meth):
                    def wrapped(person, *a, _f=getattr(pol, meth), _n=type(pol).__name__, **kw):
                        pol_state["name"], pol_state["person"], pol_state["j"] = _n, person, 0
                        return _f(person, *a, **kw)
                    setattr(pol, meth, wrapped)

    # ---- vaccination: the uniform each person consumes in VaccinationCampaigns.apply -------------
    vac_state = {"person": None, "u": {}}
    if vaccination:
        import june.epidemiology.vaccines.vaccination_campaign as vac_mod
        vrng = np.random.default_rng(99 + seed)

        def vac_random():
            u = float(vrng.random())
            vac_state["u"][pidx[vac_state["person"].id]] = u
            return u

        vac_mod.random = vac_random
        orig_apply = vac_mod.VaccinationCampaigns.apply

        def apply(self, person, date, record=None):
            vac_state["person"] = person
            return orig_apply(self, person=person, date=date, record=record)

        vac_mod.VaccinationCampaigns.apply = apply

    # ---- records: every Record.accumulate call of the step, through a duck-typed recorder ----------
    rec_state = {"rows": [], "inf": []}
    if records:
        kinds = {"deaths": 1, "recoveries": 2, "hospital_admissions": 3, "icu_admissions": 4, "discharges": 5, "symptoms": 6}
        gid_of = {}
        for sg in soa.supergroups:
            for g in getattr(world, sg.name).members:
                gid_of[(g.spec, g.id)] = gidx[id(g)]

        import types
        from june.records.records_writer import Record as RefRecord

        def fresh_events():
            return {"infections": types.SimpleNamespace(region_names=[]),
                    "hospital_admissions": types.SimpleNamespace(hospital_ids=[]),
                    "icu_admissions": types.SimpleNamespace(hospital_ids=[]),
                    "deaths": types.SimpleNamespace(dead_person_ids=[], location_specs=[], location_ids=[])}

        class Recorder:
            record_static_data = False
            mpi_rank = None
            events = fresh_events()

            def accumulate(self, table_name, **kw):
                # what the reference's event tables keep for summarise_time_step (event_records_writer.py)
                if table_name == "infections":
                    self.events["infections"].region_names.extend([kw["region_name"]] * len(kw["infected_ids"]))
                elif table_name in ("hospital_admissions", "icu_admissions"):
                    self.events[table_name].hospital_ids.append(kw["hospital_id"])
                elif table_name == "deaths":
                    self.events["deaths"].dead_person_ids.append(kw["dead_person_id"])
                    self.events["deaths"].location_specs.append(kw["location_spec"])
                    self.events["deaths"].location_ids.append(kw["location_id"])
                if table_name == "infections":
                    if kw["location_spec"] == "infection_seed":
                        return
                    g = gid_of[(kw["location_spec"], kw["location_id"])]
                    for a, b in zip(kw["infected_ids"], kw["infector_ids"]):
                        rec_state["inf"].append((pidx[a], pidx[b], g))
                elif table_name == "deaths":
                    rec_state["rows"].append((pidx[kw["dead_person_id"]], 1, gid_of[(kw["location_spec"], kw["location_id"])]))
                elif table_name == "recoveries":
                    rec_state["rows"].append((pidx[kw["recovered_person_id"]], 2, -1))
                elif table_name in ("hospital_admissions", "icu_admissions", "discharges"):
                    rec_state["rows"].append((pidx[kw["patient_id"]], kinds[table_name], gid_of[("hospital", kw["hospital_id"])]))
                elif table_name == "symptoms":
                    rec_state["rows"].append((pidx[kw["infected_id"]], 6, int(kw["symptoms"])))

            def summarise_time_step(self, timestamp, world):
                # the reference's OWN summary functions (records_writer.py:248-376) on the events of this step:
                # the eight columns of a summary.csv row (:764-773), per region
                di, ci = RefRecord.summarise_infections(self, world=world)
                dh, dicu, ch, cicu = RefRecord.summarise_hospitalisations(self, world=world)
                dd, ddh = RefRecord.summarise_deaths(self, world=world)
                rec_state["summary"] = [[d.get(r, 0) for d in (ci, di, ch, dh, cicu, dicu, ddh, dd)] for r in soa.region_names]

            def time_step(self, timestamp):
                type(self).events = fresh_events()

            def parameters(self, **kw):
                pass

            def combine_outputs(self):
                pass

        sim.record = Recorder()

    steps = []
    orig_step = Simulator.do_timestep
    orig_am = type(sim.activity_manager).do_timestep
    placement = {}

    def am_do_timestep(self, record=None):
        ret = orig_am(self, record=record)
        pg = np.full(len(people), -1, np.int32)
        pl = np.zeros(len(people), np.uint8)
        for sg in soa.supergroups:
            for g in getattr(world, sg.name).members:
                for li, s in enumerate(g.subgroups):
                    for person in s.people:
                        pg[pidx[person.id]] = gidx[id(g)]
                        pl[pidx[person.id]] = li
        placement["group"], placement["lsg"] = pg, pl
        return ret

    type(sim.activity_manager).do_timestep = am_do_timestep

    def snapshot():
        n = len(people)
        inf = np.zeros(n, np.uint8); dead = np.zeros(n, np.uint8); tag = np.full(n, -1, np.int8)
        trans = np.zeros(n); susc = np.zeros(n); med = np.full(n, -1, np.int32); stage = np.zeros(n, np.uint8)
        iid = sim.epidemiology.infection_selectors[0].infection_class.infection_id()
        for i, p in enumerate(people):
            dead[i] = p.dead
            susc[i] = p.immunity.get_susceptibility(iid)
            if p.infection is not None:
                inf[i] = 1
                tag[i] = int(p.infection.tag)
                trans[i] = p.infection.transmission.probability
                stage[i] = p.infection.symptoms.stage
            mf = p.subgroups.medical_facility if p.subgroups is not None else None
            if mf is not None:
                med[i] = (gidx[id(mf.group)] << 8) | mf.group.subgroups.index(mf)
        out = dict(infected=inf, dead=dead, tag=tag, trans=trans, susc=susc, med=med, stage=stage)
        if vaccination:
            out["vac_dose"] = np.array([-1 if p.vaccinated is None else int(p.vaccinated) for p in people], np.int8)
            out["effmult"] = np.array([p.immunity.effective_multiplier_dict.get(iid, 1.0) for p in people])
        return out

    # residence visits: the leisure destination every person DREW at its turn (leisure.py:230-275) -- the final placement
    # does not tell who went out and was fetched back when a visitor came to its household (household.py:124-149)
    lz_state = {"choice": None}
    if with_visits:
        leisure_cls = type(sim.activity_manager.leisure)
        orig_lz = leisure_cls.get_subgroup_for_person_and_housemates

        def lz(self, person, to_send_abroad=None):
            sub = orig_lz(self, person=person, to_send_abroad=to_send_abroad)
            if sub is not None:
                lz_state["choice"][pidx[person.id]] = gidx[id(sub.group)]
            return sub
        leisure_cls.get_subgroup_for_person_and_housemates = lz

    def do_timestep(self):
        rec = {}
        lz_state["choice"] = np.full(len(people), -1, np.int32)
        rec["pre_rows"] = np.array(state["rows"], np.float64).reshape(-1, ROW_LEN)
        state["rows"] = []
        state["draw_person"], state["draw_lambda"], state["draw_u"] = [], [], []
        pol_state["draws"] = []
        vac_state["u"] = {}
        rec_state["rows"], rec_state["inf"] = [], []
        timer = self.timer
        vac_day = self.epidemiology.current_date is None or timer.date.date() != self.epidemiology.current_date.date()
        rec["now"], rec["dt"] = float(timer.now), float(timer.duration)
        mask = 0
        for a in timer.activities:
            mask |= 1 << ACTIVITY_CODES[a]
        rec["activity_mask"] = mask
        orig_step(self)
        rec["place_group"], rec["place_lsg"] = placement["group"], placement["lsg"]
        if with_visits:
            rec["lz_choice"] = lz_state["choice"]
        rec["draw_person"] = np.array(state["draw_person"], np.int64)
        rec["draw_lambda"] = np.array(state["draw_lambda"], np.float64)
        rec["draw_u"] = np.array(state["draw_u"], np.float64)
        rec["new_rows"] = np.array(state["rows"], np.float64).reshape(-1, ROW_LEN)
        state["rows"] = []
        if lockdown:
            names = sorted({d[0] for d in pol_state["draws"]})
            rec["pol_names"] = np.array(names if names else ["-"])
            rec["pol_draws"] = np.array([[names.index(d[0]), d[1], d[2], d[3]] for d in pol_state["draws"]],
                                        np.float64).reshape(-1, 4)
            rec["pol_active"] = np.array([type(p).__name__ for p in
                                          self.activity_manager.policies.individual_policies.get_active(date=timer.date).policies] or ["-"])
        if vaccination and vac_day:
            rec["vac_today"] = np.array((timer.date.date() - self.timer.initial_date.date()).days)
            u = np.full(len(people), np.nan)
            for k, v in vac_state["u"].items():
                u[k] = v
            rec["vac_u"] = u
        if records:
            rec["health_events"] = np.array(sorted(rec_state["rows"]), np.int64).reshape(-1, 3)
            rec["infection_events"] = np.array(sorted(rec_state["inf"]), np.int64).reshape(-1, 3)
            rec["summary"] = np.array(rec_state.get("summary", []), np.int64).reshape(-1, 8)
        rec["post"] = snapshot()
        steps.append(rec)

    Simulator.do_timestep = do_timestep
    try:
        with contextlib.redirect_stdout(sink):
            sim.run()
    finally:
        Simulator.do_timestep = orig_step
        type(sim.activity_manager).do_timestep = orig_am
        Interaction._gets_infected = orig_gets
        Interaction._time_step_for_subgroup = orig_sub
        InfectionSelector.infect_person_at_time = orig_infect
        if with_visits:
            leisure_cls.get_subgroup_for_person_and_housemates = orig_lz
        from random import random as real_random
        inter_mod.random = real_random

    # ---- reference constants -------------------------------------------------------------------
    interaction = sim.interaction
    hig = sim.epidemiology.infection_selectors[0].health_index_generator
    prob = np.zeros((2, 2, 100, 9))
    for pi, pop in enumerate(("gp", "ch")):
        for si, sx in enumerate(("f", "m")):
            prob[pi, si] = hig.probabilities[pop][sx][:100]
    out = {}
    soa.save(os.path.join(HERE, out_name + "_world.npz"))
    out["n_steps"] = np.array(len(steps))
    out["health_index_prob"] = prob
    for k, m in interaction.contact_matrices.items():
        out["cm_" + k] = np.asarray(m, np.float64)
    if lockdown:
        from june.policy import CloseCompanies
        out["close_companies_ratios"] = np.array([CloseCompanies.furlough_ratio, CloseCompanies.key_ratio,
                                                  CloseCompanies.random_ratio], np.float64)
    out["beta_keys"] = np.array(list(interaction.betas.keys()))
    out["beta_vals"] = np.array(list(interaction.betas.values()), np.float64)
    for i, rec in enumerate(steps):
        for k, v in rec.items():
            if k == "post":
                for kk, vv in v.items():
                    out[f"s{i}_post_{kk}"] = vv
            else:
                out[f"s{i}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, out_name + "_trace.npz"), **out)
    n_draws = sum(len(r["draw_u"]) for r in steps)
    n_inf = sum(len(r["new_rows"]) for r in steps)
    print(f"{out_name}: {len(people)} people, {soa.n_groups} groups, {len(steps)} steps, {n_draws} draws, "
          f"{n_inf} infections in steps, {int(steps[-1]['post']['dead'].sum())} dead")
    return world, sim, soa


def spot_values(out_name="spot"):
    """Known-answer values of the scalar functions on the path."""
    ref_world.ref_harness.setup()
    from june.epidemiology.infection.transmission import TransmissionGamma, gamma_pdf
    from june.epidemiology.infection.transmission_xnexp import TransmissionXNExp
    from june.groups.school import InteractiveSchool

    rng = np.random.default_rng(7)
    xs = rng.uniform(-1, 25, 200); a = rng.normal(1.56, 0.08, 200); loc = rng.normal(-2.12, 0.1, 200) + rng.uniform(1, 8, 200)
    sc = 1.0 / rng.normal(0.53, 0.03, 200)
    g = np.array([gamma_pdf(float(x), float(aa), float(ll), float(ss)) for x, aa, ll, ss in zip(xs, a, loc, sc)])
    xn_in, xn_out = [], []
    for _ in range(100):
        mp, tfi, nt, n, al = rng.uniform(0.2, 2), rng.uniform(0.5, 4), rng.uniform(0.5, 2), rng.uniform(0.5, 3), rng.uniform(1, 6)
        t = TransmissionXNExp(max_probability=mp, time_first_infectious=tfi, norm_time=nt, n=n, alpha=al)
        tt = rng.uniform(0, 15)
        t.update_infection_probability(tt)
        xn_in.append([n, al, tfi, t.norm, nt, tt, mp]); xn_out.append(t.probability)
    tg = TransmissionGamma(max_infectiousness=1.3, shape=1.56, rate=0.53, shift=2.88)
    tg.update_infection_probability(5.5)
    np.savez_compressed(os.path.join(HERE, out_name + ".npz"), gamma_x=xs, gamma_a=a, gamma_loc=loc, gamma_scale=sc,
                        gamma_pdf=g, xnexp_in=np.array(xn_in), xnexp_out=np.array(xn_out),
                        tg_55=np.array(tg.probability))
    print("spot values written")


LEISURE_POLICY_CASE = dict(
    venues_to_close=["cinema"],
    activity_reductions={"pub": {"weekday": {"male": {"0-50": 0.5, "50-100": 0.2}, "female": {"0-70": 0.3, "70-100": 0.8}},
                                 "weekend": {"both_sexes": {"0-100": 0.6}}},
                         "grocery": {"both_sexes": {"0-65": 0.9, "65-100": 0.4}}},
    compliance=[1.0, 0.7, 0.35],
    slots=[("2020-03-03 09:00", 8 / 24, True), ("2020-03-03 18:00", 3 / 24, False), ("2020-03-07 12:00", 4 / 24, False)],
)


def leisure_policy_tables(out_name="leisure_policy_tables"):
    """The reference's leisure probabilities with CloseLeisureVenue + ChangeLeisureProbability active and
    region-dependent compliance (leisure_policies.py, social_venue_distributor.py:176-205, leisure.py:205-228)."""
    import datetime
    sink = io.StringIO()
    pyrandom.seed(2); np.random.seed(2)
    with contextlib.redirect_stdout(sink):
        world, dc = ref_world.build_world(2500, seed=2, with_leisure=True)
        sim = ref_world.make_simulator(world, dc, total_days=2, with_leisure=True)
    from june.policy import ChangeLeisureProbability, CloseLeisureVenue, LeisurePolicies
    leisure = sim.activity_manager.leisure
    case = LEISURE_POLICY_CASE
    pols = LeisurePolicies([CloseLeisureVenue("2020-01-01", "2021-01-01", venues_to_close=case["venues_to_close"]),
                            ChangeLeisureProbability("2020-01-01", "2021-01-01", activity_reductions=case["activity_reductions"])])
    regions = list(world.regions)
    for k, region in enumerate(regions):
        region.regional_compliance = case["compliance"][k % len(case["compliance"])]
    acts = list(leisure.leisure_distributors)
    does = np.zeros((len(case["slots"]), len(regions), 2, 100))
    probs = np.zeros((len(case["slots"]), len(regions), 2, 100, len(acts)))
    for t, (when, dt, wh) in enumerate(case["slots"]):
        date = datetime.datetime.strptime(when, "%Y-%m-%d %H:%M")
        with contextlib.redirect_stdout(sink):
            pols.apply(date=date, leisure=leisure)
            leisure.generate_leisure_probabilities_for_timestep(delta_time=dt, working_hours=wh, date=date)
        table = leisure.probabilities_by_region_sex_age
        for r, region in enumerate(regions):
            for si, sex in enumerate("fm"):
                for age in range(100):
                    e = table[region.name][sex][age]
                    does[t, r, si, age] = e["does_activity"]
                    probs[t, r, si, age] = [e["activities"][a] for a in acts]
    np.savez_compressed(os.path.join(HERE, out_name + ".npz"), does=does, probs=probs, activities=np.array(acts),
                        regions=np.array([r.name for r in regions]))
    print("leisure policy tables:", acts, does.shape, float(does.mean()))


if __name__ == "__main__":
    if "leisure_policies" in sys.argv[1:]:
        leisure_policy_tables()
        sys.exit(0)
    if "visits" in sys.argv[1:]:  # + household and care-home visits (the residents of a visited household stay / come home)
        run_trace(2500, 9, 8, "w2k_visits", cases_per_capita=0.05, with_leisure=True, with_visits=True)
    if "leisure" in sys.argv[1:]:  # leisure placement (pubs, cinemas, groceries) on top of the w3k recipe
        run_trace(2500, 9, 2, "w2k_leisure", cases_per_capita=0.04, with_leisure=True)
        sys.exit(0)
    if "vaccination" in sys.argv[1:]:  # two campaigns, doses 0 and 1, waning
        run_trace(2500, 14, 5, "w2k_vaccination", cases_per_capita=0.04, vaccination=True)
        sys.exit(0)
    if "commute" in sys.argv[1:]:  # city / inter-city transports on the two commute steps of a weekday
        run_trace(2500, 9, 4, "w2k_commute", cases_per_capita=0.05, with_commute=True)
        sys.exit(0)
    if "records" in sys.argv[1:]:  # the Record.accumulate rows of every step (infections, symptoms, deaths, ...)
        run_trace(2000, 30, 6, "w2k_records", cases_per_capita=0.15, schedule="simple", records=True)
        sys.exit(0)
    if "policies" in sys.argv[1:]:  # Quarantine, Shielding, CloseSchools, CloseCompanies switching on day by day
        run_trace(2500, 9, 3, "w2k_policies", cases_per_capita=0.05, lockdown=True)
        sys.exit(0)
    run_trace(3000, 12, 0, "w3k")
    run_trace(1200, 40, 1, "w1k_long", cases_per_capita=0.05, schedule="simple")
    spot_values()
