"""Drop-in for a real JUNE installation: ``make_b200_simulator_class(june.simulator.Simulator)``
returns a subclass whose ``do_timestep`` runs on the B200 path.

This is the binding a JUNE maintainer adds.  It imports nothing from JUNE
itself -- the reference class is passed in -- so this package stays importable without JUNE.

What it does
------------
* ``__init__`` / ``from_file`` are the reference's own (``june/simulator.py:82-303``): same
  arguments, same YAML.  On the first step the world object graph is compiled ONCE into the SoA
  form (``june_b200.adapter.compile_world``) and uploaded.
* ``do_timestep`` (``simulator.py:346-626``): applies the reference's own interaction / regional
  policies to get ``beta_reductions`` and compliances (host scalars), builds the per-step beta
  table with the reference ``Interaction.betas``, and calls ``jb_step``.
* Infections the reference creates on ``Person`` objects (seeding through its ``InfectionSeed``,
  ``infection_seed.py:209-278``) are caught where they are made (``InfectionSelector.infect_person_at_time``) and
  uploaded before the step; after the step the people whose state CHANGED -- the step's new infections and its
  health events, read as compact lists from the device -- get light-weight views written back
  (``person.infection``, ``person.dead``, immunity), so that seeding, records and checkpoints keep working on the
  reference's objects.  Per-step host work is proportional to what changed, not to the population;
  ``b200_sync_people()`` refreshes every person on demand (transmission probabilities, waning vaccine effects).
* ``b200_reference_rng = True`` makes a run reproduce the reference's own run bit for bit: the Bernoulli uniforms
  are taken from Python's ``random`` in the reference's draw order, and the new infections are created by the
  reference's own ``InfectionSelector`` (numpy / scipy) on the host and uploaded.  Slower (test and validation mode).
"""
from __future__ import annotations

import os
import sys
from typing import Optional

import numpy as np

from . import _lib
from .adapter import compile_world
from .device import DeviceWorld, Engine
from .disease import DiseaseConfig, outcome_probabilities, rates_from_dataframe
from .interaction import Interaction
from .world import ACTIVITY_CODES, PS_DEAD, PS_INFECTED


class _TransmissionView:
    __slots__ = ("probability",)

    def __init__(self, probability):
        self.probability = probability


class _SymptomsView:
    __slots__ = ("tag", "max_tag", "stage", "trajectory", "time_of_symptoms_onset")

    def __init__(self, tag, max_tag, stage, trajectory):
        self.tag, self.max_tag, self.stage, self.trajectory = tag, max_tag, stage, trajectory
        self.time_of_symptoms_onset = None


def make_b200_simulator_class(reference_simulator_cls):
    class B200Simulator(reference_simulator_cls):
        """``Simulator`` with ``do_timestep`` on the B200 path."""

        b200_engine: Optional[Engine] = None
        b200_device: int = 0
        b200_seed: int = 0
        b200_reference_rng: bool = False   # uniforms from Python's `random`, infections from the reference's sampler

        # ---- one-off compilation ---------------------------------------------------------------
        def _b200_setup(self):
            am = self.activity_manager
            a2sg = {k: v for k, v in am.activity_to_super_group_dict.items()}
            sector_betas = {}
            for g in getattr(getattr(self.world, "companies", None), "members", []) or []:
                ig = g.get_interactive_group(None)
                sector_betas = dict(type(ig).sector_betas)
                break
            llc = [p for p in (am.policies.individual_policies.policies if am.policies is not None else [])
                   if type(p).__name__ == "LimitLongCommute"]
            self._soa = compile_world(self.world, a2sg, sector_betas, leisure=am.leisure, travel=am.travel,
                                      long_commute_distance=type(llc[0]).apply_from_distance if llc else None)
            self._leisure_cache = {}
            self._leisure_stack = []
            self._policy_key = None
            ref_inter = self.interaction
            self._inter = Interaction.__new__(Interaction)
            self._inter.alpha_physical = ref_inter.alpha_physical
            self._inter.betas = dict(ref_inter.betas)
            self._inter.contact_matrices = {k: np.asarray(v, np.float64) for k, v in ref_inter.contact_matrices.items()}
            self._inter.beta_reductions = {}
            selector = self.epidemiology.infection_selectors[0]
            self._selector = selector
            dc_ref = selector.disease_config
            # the reference pops keys out of its parsed YAML while building CompletionTime objects
            # (trajectory_maker.py:59-60), so re-read the file it came from
            self._dc = None
            paths = sys.modules.get("june.paths")
            if paths is not None:
                f = os.path.join(str(paths.configs_path), "defaults", "epidemiology", "infection", "disease",
                                 f"{dc_ref.disease_name.lower()}.yaml")
                if os.path.exists(f):
                    self._dc = DiseaseConfig.from_file(f)
            if self._dc is None:
                self._dc = DiseaseConfig(dc_ref.disease_name)
            hig = selector.health_index_generator
            rates, bins = rates_from_dataframe(hig.rates_df)
            prob = outcome_probabilities(self._dc, rates, bins)
            if self.b200_engine is None:
                type(self).b200_engine = None
                self._eng = DeviceWorld(self._soa, self._inter, self._dc, prob, device=self.b200_device)
            else:
                self._eng = self.b200_engine(self._soa, self._inter, self._dc, prob)
            self._people = list(self.world.people)
            self._index = {id(p): i for i, p in enumerate(self._people)}
            self._known_infected = np.zeros(len(self._people), bool)
            self._step = 0
            # infections the reference creates from now on (seeds) are caught where they are made; the ones that
            # exist already are found by one scan
            self._pending = [p for p in self._people if p.infection is not None]
            for sel in self.epidemiology.infection_selectors:
                if getattr(sel, "_b200_wrapped", False):
                    continue

                def wrapped(person, time, _orig=sel.infect_person_at_time, _sim=self):
                    _orig(person=person, time=time)
                    _sim._pending.append(person)
                sel.infect_person_at_time = wrapped
                sel._b200_wrapped = True

        # ---- host -> device: infections created by the reference (seeding) -----------------------
        def _b200_upload_new_reference_infections(self):
            rows = []
            for p in self._pending:
                i = self._index.get(id(p))
                inf = p.infection
                if i is None or inf is None or self._known_infected[i] or isinstance(inf, _InfectionView):
                    continue
                tr = inf.transmission
                name = type(tr).__name__
                if name == "TransmissionGamma":
                    tpar = [tr.shape, tr.shift, tr.scale, tr.norm, 0.0]
                elif name == "TransmissionXNExp":
                    tpar = [tr.n, tr.alpha, tr.time_first_infectious, tr.norm, tr.norm_time]
                else:
                    tpar = [tr.probability, 0, 0, 0, 0]
                traj = inf.symptoms.trajectory
                rows.append(dict(person=i, start_time=inf.start_time, tpar=tpar, probability=tr.probability,
                                 traj_time=[t for t, _ in traj], traj_tag=[int(g) for _, g in traj],
                                 stage=inf.symptoms.stage, max_tag=int(inf.symptoms.max_tag), tag=int(inf.symptoms.tag)))
                self._known_infected[i] = True
            self._pending = []
            if rows:
                self._eng.upload_infections(rows)

        # ---- device -> host: views on the reference objects ---------------------------------------
        def _b200_write_back(self, new_members, hev):
            """The people whose state changed in this step: its new infections (device-created ones get a view) and
            its health events (symptom tag changes, recoveries, deaths)."""
            cls = self._selector.infection_class
            exposed = int(self._dc.symptom_tags.get(self._dc.default_lowest_stage, 0))
            for i in new_members:
                i = int(i)
                if i >= len(self._people):
                    continue  # a visitor from another domain
                p = self._people[i]
                if p.infection is None:
                    p.infection = _InfectionView(cls, 0.0, exposed)
                    p.immunity.add_immunity(p.infection.immunity_ids())
                self._known_infected[i] = True
            for i, k, g in zip(hev["person"], hev["kind"], hev["group"]):
                p, k = self._people[int(i)], int(k)
                if k == 1:
                    p.dead, p.infection = True, None
                    self._known_infected[int(i)] = False
                elif k == 2:
                    p.infection = None
                    self._known_infected[int(i)] = False
                elif k == 6 and p.infection is not None:
                    p.infection.symptoms.tag = int(g)

        def b200_sync_people(self):
            """Refresh every person from the device (O(population); not part of a step): transmission probability and
            tag of the infected, susceptibility and effective multiplier of everybody."""
            st = self._eng.fetch_person_state()
            vs = self._eng.fetch_vaccination()
            iid = self._selector.infection_class.infection_id()
            for i, p in enumerate(self._people):
                inf = p.infection
                if inf is not None:
                    inf.transmission.probability = float(st["trans_prob"][i])
                    inf.symptoms.tag = int(st["tag"][i])
                p.immunity.susceptibility_dict[iid] = float(st["susceptibility"][0][i])
                if vs["dose"][i] >= 0:
                    p.immunity.effective_multiplier_dict[iid] = float(vs["effective_multiplier"][0][i])

        # ---- leisure tables from the reference's own Leisure object ----------------------------------
        def _b200_leisure_table(self, activities, pol):
            """``ActivityManager.do_timestep`` (``activity_manager.py:200-209``): leisure policies, then
            ``generate_leisure_probabilities_for_timestep``.  The reference's nested dict is flattened
            into ``does[R][2][100]`` / ``cum[R][2][100][K]`` over the device's activities (venues and residence
            visits); ``ChangeVisitsProbability`` reaches the device as the residence-type split in force."""
            leisure = self.activity_manager.leisure
            if leisure is None or not self._soa.has_leisure or "leisure" not in activities:
                return 0
            date = self.timer.date
            if pol is not None:
                pol.leisure_policies.apply(date=date, leisure=leisure)
            if self._soa.has_visits:
                dist = leisure.leisure_distributors["residence_visits"]
                split = dist.policy_reductions or dist.residence_type_probabilities  # residence_visits.py:305-309
                split = (float(split.get("household", 0.0)), float(split.get("care_home", 0.0)))
                if split != getattr(self, "_b200_visit_split", None):
                    self._eng.set_visit_probabilities(*split)
                    self._b200_visit_split = split
            lkey = tuple(id(p) for p in pol.leisure_policies.get_active(date)) if pol is not None else ()
            key = (round(self.timer.duration, 9), "primary_activity" in activities, date.weekday(), date.hour, self._policy_key, lkey)
            idx = self._leisure_cache.get(key)
            if idx is None:
                leisure.generate_leisure_probabilities_for_timestep(
                    delta_time=self.timer.duration, date=date, working_hours="primary_activity" in activities)
                table = leisure.probabilities_by_region_sex_age
                specs, R = self._soa.leisure_activities, self._soa.n_regions
                does, cum = np.zeros((R, 2, 100)), np.zeros((R, 2, 100, len(specs)))
                for r, name in enumerate(self._soa.region_names):
                    per = table.get(name, table) if isinstance(table, dict) and name in table else table
                    for si, sex in enumerate("fm"):
                        for age in range(100):
                            e = per[sex][age]
                            probs = np.array([e["activities"].get(sp, 0.0) for sp in specs], np.float64)
                            kept = probs.sum()
                            does[r, si, age] = e["does_activity"] * kept
                            cum[r, si, age] = np.cumsum(probs / kept) if kept > 0 else 1.0
                self._leisure_stack.append((does, cum))
                idx = self._leisure_cache[key] = len(self._leisure_stack) - 1
                self._eng.set_leisure_tables(np.stack([t[0] for t in self._leisure_stack]),
                                             np.stack([t[1] for t in self._leisure_stack]))
            return idx + 1

        # ---- individual policies of the reference -> device records -----------------------------------
        def _b200_policy_records(self, pol, comp, tiers=()):
            date = self.timer.date
            active = [p for p in pol.individual_policies.policies if p.is_active(date)] if pol is not None else []
            key = tuple(id(p) for p in active) + tuple(comp) + tuple(tiers)
            if key == self._policy_key:
                return
            nan = float("nan")

            def opt(x):
                return nan if x is None else x

            recs = []
            for p in active:
                name = type(p).__name__
                if name == "Quarantine":
                    recs.append(dict(type=_lib.POL_QUARANTINE, mask=self._dc.tag_mask("stay_at_home_stage"), p=[p.n_days, p.compliance]))
                elif name == "Shielding":
                    recs.append(dict(type=_lib.POL_SHIELDING, p=[p.min_age, opt(p.compliance)]))
                elif name == "CloseSchools":
                    mask = 0
                    for y in p.years_to_close or []:
                        if 0 <= int(y) < 32:
                            mask |= 1 << int(y)
                    recs.append(dict(type=_lib.POL_CLOSE_SCHOOLS, flags=1 if p.full_closure else 0, mask=mask, p=[p.attending_compliance]))
                elif name == "CloseUniversities":
                    recs.append(dict(type=_lib.POL_CLOSE_UNIVERSITIES))
                elif name == "CloseCompanies":
                    recs.append(dict(type=_lib.POL_CLOSE_COMPANIES, flags=1 if p.full_closure else 0,
                                     p=[opt(p.avoid_work_probability), opt(p.furlough_probability), opt(p.key_probability),
                                        opt(type(p).furlough_ratio), opt(type(p).key_ratio), opt(type(p).random_ratio)]))
                elif name == "LimitLongCommute":
                    recs.append(dict(type=_lib.POL_LIMIT_LONG_COMMUTE, p=[p.going_to_work_probability]))
                elif name == "CloseCompaniesLockdownTiers":
                    recs.append(dict(type=_lib.POL_CLOSE_COMPANIES_TIERS,
                                     mask=sum(1 << r for r, t in enumerate(tiers) if t in type(p).TIERS)))
                elif name != "SevereSymptomsStayHome":
                    raise NotImplementedError(f"individual policy {name} has no device implementation (INTEGRATION.md section 4)")
            if recs or self._policy_key is not None:
                self._eng.set_individual_policies(recs, comp)
            self._policy_key = key

        # ---- the replaced method --------------------------------------------------------------------
        def do_timestep(self):
            if not hasattr(self, "_eng"):
                self._b200_setup()
            pol = self.activity_manager.policies
            if pol is not None:
                pol.interaction_policies.apply(date=self.timer.date, interaction=self.interaction)
                pol.regional_compliance.apply(date=self.timer.date, regions=self.world.regions)
            activities = self.timer.activities
            if getattr(self, "events", None) is not None:  # simulator.py:368-376
                self._b200_apply_events(activities)
            if not activities:
                return
            self._b200_upload_new_reference_infections()
            self._inter.beta_reductions = dict(self.interaction.beta_reductions)
            regions = {r.name: r for r in self.world.regions}
            comp = [regions[n].regional_compliance if n in regions else 1.0 for n in self._soa.region_names]
            tiers = [regions[n].policy.get("lockdown_tier") if n in regions else None for n in self._soa.region_names]
            beta = self._inter.beta_table(self._soa.beta_rules, comp, tiers)
            mask = 0
            for a in activities:
                if a in ACTIVITY_CODES:
                    mask |= 1 << ACTIVITY_CODES[a]
            # the step's health events (and, with a record, who infected whom) are what the write-back reads
            flags = _lib.STEP_RECORD_EVENTS if self.record is not None else _lib.STEP_RECORD_HEALTH
            if pol is not None:
                date = self.timer.date
                if any(type(p).__name__ == "SevereSymptomsStayHome" and p.is_active(date) for p in pol.individual_policies.policies):
                    flags |= _lib.STEP_SEVERE_STAY_HOME
                if any(type(p).__name__ == "Hospitalisation" and p.is_active(date) for p in pol.medical_care_policies.policies):
                    flags |= _lib.STEP_HOSPITALISATION
            self._b200_policy_records(pol, comp, tiers)
            table = self._b200_leisure_table(activities, pol)
            eng = self._eng
            if not self.b200_reference_rng:
                params = eng.make_params(now=self.timer.now, delta_time=self.timer.duration, step=self._step,
                                         activity_mask=mask, seed=self.b200_seed, flags=flags, beta=beta, leisure_table=table)
                self.b200_last_stats = eng.step(params)
                new_members = eng.fetch_new_infections()[0]
            else:
                new_members = self._b200_step_with_reference_rng(mask, flags, beta, table)
            self._step += 1
            self._b200_write_back(new_members, eng.fetch_health_events())
            self._b200_vaccinate()
            if self.record is not None:
                self._b200_feed_record()

        def _b200_apply_events(self, activities):
            """``Events.apply`` (``event/event.py:93-107``).  The reference's events act on its Python objects
            (``Mutation`` swaps ``person.infection``, ``DomesticCare`` moves people between households, ...), which
            the device does not see: an active event is either one that brings its device form -- a method
            ``b200_apply(simulator)``, e.g. calling ``simulator._eng.mutate`` -- or an error, never silently dropped."""
            date = self.timer.date
            for event in (self.events.events or []):
                if not event.is_active(date=date):
                    continue
                hook = getattr(event, "b200_apply", None)
                if hook is None:
                    raise NotImplementedError(
                        f"event {type(event).__name__} is active on {date} and has no device form (INTEGRATION.md "
                        "section 4): give it a b200_apply(simulator) method, or run without it")
                hook(self)

        def _b200_step_with_reference_rng(self, mask, flags, beta, table):
            """The reference's own random streams: every Bernoulli draw takes the next ``random.random()`` in the
            reference's traversal order (``interaction.py:610-641``), and the newly infected get their infection from
            the reference's ``InfectionSelector`` in the order ``Epidemiology.infect_people`` walks them
            (``epidemiology.py:317-322``) -- a run equals the reference's run with the same seeds."""
            import random
            eng, soa = self._eng, self._soa
            flags |= _lib.STEP_INJECT_UNIFORMS | _lib.STEP_NO_NEW_INFECTIONS | _lib.STEP_RECORD_LAMBDA | _lib.STEP_RECORD_EVENTS
            state = random.getstate()
            cap = len(self._people)  # nobody draws twice in a step
            eng.inject_uniforms(np.array([random.random() for _ in range(cap)]))
            params = eng.make_params(now=self.timer.now, delta_time=self.timer.duration, step=self._step,
                                     activity_mask=mask, seed=self.b200_seed, flags=flags, beta=beta, leisure_table=table)
            eng.step_begin(params)
            eng.step_interact()
            n_draws = int((~np.isnan(eng.fetch_lambda()[:cap])).sum())
            random.setstate(state)  # the reference consumed exactly one uniform per draw
            for _ in range(n_draws):
                random.random()
            ev = eng.fetch_infection_events()
            # order of Simulator.do_timestep's infected_ids: group by group, inside a group in registration order
            keys = []
            for person, g in zip(ev["infected"], ev["group"]):
                lo, hi = int(soa.grp_offset[g]), int(soa.grp_offset[g + 1])
                pos = np.nonzero(soa.reg_member[lo:hi] == person)[0]
                keys.append((int(g), int(pos[0]) if len(pos) else hi - lo + int(person)))
            order = sorted(range(len(keys)), key=keys.__getitem__)
            # numpy's global stream: the reference draws two numbers per new infection while it interacts (who is to
            # blame: the infector's subgroup and the infector, interaction.py:643-662 -- np.random.choice with p takes
            # one uniform), before any infection is created
            np.random.random_sample(2 * len(order))
            sel = self.epidemiology.infection_selectors
            iid = self._selector.infection_class.infection_id()
            for k in order:
                i = int(ev["infected"][k])
                if i < len(self._people):
                    sel.infect_person_at_time(person=self._people[i], time=self.timer.now, infection_id=iid)
            self._b200_upload_new_reference_infections()
            self.b200_last_stats = eng.step_finish()
            return ev["infected"]

        # ---- the reference's vaccination campaigns -> device tables ----------------------------------
        def _b200_vaccinate(self):
            """``Epidemiology.do_timestep`` applies the campaigns once per simulated day, after the day's first
            health update (``epidemiology.py:127-135,295-303``).  The reference's ``VaccinationCampaigns`` object
            is compiled into the device tables on first use; ``person.vaccinated`` and the immunity dicts of the
            vaccinated people are written back."""
            ep = self.epidemiology
            camp = getattr(ep, "vaccination_campaigns", None) if ep is not None else None
            if camp is None:
                return
            today = self.timer.date.date()
            if getattr(self, "_b200_vac_day", None) == today:
                return
            self._b200_vac_day = today
            if not hasattr(self, "_b200_vac_first"):
                from .vaccination import campaigns_from_reference
                cls = self._selector.infection_class
                self._b200_vac_first = self.timer.initial_date.date()
                mine = campaigns_from_reference(camp, {cls.infection_id(): cls.__name__})
                self._eng.set_vaccination(*mine.device_tables(self._b200_vac_first, [cls.__name__]))
                self._b200_vac_names = [v.name for v in mine.vaccines()]
            self._eng.vaccinate((today - self._b200_vac_first).days, seed=self.b200_seed)
            vs = self._eng.fetch_vaccination()
            prev = getattr(self, "_b200_vac_dose", None)
            if prev is None:
                prev = np.full(len(self._people), -1, np.int8)
            changed = np.nonzero(vs["dose"] != prev)[0]  # the people who got a dose today
            self._b200_vac_dose = vs["dose"].copy()
            if not len(changed):
                return
            iid = self._selector.infection_class.infection_id()
            events = getattr(self.record, "events", None) if self.record is not None else None
            susc = self._eng.fetch_person_state()["susceptibility"][0] if len(changed) else None
            for i in changed:
                p = self._people[i]
                if events is not None:  # vaccines.py:309-313
                    events["vaccines"].accumulate(p.id, self._b200_vac_names[int(vs["vaccine"][i])], int(vs["dose"][i]))
                p.vaccinated = int(vs["dose"][i])
                p.vaccine_type = self._b200_vac_names[int(vs["vaccine"][i])]
                p.immunity.susceptibility_dict[iid] = float(susc[i])
                p.immunity.effective_multiplier_dict[iid] = float(vs["effective_multiplier"][0][i])

        # ---- device events -> the reference's Record (records_writer.py:203-246) ----------------------
        def _b200_feed_record(self):
            """What the reference's stages accumulate during a step, fed from the device's event lists with the
            reference's own keyword arguments: ``infections`` (``interaction.py:664-683``, one call per location),
            ``symptoms`` / ``deaths`` / ``recoveries`` (``epidemiology.py:176-186,209-214,256-266``),
            ``hospital_admissions`` / ``icu_admissions`` / ``discharges`` (``medical_care_policies.py:463-515``);
            then ``summarise_time_step`` / ``time_step`` as ``Epidemiology.do_timestep`` does (``:161-166``)."""
            soa, rec = self._soa, self.record
            if not hasattr(self, "_b200_grp_sg"):
                self._b200_grp_sg = soa.group_supergroup()
            grp_sg = self._b200_grp_sg
            iid = self._selector.infection_class.infection_id()
            pid = lambda i: self._people[int(i)].id
            gspec = lambda g: soa.supergroups[grp_sg[g]].spec
            gid = lambda g: int(soa.group_ids[g])
            ev = self._eng.fetch_infection_events()
            if len(ev["infected"]):
                order = np.argsort(ev["group"], kind="stable")
                groups = ev["group"][order]
                cuts = np.nonzero(np.diff(groups))[0] + 1
                for sel in np.split(order, cuts):
                    g = int(ev["group"][sel[0]])
                    rec.accumulate(table_name="infections", location_spec=gspec(g), location_id=gid(g),
                                   region_name=soa.region_names[int(soa.grp_info[g] >> 16)],
                                   infected_ids=[pid(i) for i in ev["infected"][sel]], infection_ids=[iid] * len(sel),
                                   infector_ids=[pid(i) for i in ev["infector"][sel]])
            hev = self._eng.fetch_health_events()
            for p, k, g in zip(hev["person"], hev["kind"], hev["group"]):
                k, g = int(k), int(g)
                if k == 1:
                    rec.accumulate(table_name="deaths", location_spec=gspec(g), location_id=gid(g), dead_person_id=pid(p))
                elif k == 2:
                    rec.accumulate(table_name="recoveries", recovered_person_id=pid(p), infection_id=iid)
                elif k == 3:
                    rec.accumulate(table_name="hospital_admissions", hospital_id=gid(g), patient_id=pid(p))
                elif k == 4:
                    rec.accumulate(table_name="icu_admissions", hospital_id=gid(g), patient_id=pid(p))
                elif k == 5:
                    rec.accumulate(table_name="discharges", hospital_id=gid(g), patient_id=pid(p))
                elif k == 6:
                    rec.accumulate(table_name="symptoms", infected_id=pid(p), symptoms=g, infection_id=iid)
            rec.summarise_time_step(timestamp=self.timer.date, world=self.world)
            rec.time_step(timestamp=self.timer.date)

        def clear_world(self):
            """The reference empties every subgroup list after each step (``simulator.py:305-343``).
            Here membership is a per-step byte on the device, so the lists are emptied once (world
            builders leave people in them) and never filled again."""
            if not getattr(self, "_b200_cleared", False):
                super().clear_world()
                self._b200_cleared = True

    return B200Simulator


class _InfectionView:
    """What the rest of JUNE reads from ``person.infection`` (``infection.py:13-135``)."""

    __slots__ = ("transmission", "symptoms", "start_time", "_cls")

    def __init__(self, infection_class, probability, tag):
        self._cls = infection_class
        self.transmission = _TransmissionView(probability)
        self.symptoms = _SymptomsView(tag, tag, 0, [])
        self.start_time = -1

    def infection_id(self):
        return self._cls.infection_id()

    def immunity_ids(self):
        return self._cls.immunity_ids()

    @property
    def tag(self):
        return self.symptoms.tag

    @property
    def max_tag(self):
        return self.symptoms.max_tag

    @property
    def infection_probability(self):
        return self.transmission.probability
