"""``DeviceWorld``: a world resident in HBM behind the C ABI of ``libjune_b200.so``.

``Engine`` is the ABI-generic part (fills ``JbWorldDesc`` from numpy arrays and wraps every
entry point); ``DeviceWorld`` binds it to the CUDA library.  The tests bind the same ``Engine``
to the CPU oracle library, which exports the identical ABI with the ``jo_`` prefix.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import (JbVaccine, JbVaccinationCampaign, JbCommuteDesc, JbDisease, JbDist, JbIndividualPolicy, JbInfectionRow, JbLeisureDesc, JbStepParams, JbStepStats,
                   JbSupergroup, JbWorldDesc, SimulatorError, ptr)
from .disease import DiseaseConfig, health_index_cum
from .interaction import Interaction
from .world import CARE_HOME_VISITORS_SUBGROUP, KIND_SCHOOL, MAX_STAGES, N_TPAR, SCHOOL_DIM, VISITS_ACTIVITY, VISITS_SUPERGROUP, WorldSoA


def fill_disease(dc: DiseaseConfig, hi_cum: np.ndarray) -> JbDisease:
    d = JbDisease()
    d.n_tags = dc.n_tags
    d.tag_recovered_mask = dc.tag_mask("recovered_stage")
    d.tag_dead_mask = dc.tag_mask("fatality_stage")
    d.tag_hospital_mask = dc.tag_mask("hospitalised_stage")
    d.tag_icu_mask = dc.tag_mask("intensive_care_stage")
    d.tag_severe_mask = dc.tag_mask("severe_symptoms_stay_at_home_stage")
    d.tag_exposed = dc.symptom_tags[dc.default_lowest_stage]
    table = dc.trajectory_table()
    assert len(table) <= 16
    d.n_traj = len(table)
    for k, (max_tag, stages) in enumerate(table):
        d.traj_max_tag[k] = max_tag
        d.traj_n_stages[k] = len(stages)
        for s, (tag, code, p) in enumerate(stages):
            d.traj_stage_tag[k][s] = tag
            d.traj_stage_time[k][s].type = code
            for q in range(4):
                d.traj_stage_time[k][s].p[q] = p[q]
    ttype, pars = dc.transmission_table()
    d.trans_type = ttype
    for k, (code, p) in enumerate(pars):
        d.trans_par[k].type = code
        for q in range(4):
            d.trans_par[k].p[q] = p[q]
    assert hi_cum.shape == (2, 2, 100, dc.n_tags) and hi_cum.dtype == np.float64
    d.health_index_cum = hi_cum.ctypes.data
    d.infection_id = dc.infection_id
    return d


class Engine:
    """One world behind one ABI library."""

    def __init__(self, lib, prefix: str, world: WorldSoA, interaction: Interaction, disease: DiseaseConfig,
                 outcome_prob: np.ndarray, device: int = 0, n_variants: int = 1, slot_capacity: int = 0):
        self._lib = lib
        self._p = prefix
        self.device_index = device
        self.world = world
        self.interaction = interaction
        self.disease = disease
        self.n_variants = n_variants
        world.validate()
        self._keep: List[object] = []
        # contact matrices, one block per supergroup spec
        cm_blocks, sgs = [], (JbSupergroup * len(world.supergroups))()
        off = 0
        for k, sg in enumerate(world.supergroups):
            m = np.ascontiguousarray(interaction.contact_matrices[sg.spec], dtype=np.float64)
            if sg.kind == KIND_SCHOOL:
                assert m.shape == (SCHOOL_DIM, SCHOOL_DIM)
            dim = m.shape[0]
            sgs[k].first_group, sgs[k].n_groups = sg.first_group, sg.n_groups
            sgs[k].kind, sgs[k].dim, sgs[k].cm_offset = sg.kind, dim, off
            sgs[k].activity, sgs[k].launch_mode, sgs[k].has_dynamic = sg.activity, sg.launch_mode, sg.has_dynamic
            cm_blocks.append(m.ravel())
            off += dim * dim
        cm = np.concatenate(cm_blocks) if cm_blocks else np.zeros(1)
        self._hi = np.ascontiguousarray(health_index_cum(outcome_prob))
        d = JbWorldDesc()
        d.abi_version = _lib.ABI_VERSION
        d.n_people, d.n_halo, d.n_groups, d.n_members = world.n_people, world.n_halo, world.n_groups, world.n_members
        d.n_supergroups, d.n_regions, d.n_super_areas = len(world.supergroups), world.n_regions, world.n_super_areas
        d.n_variants, d.n_beta, d.n_scale = n_variants, len(world.beta_rules), len(world.beta_scale)
        d.n_cm_doubles, d.n_school_years = cm.size, world.school_years.size
        d.slot_capacity, d.newinf_capacity = slot_capacity, 0
        d.person_id_base = world.person_id_base
        arrays = dict(
            age=world.age, sex=world.sex, care_home_resident=world.care_home_resident,
            has_activity=world.has_activity, super_area=world.super_area, sa_hospital=world.sa_hospital,
            sa_region=world.sa_region, grp_offset=world.grp_offset, grp_info=world.grp_info, grp_aux=world.grp_aux,
            reg_member=world.reg_member, reg_info=world.reg_info, school_years=world.school_years,
            cm=cm, beta_scale=world.beta_scale)
        for name, arr in arrays.items():
            arr = np.ascontiguousarray(arr)
            self._keep.append(arr)
            setattr(d, name, arr.ctypes.data)
        d.supergroups = C.addressof(sgs)
        d.disease = fill_disease(disease, self._hi)
        self._keep += [sgs, cm]
        handle = C.c_void_p()
        self._check(self._f("world_create")(C.byref(d), device, C.byref(handle)))
        self._h = handle
        if world.has_leisure:
            self._set_leisure(world)
        if world.has_commute:
            self._set_commute(world)
        if world.person_attr is not None:
            pa = np.ascontiguousarray(world.person_attr, dtype=np.uint8)
            self._check(self._f("set_person_attributes")(self._h, ptr(pa)))
        if getattr(world, "person_work_region", None) is not None:
            wr = np.ascontiguousarray(world.person_work_region, dtype=np.uint8)
            self._check(self._f("set_person_work_regions")(self._h, ptr(wr)))
        if world.n_halo:
            hg = np.ascontiguousarray(world.halo_gid, dtype=np.uint32)
            self._check(self._f("set_halo_gids")(self._h, ptr(hg), hg.size))
        if world.out_person.size:
            op = np.ascontiguousarray(world.out_person, dtype=np.uint32)
            self._check(self._f("set_halo")(self._h, ptr(op), op.size))

    # ---- leisure --------------------------------------------------------------------------------
    def _set_leisure(self, world: WorldSoA):
        names = [sg.name for sg in world.supergroups]
        K = len(world.leisure_activities)
        sgi = np.array([-1 if n == VISITS_SUPERGROUP else names.index(n) for n in world.leisure_supergroups], np.int32)
        lsg = np.array(world.leisure_subgroups, np.int32)
        pa = np.ascontiguousarray(world.person_area, dtype=np.uint32)
        off = np.ascontiguousarray(world.area_venue_off, dtype=np.uint32)
        ven = np.ascontiguousarray(world.area_venue, dtype=np.uint32)
        assert off.size == world.n_areas * K + 1 and int(off[-1]) == ven.size
        d = JbLeisureDesc()
        d.n_activities, d.n_areas = K, world.n_areas
        d.supergroup, d.subgroup = sgi.ctypes.data, lsg.ctypes.data
        d.person_area, d.area_off, d.area_venue = pa.ctypes.data, off.ctypes.data, (ven if ven.size else np.zeros(1, np.uint32)).ctypes.data
        self._check(self._f("set_leisure")(self._h, C.byref(d)))
        if world.has_visits:
            self._set_visits(world)

    def _set_visits(self, world: WorldSoA, probabilities=(0.66, 0.34)):
        """Residence visits (``residence_visits.py``): the households' links and ``residence_type_probabilities``
        (defaults of ``configs/defaults/groups/leisure/visits.yaml``; ``set_visit_probabilities`` changes them)."""
        names = [sg.name for sg in world.supergroups]
        spec_of = {sg.spec: i for i, sg in enumerate(world.supergroups)}
        d = _lib.JbVisitsDesc()
        d.activity = world.leisure_activities.index(VISITS_ACTIVITY)
        d.household_supergroup = spec_of["household"]
        has_ch = world.visit_care_home_off is not None and int(world.visit_care_home_off[-1]) > 0
        d.care_home_supergroup = spec_of["care_home"] if has_ch else -1
        d.care_home_subgroup = CARE_HOME_VISITORS_SUBGROUP
        arrs = [np.ascontiguousarray(world.visit_household_off, np.uint32), np.ascontiguousarray(world.visit_household, np.uint32)]
        d.household_off, d.household_target = arrs[0].ctypes.data, (arrs[1] if arrs[1].size else np.zeros(1, np.uint32)).ctypes.data
        if has_ch:
            arrs += [np.ascontiguousarray(world.visit_care_home_off, np.uint32), np.ascontiguousarray(world.visit_care_home, np.uint32)]
            d.care_home_off, d.care_home_target = arrs[2].ctypes.data, arrs[3].ctypes.data
        d.p_household, d.p_care_home = float(probabilities[0]), float(probabilities[1])
        keys = [r.key for r in world.beta_rules]
        assert "household_visits" in keys, "a world with residence visits needs the beta rule 'household_visits'"
        d.household_visits_beta = keys.index("household_visits")
        assert arrs[0].size == world.supergroups[d.household_supergroup].n_groups + 1
        self._check(self._f("set_visits")(self._h, C.byref(d)))

    def set_visit_probabilities(self, p_household: float, p_care_home: float):
        """``ChangeVisitsProbability`` (``leisure_policies.py:180-220``)."""
        self._check(self._f("set_visit_probabilities")(self._h, float(p_household), float(p_care_home)))

    def set_leisure_tables(self, does: np.ndarray, cum: np.ndarray):
        """``does[T][R][2][100]``, ``cum[T][R][2][100][K]`` (see ``june_b200.leisure``)."""
        does = np.ascontiguousarray(does, dtype=np.float64)
        cum = np.ascontiguousarray(cum, dtype=np.float64)
        K = len(self.world.leisure_activities)
        assert does.shape[1:] == (self.world.n_regions, 2, 100) and cum.shape == does.shape + (K,)
        self._check(self._f("set_leisure_tables")(self._h, does.shape[0], ptr(does), ptr(cum)))

    def inject_leisure(self, group_per_person: np.ndarray):
        g = np.ascontiguousarray(group_per_person, dtype=np.int32)
        assert g.shape == (self.world.n_people,)
        self._check(self._f("inject_leisure")(self._h, ptr(g)))

    # ---- vaccination -------------------------------------------------------------------------------
    def set_vaccination(self, vaccines: Sequence[dict], campaigns: Sequence[dict]):
        """Tables produced by ``VaccinationCampaigns.device_tables``."""
        va = (JbVaccine * max(1, len(vaccines)))()
        keep = []
        for k, v in enumerate(vaccines):
            va[k].n_doses, va[k].waning_factor = int(v["n_doses"]), float(v["waning_factor"])
            for q in range(v["n_doses"]):
                va[k].days_administered_to_effective[q] = int(v["days_administered_to_effective"][q])
                va[k].days_effective_to_waning[q] = int(v["days_effective_to_waning"][q])
                va[k].days_waning[q] = int(v["days_waning"][q])
            for name in ("sterilisation", "symptomatic"):
                arr = np.ascontiguousarray(v[name], dtype=np.float64)
                assert arr.shape == (v["n_doses"], self.n_variants, 100), arr.shape
                keep.append(arr)
                setattr(va[k], name, arr.ctypes.data)
        ca = (JbVaccinationCampaign * max(1, len(campaigns)))()
        for k, c in enumerate(campaigns):
            ca[k].vaccine, ca[k].group_by, ca[k].lo, ca[k].hi = int(c["vaccine"]), int(c["group_by"]), int(c["lo"]), int(c["hi"])
            ca[k].n_doses = len(c["dose_numbers"])
            for q, (dn, dd) in enumerate(zip(c["dose_numbers"], c["days_to_next_dose"])):
                ca[k].dose_numbers[q], ca[k].days_to_next_dose[q] = int(dn), int(dd)
            ca[k].start_day, ca[k].end_day = int(c["start_day"]), int(c["end_day"])
            ca[k].last_dose_vaccines, ca[k].coverage = int(c.get("last_dose_vaccines", 0)), float(c["coverage"])
        self._check(self._f("set_vaccination")(self._h, va, len(vaccines), ca, len(campaigns)))

    def vaccinate(self, today: int, seed: int = 0):
        """One simulated day of ``VaccinationCampaigns.apply`` + ``update_vaccine_effect`` for everybody."""
        self._check(self._f("vaccinate")(self._h, int(today), int(seed)))

    def inject_vaccination_uniforms(self, u: Optional[np.ndarray]):
        if u is None:
            self._check(self._f("inject_vaccination_uniforms")(self._h, None))
            return
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.shape == (self.world.n_people,)
        self._check(self._f("inject_vaccination_uniforms")(self._h, ptr(u)))

    def fetch_vaccination(self) -> Dict[str, np.ndarray]:
        n = self.world.n_people
        out = dict(dose=np.zeros(n, np.int8), vaccine=np.zeros(n, np.int8), effective_multiplier=np.ones((self.n_variants, n)))
        self._check(self._f("fetch_vaccination")(self._h, ptr(out["dose"]), ptr(out["vaccine"]), ptr(out["effective_multiplier"])))
        return out

    _VACC_STATE = (("dose", np.int8, False), ("vaccine", np.int8, False), ("campaign", np.int8, False),
                   ("first_day", np.int32, False), ("stage", np.uint8, False), ("prior_infection", np.float64, True),
                   ("prior_symptoms", np.float64, True), ("effective_multiplier", np.float64, True))

    def fetch_vaccination_state(self) -> Dict[str, np.ndarray]:
        """Everything the device keeps per person about vaccination (checkpoints)."""
        n = self.world.n_people
        out = {k: np.zeros((self.n_variants, n) if per_variant else n, dt) for k, dt, per_variant in self._VACC_STATE}
        st = _lib.JbVaccinationState(*(out[k].ctypes.data for k, _, _ in self._VACC_STATE))
        self._check(self._f("vaccination_state")(self._h, 0, C.byref(st)))
        return out

    def set_vaccination_state(self, state: Dict[str, np.ndarray]) -> None:
        """Write a saved vaccination state back (after ``set_vaccination``)."""
        n = self.world.n_people
        arrs = {k: np.ascontiguousarray(state[k], dt).reshape((self.n_variants, n) if per_variant else n)
                for k, dt, per_variant in self._VACC_STATE}
        st = _lib.JbVaccinationState(*(arrs[k].ctypes.data for k, _, _ in self._VACC_STATE))
        self._check(self._f("vaccination_state")(self._h, 1, C.byref(st)))

    # ---- commute --------------------------------------------------------------------------------
    def _set_commute(self, world: WorldSoA):
        arrs = [np.ascontiguousarray(world.person_commute, dtype=np.int32)] + [
            np.ascontiguousarray(x, dtype=np.uint32) for x in (world.city_station_off, world.city_station,
                                                               world.station_transport_off, world.station_transport)]
        arrs = [a if a.size else np.zeros(1, a.dtype) for a in arrs]
        d = JbCommuteDesc()
        d.n_cities, d.n_stations = len(world.city_station_off) - 1, len(world.station_transport_off) - 1
        (d.person_commute, d.city_station_off, d.city_station, d.station_transport_off,
         d.station_transport) = (a.ctypes.data for a in arrs)
        self._check(self._f("set_commute")(self._h, C.byref(d)))

    def inject_commute(self, group_per_person: np.ndarray):
        g = np.ascontiguousarray(group_per_person, dtype=np.int32)
        assert g.shape == (self.world.n_people,)
        self._check(self._f("inject_commute")(self._h, ptr(g)))

    # ---- individual policies ---------------------------------------------------------------------
    def set_individual_policies(self, records: Sequence[dict], regional_compliance: Optional[Sequence[float]] = None):
        """``records``: dicts with ``type`` (``_lib.POL_*``), optional ``flags``, ``mask`` and up to six
        doubles ``p`` (None -> NaN), in the reference's policy-list order."""
        n = len(records)
        arr = (JbIndividualPolicy * max(1, n))()
        for k, r in enumerate(records):
            arr[k].type, arr[k].flags, arr[k].mask = int(r["type"]), int(r.get("flags", 0)), int(r.get("mask", 0))
            for q, v in enumerate(list(r.get("p", []))[:6]):
                arr[k].p[q] = float("nan") if v is None else float(v)
        rc = np.ascontiguousarray(regional_compliance if regional_compliance is not None else np.ones(self.world.n_regions),
                                  dtype=np.float64)
        assert rc.shape == (self.world.n_regions,)
        self._check(self._f("set_individual_policies")(self._h, arr, n, ptr(rc)))

    def inject_policy_uniforms(self, u: np.ndarray):
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.shape == (_lib.MAX_POLICIES, _lib.POLICY_DRAWS, self.world.n_people)
        self._check(self._f("inject_policy_uniforms")(self._h, ptr(u)))

    # ---- plumbing ------------------------------------------------------------------------------
    def _f(self, name):
        return getattr(self._lib, self._p + name)

    def _check(self, rc: int):
        if rc == 0:
            return
        msg = self._f("last_error")().decode()
        if rc in (_lib.JB_ENOACT, _lib.JB_ECAP):
            raise SimulatorError(msg)
        if rc == _lib.JB_ENOGPU:
            raise RuntimeError(msg)
        raise RuntimeError(f"{self._p}* call failed (code {rc}): {msg}")

    def close(self):
        if getattr(self, "_h", None):
            self._f("world_destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # ---- state -----------------------------------------------------------------------------------
    def set_person_state(self, pstate: Optional[np.ndarray] = None, susceptibility: Optional[np.ndarray] = None):
        if pstate is not None:
            pstate = np.ascontiguousarray(pstate, dtype=np.uint8)
            assert pstate.shape == (self.world.n_people,)
        if susceptibility is not None:
            susceptibility = np.ascontiguousarray(susceptibility, dtype=np.float64).reshape(self.n_variants, self.world.n_people)
        self._check(self._f("set_person_state")(self._h, ptr(pstate), ptr(susceptibility)))

    def set_halo_activities(self, activity_mask: int):
        """Activities under which halo persons are registered in any domain (same value on every rank)."""
        self._check(self._f("set_halo_activities")(self._h, int(activity_mask)))

    def set_variants(self, trans_par: Optional[Sequence[Sequence[tuple]]] = None, cross_immunity: Optional[Sequence[int]] = None):
        """Per-variant transmission distributions (``[(type code, p[4])] * 6`` per variant, as
        ``DiseaseConfig.transmission_table`` returns them) and cross-immunity bit masks."""
        from ._lib import JbDist
        V = self.n_variants
        tp = None
        if trans_par is not None:
            assert len(trans_par) == V
            tp = (JbDist * (6 * V))()
            for v, row in enumerate(trans_par):
                for k, (code, p) in enumerate(row):
                    tp[v * 6 + k].type = int(code)
                    for q in range(4):
                        tp[v * 6 + k].p[q] = float(p[q])
        ci = None if cross_immunity is None else np.ascontiguousarray(cross_immunity, dtype=np.uint32)
        assert ci is None or ci.shape == (V,)
        self._check(self._f("set_variants")(self._h, V, tp, ptr(ci)))

    def mutate(self, variant: int, regional_probability: Sequence[float], seed: int, step: int):
        """``Mutation.apply`` (``event/mutation.py:53-66``)."""
        pr = np.ascontiguousarray(regional_probability, dtype=np.float64)
        assert pr.shape == (self.world.n_regions,)
        self._check(self._f("mutate")(self._h, int(variant), ptr(pr), int(seed), int(step)))

    def set_medical(self, medical: np.ndarray):
        """Restore ``person.subgroups.medical_facility`` (``hospital group << 8 | ward``, -1 = none); call
        before ``upload_infections``."""
        medical = np.ascontiguousarray(medical, dtype=np.int32)
        assert medical.shape == (self.world.n_people,)
        self._check(self._f("set_medical")(self._h, ptr(medical)))

    def inject_uniforms(self, u: np.ndarray):
        u = np.ascontiguousarray(u, dtype=np.float64)
        self._check(self._f("inject_uniforms")(self._h, ptr(u), u.size))

    def infect(self, persons: Sequence[int], time: float, seed: int, step: int, variants: Optional[Sequence[int]] = None):
        p = np.ascontiguousarray(persons, dtype=np.uint32)
        v = None if variants is None else np.ascontiguousarray(variants, dtype=np.uint32)
        self._check(self._f("infect")(self._h, ptr(p), ptr(v), p.size, float(time), int(seed), int(step)))

    def upload_infections(self, rows: List[dict]):
        arr = (JbInfectionRow * len(rows))()
        for i, r in enumerate(rows):
            a = arr[i]
            a.person, a.variant, a.start_time = int(r["person"]), int(r.get("variant", 0)), float(r["start_time"])
            for q in range(N_TPAR):
                a.tpar[q] = float(r["tpar"][q]) if q < len(r["tpar"]) else 0.0
            a.probability = float(r.get("probability", 0.0))
            tt, tg = list(r["traj_time"]), list(r["traj_tag"])
            a.n_stages, a.stage = len(tt), int(r.get("stage", 0))
            a.max_tag, a.tag = int(r.get("max_tag", max(tg))), int(r.get("tag", tg[a.stage]))
            for q in range(len(tt)):
                a.traj_time[q], a.traj_tag[q] = float(tt[q]), int(tg[q])
        self._check(self._f("upload_infections")(self._h, arr, len(rows)))

    # ---- stepping --------------------------------------------------------------------------------
    def make_params(self, now: float, delta_time: float, step: int, activity_mask: int, seed: int, flags: int,
                    beta: np.ndarray, leisure_table: int = 0) -> JbStepParams:
        cache = self.__dict__.setdefault("_beta_cache", {})
        hit = cache.get(id(beta))
        if hit is None or hit[0] is not beta:  # callers pass the same table object while the policies do not change
            b = np.ascontiguousarray(beta, dtype=np.float64)
            assert b.shape == (len(self.world.beta_rules), self.world.n_regions), b.shape
            if len(cache) > 64:
                cache.clear()
            hit = cache[id(beta)] = (beta, b, b.ctypes.data)
        p = JbStepParams()
        p.now, p.delta_time, p.step, p.activity_mask = now, delta_time, step, activity_mask
        p.seed, p.flags, p.leisure_table = seed, flags, leisure_table
        p.beta = hit[2]
        p._beta_keep = hit[1]
        return p

    def step(self, params: JbStepParams, want_stats: bool = True) -> Optional[Dict[str, int]]:
        stats = JbStepStats()
        self._check(self._f("step")(self._h, C.byref(params), C.byref(stats) if want_stats else None))
        return stats.as_dict() if want_stats else None

    def describe(self) -> str:
        """Text summary of the device layout (kernel flavour per supergroup)."""
        buf = C.create_string_buffer(8192)
        self._check(self._f("world_describe")(self._h, buf, len(buf)))
        return buf.value.decode()

    def step_prepare(self, params: JbStepParams) -> None:
        """Build (capture + instantiate) the graph of a step with these activities / flags without running it."""
        self._check(self._f("step_prepare")(self._h, C.byref(params)))

    def step_begin(self, params: JbStepParams, d_send: int = 0):
        self._check(self._f("step_begin")(self._h, C.byref(params), C.c_void_p(d_send or None)))

    def step_interact(self, d_recv: int = 0, d_ret_send: int = 0):
        self._check(self._f("step_interact")(self._h, C.c_void_p(d_recv or None), C.c_void_p(d_ret_send or None)))

    def step_finish(self, d_ret_recv: int = 0, want_stats: bool = True) -> Optional[Dict[str, int]]:
        stats = JbStepStats()
        self._check(self._f("step_finish")(self._h, C.c_void_p(d_ret_recv or None), C.byref(stats) if want_stats else None))
        return stats.as_dict() if want_stats else None

    def sync(self):
        self._check(self._f("sync")(self._h))

    def halo_record_size(self) -> int:
        return int(self._f("halo_record_size_w")(self._h))

    # ---- readback --------------------------------------------------------------------------------
    def fetch_new_infections(self) -> Tuple[np.ndarray, np.ndarray]:
        cap = self.world.n_people + self.world.n_halo
        m, v = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
        n = C.c_size_t()
        self._check(self._f("fetch_new_infections")(self._h, ptr(m), ptr(v), cap, C.byref(n)))
        return m[: n.value].copy(), v[: n.value].copy()

    def fetch_infection_events(self) -> Dict[str, np.ndarray]:
        """Rows of the reference's ``infections`` record table for the last step that ran with
        ``STEP_RECORD_EVENTS``: ``infected``, ``infector`` (person indices), ``group``, ``variant``."""
        cap = self.world.n_people + self.world.n_halo
        a = [np.zeros(cap, np.uint32) for _ in range(4)]
        n = C.c_size_t()
        self._check(self._f("fetch_infection_events")(self._h, ptr(a[0]), ptr(a[1]), ptr(a[2]), ptr(a[3]), cap, C.byref(n)))
        return dict(infected=a[0][: n.value].copy(), infector=a[1][: n.value].copy(), group=a[2][: n.value].copy(),
                    variant=a[3][: n.value].copy())

    HEALTH_EVENT_KINDS = {1: "deaths", 2: "recoveries", 3: "hospital_admissions", 4: "icu_admissions", 5: "discharges", 6: "symptoms"}

    def fetch_health_events(self) -> Dict[str, np.ndarray]:
        """Rows of the reference's ``deaths`` / ``recoveries`` / ``hospital_admissions`` /
        ``icu_admissions`` / ``discharges`` record tables for the last step that ran with
        ``STEP_RECORD_EVENTS``: ``person``, ``kind`` (``HEALTH_EVENT_KINDS``), ``group`` (-1: no location)."""
        cap = 4 * self.world.n_people
        a = [np.zeros(cap, np.uint32), np.zeros(cap, np.uint32), np.zeros(cap, np.int32)]
        n = C.c_size_t()
        self._check(self._f("fetch_health_events")(self._h, ptr(a[0]), ptr(a[1]), ptr(a[2]), cap, C.byref(n)))
        return dict(person=a[0][: n.value].copy(), kind=a[1][: n.value].copy(), group=a[2][: n.value].copy())

    def fetch_person_state(self) -> Dict[str, np.ndarray]:
        n = self.world.n_people
        out = dict(pstate=np.zeros(n, np.uint8), tag=np.zeros(n, np.int8), trans_prob=np.zeros(n, np.float64),
                   susceptibility=np.zeros((self.n_variants, n), np.float64), medical=np.zeros(n, np.int32))
        self._check(self._f("fetch_person_state")(self._h, ptr(out["pstate"]), ptr(out["tag"]), ptr(out["trans_prob"]),
                                                  ptr(out["susceptibility"]), ptr(out["medical"])))
        return out

    def fetch_lambda(self) -> np.ndarray:
        out = np.zeros(self.world.n_people + self.world.n_halo, np.float64)
        self._check(self._f("fetch_lambda")(self._h, ptr(out)))
        return out

    def fetch_infections(self) -> List[dict]:
        n = C.c_size_t()
        self._check(self._f("fetch_infections")(self._h, None, 0, C.byref(n)))
        arr = (JbInfectionRow * max(1, n.value))()
        self._check(self._f("fetch_infections")(self._h, arr, n.value, C.byref(n)))
        out = []
        for i in range(n.value):
            a = arr[i]
            ns = a.n_stages
            out.append(dict(person=a.person, variant=a.variant, start_time=a.start_time, tpar=list(a.tpar),
                            probability=a.probability, stage=a.stage, max_tag=a.max_tag, tag=a.tag,
                            traj_time=list(a.traj_time)[:ns], traj_tag=list(a.traj_tag)[:ns]))
        return out

    # include/june_b200.h, JB_N_SUMMARY ("daily_infected" here = infections created in this domain by the PERSON's
    # region; Simulator.summary_rows replaces it by the reference's count by the region of the infecting group)
    SUMMARY_COLUMNS = ("daily_infected", "current_infected", "current_hospitalised", "current_intensive_care",
                       "daily_deaths", "daily_hospital_deaths", "daily_recovered", "daily_hospitalised",
                       "daily_intensive_care")

    def fetch_summary(self) -> np.ndarray:
        """Per-region counts of the last step, columns ``SUMMARY_COLUMNS`` (the quantities of the
        reference's ``summary.csv``, ``records_writer.py:151-164``)."""
        out = np.zeros((self.world.n_regions, len(self.SUMMARY_COLUMNS)), np.uint32)
        self._check(self._f("fetch_summary")(self._h, ptr(out)))
        return out

    def flush_l2(self, n_bytes: int = 256 << 20):
        """Evict the L2 on the library's stream, in front of the next step's event bracket."""
        self._check(self._f("flush_l2")(self._h, int(n_bytes)))

    def set_concurrency(self, on: bool = True):
        self._check(self._f("set_concurrency")(self._h, 1 if on else 0))

    def timing_enable(self, level: int = 1):
        """0 off; 1 whole-step CUDA events (graph replay stays on); 2 + per-kernel events (plain launches)."""
        self._check(self._f("timing_enable")(self._h, int(level)))

    def timing_read_supergroups(self) -> List[Dict[str, float]]:
        k = len(self.world.supergroups)
        ms, ln = (C.c_double * k)(), (C.c_uint64 * k)()
        self._check(self._f("timing_read_supergroups")(self._h, ms, ln, k))
        return [dict(name=sg.name, ms=ms[i], launches=int(ln[i])) for i, sg in enumerate(self.world.supergroups)]

    def timing_read(self, reset: bool = False) -> Dict[str, float]:
        a, b = C.c_double(), C.c_double()
        n, l = C.c_uint64(), C.c_uint64()
        self._check(self._f("timing_read")(self._h, C.byref(a), C.byref(b), C.byref(n), C.byref(l), 1 if reset else 0))
        return dict(interaction_ms=a.value, step_ms=b.value, n_steps=n.value, n_launches=l.value)


class DeviceWorld(Engine):
    """World resident on one CUDA device (the product path; raises without the CUDA library/GPU)."""

    def __init__(self, world: WorldSoA, interaction: Interaction, disease: DiseaseConfig, outcome_prob: np.ndarray,
                 device: int = 0, n_variants: int = 1, slot_capacity: int = 0):
        super().__init__(_lib.load(), "jb_", world, interaction, disease, outcome_prob, device=device,
                         n_variants=n_variants, slot_capacity=slot_capacity)
