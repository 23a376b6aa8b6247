"""World compiler: walk a reference ``World`` / ``Domain`` object graph ONCE and emit the
struct-of-arrays form the device consumes (``june_b200.world.WorldSoA``).

Duck-typed on purpose: it needs no ``import june``.  It reads exactly the attributes the
reference hot path reads:

* ``world.people`` (iterable of persons with ``id, age, sex, area, subgroups, dead``;
  ``june/demography/person.py:35-81``),
* ``getattr(world, supergroup_name).members`` (``june/simulator.py:404``), each group having
  ``spec, id, subgroups, super_area`` (+ ``sector`` for companies/schools, ``years`` for schools;
  ``june/groups/company.py:309-313``, ``june/groups/school.py:418-522``),
* ``world.super_areas.members`` with ``region`` and ``closest_hospitals``
  (``june/policy/medical_care_policies.py:454``).

See INTEGRATION.md for how a reference ``Simulator`` subclass uses it.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from .world import (ACT_MEDICAL, ACT_PRIMARY, ACT_RESIDENCE, ACTIVITY_CODES, KIND_MATRIX, KIND_SCHOOL, LOCKDOWN_CODES,
                    PRIMARY_KIND_CODES, VISITS_ACTIVITY, VISITS_SUPERGROUP, BetaRule, Supergroup, WorldSoA, make_ginfo, make_info,
                    make_person_attr)

# activity hierarchy, june/activity/activity_manager.py:24-33 (activities without device support
# in this round are listed so that the supergroup order stays the reference's)
ACTIVITY_HIERARCHY = ["medical_facility", "commute", "primary_activity", "leisure", "residence"]

_SINGULAR_SPEC = {
    "households": "household", "companies": "company", "schools": "school", "hospitals": "hospital",
    "care_homes": "care_home", "universities": "university", "pubs": "pub", "cinemas": "cinema",
    "groceries": "grocery", "gyms": "gym", "city_transports": "city_transport",
    "inter_city_transports": "inter_city_transport",
}


def supergroup_order(activity_to_super_groups: Dict[str, Sequence[str]]) -> List[Tuple[str, str]]:
    """(supergroup name, activity) in the order ``Simulator.do_timestep`` visits them when all
    activities are active (``activity_manager.py:172-188`` + ``simulator.py:399-409``)."""
    out = []
    seen = set()
    for act in ACTIVITY_HIERARCHY:
        for name in activity_to_super_groups.get(act, []):
            if "visits" in name or name in seen:
                continue
            seen.add(name)
            out.append((name, act))
    return out


def school_beta_key(sector) -> str:
    # school.py:493-498
    if sector is None:
        return "school"
    if "secondary" in sector:
        return "secondary_school"
    return "primary_school"


# specs whose groups go through the warp-per-group kernel: per-school matrices, dynamic members
# (hospital wards), contact matrices larger than 4x4.  Everything else is tiled.
WARP_MODE_SPECS = ("school", "hospital", "university")


LEISURE_BIN_CAP = 2048  # visitors per social venue and step


def _haversine(origin, destination) -> float:
    """Great-circle distance in km between (lat, lon) pairs in degrees (``june/utils/distances.py``)."""
    lat1, lon1, lat2, lon2 = map(np.radians, (origin[0], origin[1], destination[0], destination[1]))
    a = np.sin((lat2 - lat1) / 2) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin((lon2 - lon1) / 2) ** 2
    return float(6371 * 2 * np.arcsin(np.sqrt(a)))


def compile_world(world, activity_to_super_groups: Dict[str, Sequence[str]], sector_betas: Dict[str, float],
                  warp_mode_specs=WARP_MODE_SPECS, leisure=None, travel=None, long_commute_distance=None) -> WorldSoA:
    """``leisure``: the reference's ``Leisure`` object (or None).  Its social-venue distributors
    (``leisure.leisure_distributors``, in order) become the leisure activities of the device world;
    ``area.social_venues[spec]`` (``social_venue_distributor.py:258``) the per-area candidates."""
    people = list(world.people)
    n = len(people)
    pidx = {id(p): i for i, p in enumerate(people)}

    super_areas = list(world.super_areas.members) if hasattr(world.super_areas, "members") else list(world.super_areas)
    sa_idx = {id(sa): i for i, sa in enumerate(super_areas)}
    region_names: List[str] = []
    sa_region = np.zeros(len(super_areas), np.uint16)
    for i, sa in enumerate(super_areas):
        name = getattr(getattr(sa, "region", None), "name", None) or "region"
        if name not in region_names:
            region_names.append(name)
        sa_region[i] = region_names.index(name)

    # ---- groups, supergroup after supergroup ---------------------------------------------
    beta_rules: List[BetaRule] = []

    def beta_index(rule: BetaRule) -> int:
        for i, r in enumerate(beta_rules):
            if r.key == rule.key and r.fallback == rule.fallback and r.uses_tier == rule.uses_tier:
                return i
        beta_rules.append(rule)
        return len(beta_rules) - 1

    scales: List[float] = [1.0]

    def scale_index(x: float) -> int:
        for i, s in enumerate(scales):
            if s == x:
                return i
        scales.append(float(x))
        return len(scales) - 1

    supergroups: List[Supergroup] = []
    groups = []          # reference group objects in device order
    gidx: Dict[int, int] = {}
    beta_i: List[int] = []
    scale_i: List[int] = []
    region_i: List[int] = []
    gsa_i: List[int] = []
    aux: List[int] = []
    years_flat: List[int] = []
    for name, act in supergroup_order(activity_to_super_groups):
        sg_obj = getattr(world, name, None)
        if sg_obj is None:
            continue
        members = [g for g in sg_obj.members if not getattr(g, "external", False)]
        if not members:
            continue
        spec = members[0].spec
        first = len(groups)
        kind = KIND_SCHOOL if spec == "school" else KIND_MATRIX
        if spec == "household":
            # household / household_visits / care_visits must be consecutive rows (household.py:280-291)
            hb = beta_index(BetaRule("household", None, False))
            beta_index(BetaRule("household_visits", None, False))
            beta_index(BetaRule("care_visits", None, False))
        for g in members:
            gidx[id(g)] = len(groups)
            groups.append(g)
            sa = getattr(g, "super_area", None)
            region_i.append(int(sa_region[sa_idx[id(sa)]]) if sa is not None and id(sa) in sa_idx else 0)
            gsa_i.append(sa_idx[id(sa)] if sa is not None and id(sa) in sa_idx else 0)
            sc = 0
            a = 0
            if spec == "household":
                b = hb
            elif spec == "school":
                key = school_beta_key(getattr(g, "sector", None))
                b = beta_index(BetaRule(key, "school" if key != "school" else None, True))
                yrs = list(g.years)
                assert len(yrs) == len(g.subgroups) - 1, "school years must match its classes"
                assert len(yrs) < 256
                a = (len(years_flat) << 8) | len(yrs)
                years_flat.extend(int(y) for y in yrs)
            elif spec == "company":
                b = beta_index(BetaRule("company", None, True))
                sc = scale_index(float(sector_betas.get(getattr(g, "sector", None), 1.0)))
            else:
                b = beta_index(BetaRule(spec, None, True))
            beta_i.append(b)
            scale_i.append(sc)
            aux.append(a)
        is_venue = act in ("leisure", "commute")
        # care homes take their visitors as dynamic members when the reference's Leisure has residence visits
        visited_ch = spec == "care_home" and leisure is not None and VISITS_ACTIVITY in getattr(leisure, "leisure_distributors", {})
        has_dyn = 1 if spec == "hospital" else LEISURE_BIN_CAP if is_venue else 64 if visited_ch else 0
        supergroups.append(Supergroup(name=name, spec=spec, activity=ACTIVITY_CODES[act], first_group=first,
                                      n_groups=len(members), kind=kind,
                                      launch_mode=1 if (spec in warp_mode_specs or is_venue or visited_ch) else 0, has_dynamic=has_dyn))

    # ---- registered members -----------------------------------------------------------------
    lsg_of: Dict[int, Tuple[int, int]] = {}
    for gi, g in enumerate(groups):
        for li, sub in enumerate(g.subgroups):
            lsg_of[id(sub)] = (gi, li)
    rows = []
    has_activity = np.zeros(n, np.uint8)
    ch = np.zeros(n, np.uint8)
    age = np.zeros(n, np.uint8)
    sex = np.zeros(n, np.uint8)
    psa = np.zeros(n, np.uint32)
    person_ids = np.zeros(n, np.int64)
    for i, p in enumerate(people):
        age[i] = min(int(p.age), 255)
        sex[i] = 1 if p.sex == "m" else 0
        person_ids[i] = int(p.id)
        area = getattr(p, "area", None)
        if area is not None and id(area.super_area) in sa_idx:
            psa[i] = sa_idx[id(area.super_area)]
        for act_name, act in (("residence", ACT_RESIDENCE), ("primary_activity", ACT_PRIMARY)):
            sub = getattr(p.subgroups, act_name, None) if p.subgroups is not None else None
            if sub is None:
                continue
            has_activity[i] |= 1 << act
            if getattr(sub, "external", False):
                raise NotImplementedError("external subgroups: compile the whole world, then partition (june_b200.domains)")
            key = lsg_of.get(id(sub))
            if key is None:
                continue  # subgroup of a supergroup that is not simulated
            rows.append((key[0], key[1], i, act))
            if act == ACT_RESIDENCE and sub.group.spec == "care_home":
                ch[i] = 1
    rows.sort()
    g_arr = np.array([r[0] for r in rows], np.int64)
    reg_member = np.array([r[2] for r in rows], np.uint32)
    reg_info = make_info(np.array([r[1] for r in rows], np.uint16), np.array([r[3] for r in rows], np.uint16)).astype(np.uint16)
    counts = np.bincount(g_arr, minlength=len(groups)) if len(rows) else np.zeros(len(groups), np.int64)
    grp_offset = np.zeros(len(groups) + 1, np.uint32)
    grp_offset[1:] = np.cumsum(counts)

    sa_hospital = np.full(len(super_areas), -1, np.int32)
    for i, sa in enumerate(super_areas):
        hs = getattr(sa, "closest_hospitals", None)
        if hs:
            sa_hospital[i] = gidx.get(id(hs[0]), -1)

    w = WorldSoA(
        age=age, sex=sex, care_home_resident=ch, has_activity=has_activity, super_area=psa,
        sa_hospital=sa_hospital, sa_region=sa_region, region_names=region_names,
        grp_offset=grp_offset, grp_info=make_ginfo(np.array(beta_i, np.uint32), np.array(scale_i, np.uint32),
                                                   np.array(region_i, np.uint32)).astype(np.uint32),
        grp_aux=np.array(aux, np.uint32), reg_member=reg_member, reg_info=reg_info,
        school_years=np.array(years_flat, np.uint8), supergroups=supergroups, beta_rules=beta_rules,
        beta_scale=np.array(scales, np.float64), person_ids=person_ids,
        group_ids=np.array([int(g.id) for g in groups], np.int64),
        grp_super_area=np.array(gsa_i, np.uint32),
    )
    # what the individual policies read per person (individual_policies.py:497-510,651-770)
    lock = np.array([LOCKDOWN_CODES.get(getattr(p, "lockdown_status", None), 0) for p in people], np.uint8)
    kind = np.zeros(n, np.uint8)
    kw2 = np.zeros(n, bool)
    for i, p in enumerate(people):
        sub = getattr(p.subgroups, "primary_activity", None) if p.subgroups is not None else None
        if sub is not None:
            kind[i] = PRIMARY_KIND_CODES.get(sub.group.spec, 4)
        res = getattr(p.subgroups, "residence", None) if p.subgroups is not None else None
        if res is not None and hasattr(res.group, "residents"):
            kw2[i] = sum(1 for m in res.group.residents if getattr(m, "lockdown_status", None) == "key_worker") > 1
    # LimitLongCommute._does_long_commute (individual_policies.py:803-813): haversine distance between
    # the home area and the work super area above the class-level threshold
    long_c = np.zeros(n, bool)
    if long_commute_distance is not None:
        for i, p in enumerate(people):
            wsa = getattr(p, "work_super_area", None)
            if wsa is not None and getattr(p, "area", None) is not None:
                long_c[i] = _haversine(p.area.coordinates, wsa.coordinates) > long_commute_distance
    w.person_attr = make_person_attr(lock, kind, kw2, long_c)
    # CloseCompaniesLockdownTiers (individual_policies.py:570-599): the work super area's region, when the
    # person works outside the home super area
    work_region = np.full(n, 0xFF, np.uint8)
    for i, p in enumerate(people):
        wsa = getattr(p, "work_super_area", None)
        home = getattr(getattr(p, "area", None), "super_area", None)
        reg = getattr(getattr(wsa, "region", None), "name", None)
        if wsa is not None and wsa is not home and reg in w.region_names:
            work_region[i] = w.region_names.index(reg)
    w.person_work_region = work_region
    if travel is not None and getattr(world, "cities", None) is not None:
        # Travel.get_commute_subgroup (travel.py:125-131): work city, public transport, internal commuter
        # of the city or commuter of the home super area's inter-city station (city.py:83-102)
        cities = list(world.cities.members) if hasattr(world.cities, "members") else list(world.cities)
        stations, st_idx = [], {}
        cs_off, cs = [0], []
        for c in cities:
            for st in list(c.city_stations) + list(c.inter_city_stations):
                st_idx[id(st)] = len(stations)
                stations.append(st)
            cs += [st_idx[id(st)] for st in c.city_stations]
            cs_off.append(len(cs))
        st_off, st_tr = [0], []
        for st in stations:
            trs = getattr(st, "city_transports", None) or getattr(st, "inter_city_transports", None) or []
            st_tr += [gidx[id(t)] for t in trs if id(t) in gidx]
            st_off.append(len(st_tr))
        cidx = {id(c): i for i, c in enumerate(cities)}
        pc = np.full(n, -1, np.int32)
        for i, p in enumerate(people):
            city = p.work_city
            mode = getattr(p, "mode_of_transport", None)
            if city is None or mode is None or not mode.is_public or not city.has_stations:
                continue
            if p.id in city.internal_commuter_ids:
                pc[i] = cidx[id(city)] << 1
            else:
                st = p.super_area.closest_inter_city_station_for_city.get(city.name)
                if st is not None and p.id in st.commuter_ids:
                    pc[i] = (st_idx[id(st)] << 1) | 1
        w.person_commute = pc
        w.city_station_off, w.city_station = np.array(cs_off, np.uint32), np.array(cs, np.uint32)
        w.station_transport_off, w.station_transport = np.array(st_off, np.uint32), np.array(st_tr, np.uint32)
    if leisure is not None:
        areas = list(world.areas.members) if hasattr(world.areas, "members") else list(world.areas)
        aidx = {id(a): i for i, a in enumerate(areas)}
        specs = [s_ for s_ in leisure.leisure_distributors if "visits" not in s_]
        by_spec = {sg.spec: sg for sg in supergroups}
        specs = [s_ for s_ in specs if s_ in by_spec]
        off, ven = [0], []
        for a in areas:
            for s_ in specs:
                for v in (getattr(a, "social_venues", {}) or {}).get(s_, ()) or ():
                    if id(v) in gidx:
                        ven.append(gidx[id(v)])
                off.append(len(ven))
        w.person_area = np.array([aidx[id(p.area)] for p in people], np.uint32)
        w.leisure_activities = specs
        w.leisure_supergroups = [by_spec[s_].name for s_ in specs]
        w.leisure_subgroups = [int(leisure.leisure_distributors[s_].leisure_subgroup_type) for s_ in specs]
        hh_sg = by_spec.get("household")
        if VISITS_ACTIVITY in leisure.leisure_distributors and hh_sg is not None:
            # residence visits (residence_visits.py): the activity keeps its place in the distributor order (an empty
            # column of the venue table); Household.residences_to_visit becomes two CSRs over the households
            k_vis = [s_ for s_ in leisure.leisure_distributors if s_ in by_spec or s_ == VISITS_ACTIVITY].index(VISITS_ACTIVITY)
            K = len(specs)
            off2, pos = [0], 0
            for a_i in range(len(areas)):
                for k in range(K + 1):
                    if k != k_vis:
                        kk = k if k < k_vis else k - 1
                        pos += off[a_i * K + kk + 1] - off[a_i * K + kk]
                    off2.append(pos)
            off = off2
            specs.insert(k_vis, VISITS_ACTIVITY)
            w.leisure_supergroups.insert(k_vis, VISITS_SUPERGROUP)
            w.leisure_subgroups.insert(k_vis, 0)
            vh_off, vh, vc_off, vc = [0], [], [0], []
            for g in groups[hh_sg.first_group: hh_sg.first_group + hh_sg.n_groups]:
                links = getattr(g, "residences_to_visit", None) or {}
                vh.extend(gidx[id(t)] for t in links.get("household", ()) if id(t) in gidx)
                vc.extend(gidx[id(t)] for t in links.get("care_home", ()) if id(t) in gidx)
                vh_off.append(len(vh))
                vc_off.append(len(vc))
            w.visit_household_off, w.visit_household = np.array(vh_off, np.uint32), np.array(vh, np.uint32)
            w.visit_care_home_off, w.visit_care_home = np.array(vc_off, np.uint32), np.array(vc, np.uint32)
        w.area_venue_off, w.area_venue = np.array(off, np.uint32), np.array(ven, np.uint32)
    w.validate()
    return w


def group_index_map(world_soa: WorldSoA, world) -> Dict[Tuple[str, int], int]:
    """(spec, reference group id) -> device group index."""
    out = {}
    for sg in world_soa.supergroups:
        for k in range(sg.n_groups):
            gi = sg.first_group + k
            out[(sg.spec, int(world_soa.group_ids[gi]))] = gi
    return out
