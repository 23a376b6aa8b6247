"""Synthetic worlds generated directly in struct-of-arrays form (no Python objects), following
the recipe of SURVEY.md section 8(d): households by size pmf, schools with year classes split at 40
pupils, lognormal companies with sectors A-U, care homes, hospitals; areas of ~300 people, super
areas of ~8000, 9 named regions.  The same recipe is used at 10 k, 2.7 M and 56 M people; the
reference's own world builder (``june/distributors``) needs census data that does not ship.

Multi-GPU: ``generate_domain`` builds one rank's domain; a fixed fraction of workers commute to
a company of a neighbouring domain.  ``wire_halos`` (single process, any number of virtual
domains) or ``wire_halos_distributed`` (one process per rank over ``torch.distributed``) exchange
the commuter lists once at set-up and append the visitors as halo members to the destination
companies -- the static form of the reference's per-step ``MovablePeople`` exchange
(``june/mpi_wrapper.py:99-293``, ``june/domains/domain.py``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .world import (ACT_COMMUTE, ACT_LEISURE, ACT_MEDICAL, ACT_PRIMARY, ACT_RESIDENCE, KIND_MATRIX, KIND_SCHOOL, BetaRule, Supergroup,
                    WorldSoA, make_ginfo, make_info, make_person_attr)

REGIONS = ["North East", "North West", "Yorkshire and The Humber", "East Midlands", "West Midlands",
           "East of England", "London", "South East", "South West"]
SECTORS = list("ABCDEFGHIJKLMNOPQRSTU")
HH_SIZES = np.array([1, 2, 3, 4, 5, 6])
HH_P = np.array([0.30, 0.34, 0.16, 0.13, 0.05, 0.02])
MAX_CLASSROOM = 40  # school_distributor.yaml: max_classroom_size
# leisure venues (config 4): supergroup name, spec, people per venue, candidates per area
# (pubs.yaml / cinemas.yaml / groceries.yaml / gyms.yaml: neighbours_to_consider), in the order of
# activity_to_super_groups["leisure"]; LEISURE_ORDER is the reference's distributor order
# (generate_leisure_for_world, leisure.py:33-91)
LEISURE_VENUES = [("pubs", "pub", 1000, 7), ("cinemas", "cinema", 20000, 5), ("groceries", "grocery", 2000, 15),
                  ("gyms", "gym", 5000, 7)]
LEISURE_ORDER = ["pubs", "gyms", "cinemas", "groceries"]
LEISURE_BIN_CAP = 2048  # visitors per venue and step
# commute (config 2 of BASELINE.json): one city per CITY_SUPER_AREAS super areas with CITY_STATIONS
# stations; 30 % of the company workers use public transport, 60 % of them inside their city, the
# others through the inter-city station of their home super area; ~50 commuters per transport
CITY_SUPER_AREAS, CITY_STATIONS, SEATS = 20, 4, 50
TRANSPORT_BIN_CAP = 512


def age_pmf() -> np.ndarray:
    """Coarse England-like age pyramid over 0..99 (fixed table, part of the recipe)."""
    w = np.empty(100)
    w[:20] = 1.18
    w[20:40] = 1.32
    w[40:60] = 1.33
    w[60:70] = 1.08
    w[70:80] = 0.85
    w[80:90] = 0.42
    w[90:] = 0.09
    return w / w.sum()


@dataclass
class DomainDraft:
    world: WorldSoA
    # outbound commuters: local person indices, grouped by destination rank
    out_person: Dict[int, np.ndarray] = field(default_factory=dict)
    company_first: int = 0
    company_count: int = 0
    company_size: Optional[np.ndarray] = None
    rank: int = 0
    world_size: int = 1
    seed: int = 0


def _csr_from_rows(group: np.ndarray, lsg: np.ndarray, member: np.ndarray, act: np.ndarray, n_groups: int):
    key = (group.astype(np.int64) << 36) | (lsg.astype(np.int64) << 28) | member.astype(np.int64)
    order = np.argsort(key, kind="stable")
    group, lsg, member, act = group[order], lsg[order], member[order], act[order]
    counts = np.bincount(group, minlength=n_groups)
    off = np.zeros(n_groups + 1, np.uint32)
    off[1:] = np.cumsum(counts)
    return off, member.astype(np.uint32), make_info(lsg, act).astype(np.uint16)


def generate_domain(n_people: int, rank: int = 0, world_size: int = 1, seed: int = 0, commute_fraction: float = 0.05,
                    person_id_base: Optional[int] = None, sector_betas: Optional[Dict[str, float]] = None,
                    leisure: bool = False, commute: bool = False, visits: bool = False) -> DomainDraft:
    """``visits`` (with ``leisure``): households are linked to 2-4 households of their super area and, one household
    per care-home resident, to that care home (``residence_visits.py:189-296``); "residence_visits" becomes the last
    leisure activity."""
    rng = np.random.default_rng([seed, rank, 20200301])
    pmf = age_pmf()
    adult_pmf = pmf.copy()
    adult_pmf[:18] = 0
    adult_pmf /= adult_pmf.sum()

    # ---- households ---------------------------------------------------------------------------
    n_hh_guess = int(n_people / float((HH_SIZES * HH_P).sum()) * 1.02) + 8
    sizes = rng.choice(HH_SIZES, size=n_hh_guess, p=HH_P)
    csum = np.cumsum(sizes)
    n_hh = int(np.searchsorted(csum, n_people, side="left")) + 1
    sizes = sizes[:n_hh].copy()
    sizes[-1] -= int(csum[n_hh - 1] - n_people)
    if sizes[-1] <= 0:
        sizes = sizes[:-1]
        n_hh -= 1
        sizes[-1] += n_people - int(sizes.sum())
    hh_start = np.zeros(n_hh + 1, np.int64)
    hh_start[1:] = np.cumsum(sizes)
    n = int(hh_start[-1])
    assert n == n_people
    hh_of = np.repeat(np.arange(n_hh, dtype=np.int64), sizes)
    age = rng.choice(100, size=n, p=pmf).astype(np.uint8)
    first = hh_start[:-1]
    age[first] = rng.choice(100, size=n_hh, p=adult_pmf).astype(np.uint8)  # at least one adult per household
    sex = (rng.random(n) < 0.5).astype(np.uint8)

    # ---- geography: areas ~300 people, super areas ~8000, 9 regions ----------------------------
    area_of_hh = (hh_start[:-1] // 300).astype(np.int64)
    n_areas = int(area_of_hh.max()) + 1
    sa_of_area = np.arange(n_areas) // 27
    n_sa = int(sa_of_area.max()) + 1
    sa_of_hh = sa_of_area[area_of_hh]
    psa = sa_of_hh[hh_of].astype(np.uint32)
    # regions are contiguous blocks of super areas; with domain decomposition each rank covers a slice
    g_sa = rank * n_sa + np.arange(n_sa)
    sa_region = ((g_sa * len(REGIONS)) // max(1, n_sa * world_size)).astype(np.uint16)

    rows_g: List[np.ndarray] = []
    rows_l: List[np.ndarray] = []
    rows_m: List[np.ndarray] = []
    rows_a: List[np.ndarray] = []

    def add_rows(g, l, m, a):
        rows_g.append(np.asarray(g, np.int64))
        rows_l.append(np.asarray(l, np.int64))
        rows_m.append(np.asarray(m, np.int64))
        rows_a.append(np.full(len(m), a, np.int64))

    has_activity = np.zeros(n, np.uint8)
    ch_res = np.zeros(n, np.uint8)
    free_worker = (age >= 19) & (age <= 64) & (rng.random(n) < 0.75)

    # ---- care homes: 1 per ~3700 people, ~25 residents aged >= 65 -------------------------------
    n_ch = max(1, n // 3700)
    old = np.nonzero(age >= 65)[0]
    rng.shuffle(old)
    n_res = min(len(old), n_ch * 25)
    ch_residents = np.sort(old[:n_res])
    ch_of_res = (np.arange(n_res) * n_ch // max(1, n_res)).astype(np.int64)
    in_care = np.zeros(n, bool)
    in_care[ch_residents] = True
    ch_res[ch_residents] = 1

    # ---- schools: per super area 2 primary (4-10) + 1 secondary (11-18) -------------------------
    kid = np.nonzero((age >= 4) & (age <= 18) & ~in_care)[0]
    ksa = psa[kid].astype(np.int64)
    is_sec = age[kid] >= 11
    pick = rng.integers(0, 2, size=len(kid))
    school_local = np.where(is_sec, 2, pick)          # 0,1 primary; 2 secondary
    school_of_kid = ksa * 3 + school_local
    n_schools = n_sa * 3
    school_min = np.tile(np.array([4, 4, 11]), n_sa)
    school_max = np.tile(np.array([10, 10, 18]), n_sa)
    school_sector = np.tile(np.array([0, 0, 1]), n_sa)  # 0 primary, 1 secondary
    school_sa = np.repeat(np.arange(n_sa), 3)
    # year classes split at MAX_CLASSROOM pupils (School.limit_classroom_sizes, school.py:123-155)
    year = age[kid].astype(np.int64) - school_min[school_of_kid]
    sy_key = school_of_kid * 32 + year
    order = np.argsort(sy_key, kind="stable")
    kid_s, sy_s = kid[order], sy_key[order]
    uniq, start, cnt = np.unique(sy_s, return_index=True, return_counts=True)
    n_classes_per = np.where(cnt > MAX_CLASSROOM, -(-cnt // MAX_CLASSROOM), 1)
    # class index of every pupil inside its (school, year): np.array_split semantics
    rank_in = np.arange(len(kid_s)) - np.repeat(start, cnt)
    cnt_r, ncl_r = np.repeat(cnt, cnt), np.repeat(n_classes_per, cnt)
    base, extra = cnt_r // ncl_r, cnt_r % ncl_r
    big = extra * (base + 1)
    cls_in = np.where(rank_in < big, rank_in // np.maximum(base + 1, 1), extra + (rank_in - big) // np.maximum(base, 1))
    # local subgroup = 1 + (classes of earlier years of that school) + class within year
    sch_of_u = uniq // 32
    cls_before = np.zeros(len(uniq), np.int64)
    new_school = np.r_[True, sch_of_u[1:] != sch_of_u[:-1]]
    csum_cls = np.cumsum(n_classes_per) - n_classes_per
    first_of_school = np.maximum.accumulate(np.where(new_school, np.arange(len(uniq)), 0))
    cls_before = csum_cls - csum_cls[first_of_school]
    lsg_kid = 1 + np.repeat(cls_before, cnt) + cls_in
    # school years table (age of every class), per school
    years_age = np.repeat(uniq % 32 + school_min[sch_of_u], n_classes_per).astype(np.uint8)
    years_school = np.repeat(sch_of_u, n_classes_per)
    n_years_per_school = np.bincount(years_school, minlength=n_schools)
    years_off = np.zeros(n_schools + 1, np.int64)
    years_off[1:] = np.cumsum(n_years_per_school)
    assert n_years_per_school.max() < 255
    pupils_per_school = np.bincount(school_of_kid, minlength=n_schools)
    has_activity[kid] |= 1 << ACT_PRIMARY

    # ---- workers ---------------------------------------------------------------------------------
    workers = np.nonzero(free_worker & ~in_care)[0]
    rng.shuffle(workers)
    wpos = 0

    def take(k):
        nonlocal wpos
        out = workers[wpos: wpos + k]
        wpos += len(out)
        return out

    # teachers: 1 per ~20 pupils (subgroup 0)
    n_teach = np.maximum(1, pupils_per_school // 20)
    n_teach[pupils_per_school == 0] = 0
    teachers = take(int(n_teach.sum()))
    teach_school = np.repeat(np.arange(n_schools), n_teach)[: len(teachers)]
    # hospitals: 1 per ~50000 people, ~0.6 % of the population works there
    n_hosp = max(1, n // 50000)
    hosp_workers = take(max(3 * n_hosp, int(0.006 * n)))
    hosp_of_worker = (np.arange(len(hosp_workers)) * n_hosp // max(1, len(hosp_workers))).astype(np.int64)
    # care home workers: 1 per 10 residents (care_home.yaml)
    n_chw = np.maximum(1, np.bincount(ch_of_res, minlength=n_ch) // 10)
    ch_workers = take(int(n_chw.sum()))
    chw_home = np.repeat(np.arange(n_ch), n_chw)[: len(ch_workers)]
    # commuters to another domain (static list; their company lives on the destination rank)
    rest = workers[wpos:]
    out_person: Dict[int, np.ndarray] = {}
    if world_size > 1 and commute_fraction > 0:
        n_comm = int(len(rest) * commute_fraction)
        comm = rest[:n_comm]
        rest = rest[n_comm:]
        # destinations: neighbouring domains first (ring), like commuting to a nearby super area
        hop = rng.choice(np.array([1, -1, 2, -2]), size=n_comm, p=[0.4, 0.4, 0.1, 0.1])
        dest = (rank + hop) % world_size
        dest = np.where(dest == rank, (rank + 1) % world_size, dest)
        for q in range(world_size):
            sel = np.sort(comm[dest == q])
            if len(sel):
                out_person[q] = sel.astype(np.uint32)
        has_activity[comm] |= 1 << ACT_PRIMARY
    # companies: lognormal sizes (median 8, cap 1500), sector uniform over A-U
    n_w = len(rest)
    est = int(n_w / 14.0 * 1.3) + 16
    csize = np.minimum(1500, np.maximum(1, np.rint(rng.lognormal(np.log(8.0), 1.1, est)))).astype(np.int64)
    ccum = np.cumsum(csize)
    while ccum[-1] < n_w:
        more = np.minimum(1500, np.maximum(1, np.rint(rng.lognormal(np.log(8.0), 1.1, est)))).astype(np.int64)
        csize = np.r_[csize, more]
        ccum = np.cumsum(csize)
    n_comp = int(np.searchsorted(ccum, n_w, side="left")) + 1
    csize = csize[:n_comp].copy()
    csize[-1] -= int(ccum[n_comp - 1] - n_w)
    comp_of_worker = np.repeat(np.arange(n_comp, dtype=np.int64), csize)
    # workers of a company mostly live near it: sort the remaining workers by home super area, then
    # shuffle inside windows so that companies draw from a neighbourhood, not the whole domain
    rest = rest[np.argsort(psa[rest] + rng.normal(0, 1.5, len(rest)), kind="stable")]
    comp_sa = psa[rest][np.minimum(np.cumsum(csize) - 1, n_w - 1)].astype(np.int64) if n_w else np.zeros(n_comp, np.int64)
    comp_sector = rng.integers(0, len(SECTORS), n_comp)
    has_activity[teachers] |= 1 << ACT_PRIMARY
    has_activity[hosp_workers] |= 1 << ACT_PRIMARY
    has_activity[ch_workers] |= 1 << ACT_PRIMARY
    has_activity[rest] |= 1 << ACT_PRIMARY
    has_activity[:] |= 1 << ACT_RESIDENCE

    # ---- group table, in the reference's supergroup order ----------------------------------------
    sector_betas = sector_betas or {}
    g_hosp0 = 0
    # commuters and their transports (commute precedes the primary activity in the hierarchy)
    n_city = max(1, -(-n_sa // CITY_SUPER_AREAS)) if commute else 0
    person_commute = np.full(n, -1, np.int32)
    n_ct = n_it = 0
    if commute:
        crng = np.random.default_rng([seed, rank, 5150])
        pub = rest[crng.random(len(rest)) < 0.30]
        internal = crng.random(len(pub)) < 0.60
        city_of = (psa[pub] // CITY_SUPER_AREAS).astype(np.int64)
        person_commute[pub[internal]] = (city_of[internal] << 1).astype(np.int32)
        # inter-city stations: one per super area (closest_inter_city_station_for_city), index after the city stations
        n_cstation = n_city * CITY_STATIONS
        person_commute[pub[~internal]] = (((n_cstation + psa[pub[~internal]].astype(np.int64)) << 1) | 1).astype(np.int32)
        per_cstation = np.bincount(city_of[internal], minlength=n_city) / CITY_STATIONS
        ct_per_station = np.maximum(1, np.ceil(np.repeat(per_cstation, CITY_STATIONS) / SEATS)).astype(np.int64)
        it_per_station = np.maximum(1, np.ceil(np.bincount(psa[pub[~internal]], minlength=n_sa) / SEATS)).astype(np.int64)
        n_ct, n_it = int(ct_per_station.sum()), int(it_per_station.sum())
    g_ct0 = g_hosp0 + n_hosp
    g_it0 = g_ct0 + n_ct
    g_school0 = g_it0 + n_it
    g_comp0 = g_school0 + n_schools
    # leisure venues sit between the primary-activity and the residence supergroups (hierarchy order)
    n_venues = [max(1, n // per) for _, _, per, _ in LEISURE_VENUES] if leisure else []
    g_ven0 = [g_comp0 + n_comp + int(sum(n_venues[:i])) for i in range(len(n_venues))]
    g_hh0 = g_comp0 + n_comp + int(sum(n_venues))
    g_ch0 = g_hh0 + n_hh
    G = g_ch0 + n_ch
    rules = [BetaRule("hospital"), BetaRule("primary_school", "school"), BetaRule("secondary_school", "school"),
             BetaRule("company"), BetaRule("household", None, False), BetaRule("household_visits", None, False),
             BetaRule("care_visits", None, False), BetaRule("care_home")]
    if leisure:
        rules += [BetaRule(spec) for _, spec, _, _ in LEISURE_VENUES]
    if commute:
        rules += [BetaRule("city_transport"), BetaRule("inter_city_transport")]
    scales = [1.0]
    sec_scale = np.zeros(len(SECTORS), np.int64)
    for i, s in enumerate(SECTORS):
        x = float(sector_betas.get(s, 1.0))
        if x not in scales:
            scales.append(x)
        sec_scale[i] = scales.index(x)
    beta_idx = np.zeros(G, np.uint32)
    scale_idx = np.zeros(G, np.uint32)
    region = np.zeros(G, np.uint32)
    aux = np.zeros(G, np.uint32)
    hosp_sa = (np.arange(n_hosp) * n_sa // n_hosp).astype(np.int64)
    beta_idx[g_hosp0:g_hosp0 + n_hosp] = 0
    region[g_hosp0:g_hosp0 + n_hosp] = sa_region[hosp_sa]
    beta_idx[g_school0:g_comp0] = 1 + school_sector
    region[g_school0:g_comp0] = sa_region[school_sa]
    aux[g_school0:g_comp0] = (years_off[:-1].astype(np.uint32) << 8) | n_years_per_school.astype(np.uint32)
    beta_idx[g_comp0:g_comp0 + n_comp] = 3
    scale_idx[g_comp0:g_comp0 + n_comp] = sec_scale[comp_sector]
    region[g_comp0:g_comp0 + n_comp] = sa_region[comp_sa]
    beta_idx[g_hh0:g_ch0] = 4
    region[g_hh0:g_ch0] = sa_region[sa_of_hh]
    ch_sa = (np.arange(n_ch) * n_sa // n_ch).astype(np.int64)
    beta_idx[g_ch0:] = 7
    region[g_ch0:] = sa_region[ch_sa]
    grp_sa = np.zeros(G, np.uint32)
    grp_sa[g_hosp0:g_hosp0 + n_hosp] = hosp_sa
    grp_sa[g_school0:g_comp0] = school_sa
    grp_sa[g_comp0:g_comp0 + n_comp] = comp_sa
    grp_sa[g_hh0:g_ch0] = sa_of_hh
    grp_sa[g_ch0:] = ch_sa
    commute_tables = None
    if commute:
        bi = len(rules) - 2
        st_sa = np.concatenate([np.repeat(np.arange(n_city) * CITY_SUPER_AREAS, CITY_STATIONS), np.arange(n_sa)]).clip(0, n_sa - 1)
        per = np.concatenate([ct_per_station, it_per_station])
        st_off = np.zeros(len(per) + 1, np.uint32)
        st_off[1:] = np.cumsum(per)
        tr_station = np.repeat(np.arange(len(per)), per)
        tr_group = (g_ct0 + np.arange(n_ct + n_it)).astype(np.uint32)
        beta_idx[g_ct0:g_it0] = bi
        beta_idx[g_it0:g_school0] = bi + 1
        region[g_ct0:g_school0] = sa_region[st_sa[tr_station]]
        grp_sa[g_ct0:g_school0] = st_sa[tr_station]
        cs_off = (np.arange(n_city + 1) * CITY_STATIONS).astype(np.uint32)
        commute_tables = (cs_off, np.arange(n_city * CITY_STATIONS, dtype=np.uint32), st_off, tr_group)
    # venues are spread evenly over the areas; an area's candidates are the nearest ones by index
    area_venue_off = area_venue = None
    if leisure:
        K = len(LEISURE_VENUES)
        cand = []  # [activity in LEISURE_ORDER][area] -> group indices
        for name in LEISURE_ORDER:
            i = [v[0] for v in LEISURE_VENUES].index(name)
            nv, ncand = n_venues[i], min(n_venues[i], LEISURE_VENUES[i][3])
            ven_area = (np.arange(nv) * n_areas // nv).astype(np.int64)
            ven_sa = sa_of_area[ven_area]
            beta_idx[g_ven0[i]: g_ven0[i] + nv] = 8 + i
            region[g_ven0[i]: g_ven0[i] + nv] = sa_region[ven_sa]
            grp_sa[g_ven0[i]: g_ven0[i] + nv] = ven_sa
            centre = np.clip((np.arange(n_areas) * nv) // n_areas - ncand // 2, 0, nv - ncand)
            cand.append(g_ven0[i] + centre[:, None] + np.arange(ncand)[None, :])
        if visits:  # the visits activity has no venues: an empty column of the (area, activity) table
            cand.append(np.zeros((n_areas, 0), np.int64))
            K += 1
        counts = np.stack([np.full(n_areas, c.shape[1]) for c in cand], axis=1).ravel()  # (area, activity) order
        area_venue_off = np.zeros(n_areas * K + 1, np.uint32)
        area_venue_off[1:] = np.cumsum(counts)
        area_venue = np.concatenate([np.concatenate([cand[k][a] for k in range(K)]) for a in range(n_areas)]).astype(np.uint32)
    visit_tables = None
    if leisure and visits:
        vrng = np.random.default_rng([seed, rank, 777])
        # link_households_to_households: every household gets randint(2, 4) households of its super area
        n_link = vrng.integers(2, 5, n_hh)
        sa_first = np.searchsorted(sa_of_hh, np.arange(n_sa), side="left")
        sa_cnt = np.bincount(sa_of_hh, minlength=n_sa)
        src = np.repeat(np.arange(n_hh), n_link)
        pick = sa_first[sa_of_hh[src]] + (vrng.random(len(src)) * sa_cnt[sa_of_hh[src]]).astype(np.int64)
        pick = np.where(pick == src, sa_first[sa_of_hh[src]] + (pick - sa_first[sa_of_hh[src]] + 1) % np.maximum(1, sa_cnt[sa_of_hh[src]]), pick)
        keep = pick != src
        vh_off = np.zeros(n_hh + 1, np.uint32)
        vh_off[1:] = np.cumsum(np.bincount(src[keep], minlength=n_hh))
        vh = (g_hh0 + pick[keep]).astype(np.uint32)
        # link_households_to_care_homes: one household of two or more people in the care home's super area per resident
        big = np.nonzero(sizes >= 2)[0]
        big_sa = sa_of_hh[big]
        want_sa = ch_sa[ch_of_res]
        lo = np.searchsorted(big_sa, want_sa, side="left")
        hi = np.searchsorted(big_sa, want_sa, side="right")
        ok = hi > lo
        chosen = big[(lo[ok] + (vrng.random(int(ok.sum())) * (hi[ok] - lo[ok])).astype(np.int64))]
        order = np.argsort(chosen, kind="stable")
        vc_off = np.zeros(n_hh + 1, np.uint32)
        vc_off[1:] = np.cumsum(np.bincount(chosen, minlength=n_hh))
        vc = (g_ch0 + ch_of_res[ok][order]).astype(np.uint32)
        visit_tables = (vh_off, vh, vc_off, vc)

    # ---- registrations ----------------------------------------------------------------------------
    add_rows(g_hosp0 + hosp_of_worker, np.zeros(len(hosp_workers)), hosp_workers, ACT_PRIMARY)
    add_rows(g_school0 + teach_school, np.zeros(len(teachers)), teachers, ACT_PRIMARY)
    add_rows(g_school0 + (sy_s // 32), lsg_kid, kid_s, ACT_PRIMARY)
    add_rows(g_comp0 + comp_of_worker, np.zeros(n_w), rest, ACT_PRIMARY)
    home = np.nonzero(~in_care)[0]
    a = age[home]
    hh_lsg = np.where(a < 18, 0, np.where(a <= 25, 1, np.where(a < 65, 2, 3)))  # household.py:109-122
    add_rows(g_hh0 + hh_of[home], hh_lsg, home, ACT_RESIDENCE)
    add_rows(g_ch0 + chw_home, np.zeros(len(ch_workers)), ch_workers, ACT_PRIMARY)
    add_rows(g_ch0 + ch_of_res, np.ones(n_res), ch_residents, ACT_RESIDENCE)
    grp_offset, reg_member, reg_info = _csr_from_rows(np.concatenate(rows_g), np.concatenate(rows_l),
                                                      np.concatenate(rows_m), np.concatenate(rows_a), G)

    sa_hospital = (g_hosp0 + (np.arange(n_sa) * n_hosp // n_sa)).astype(np.int32)
    supergroups = [
        Supergroup("hospitals", "hospital", ACT_MEDICAL, g_hosp0, n_hosp, KIND_MATRIX, 1, 1),
    ] + ([Supergroup("city_transports", "city_transport", ACT_COMMUTE, g_ct0, n_ct, KIND_MATRIX, 1, TRANSPORT_BIN_CAP),
          Supergroup("inter_city_transports", "inter_city_transport", ACT_COMMUTE, g_it0, n_it, KIND_MATRIX, 1, TRANSPORT_BIN_CAP)]
         if commute else []) + [
        Supergroup("schools", "school", ACT_PRIMARY, g_school0, n_schools, KIND_SCHOOL, 1, 0),
        Supergroup("companies", "company", ACT_PRIMARY, g_comp0, n_comp, KIND_MATRIX, 0, 0),
    ] + [Supergroup(name, spec, ACT_LEISURE, g_ven0[i], n_venues[i], KIND_MATRIX, 1, LEISURE_BIN_CAP)
         for i, (name, spec, _, _) in enumerate(LEISURE_VENUES if leisure else [])] + [
        Supergroup("households", "household", ACT_RESIDENCE, g_hh0, n_hh, KIND_MATRIX, 0, 0),
        # (care homes take visitors as dynamic members when residence visits are on)
        Supergroup("care_homes", "care_home", ACT_RESIDENCE, g_ch0, n_ch, KIND_MATRIX, 1 if visit_tables else 0, 64 if visit_tables else 0),
    ]
    w = WorldSoA(
        age=age, sex=sex, care_home_resident=ch_res, has_activity=has_activity, super_area=psa,
        sa_hospital=sa_hospital, sa_region=sa_region, region_names=list(REGIONS),
        grp_offset=grp_offset, grp_info=make_ginfo(beta_idx, scale_idx, region).astype(np.uint32), grp_aux=aux,
        reg_member=reg_member, reg_info=reg_info, school_years=years_age, supergroups=supergroups,
        beta_rules=rules, beta_scale=np.array(scales, np.float64),
        person_id_base=(rank * n_people if person_id_base is None else person_id_base),
        grp_super_area=grp_sa,
    )
    # lockdown status of the workers (key 19 %, furlough 4 %, else "random": policy_covid19.yaml
    # close_companies), spec of the primary activity, and the two-key-worker-household flag
    lock = np.zeros(n, np.uint8)
    wk = np.concatenate([teachers, hosp_workers, ch_workers, rest] + list(out_person.values())).astype(np.int64)
    draw = np.random.default_rng([seed, rank, 4242]).random(len(wk))
    lock[wk] = np.where(draw < 0.19, 1, np.where(draw < 0.23, 3, 2))
    lock[hosp_workers] = 1
    kind = np.zeros(n, np.uint8)
    kind[kid] = 1
    kind[teachers] = 1
    kind[rest] = 2
    kind[hosp_workers] = 4
    kind[ch_workers] = 4
    for sel in out_person.values():
        kind[sel] = 2
    n_key = np.bincount(hh_of, weights=(lock == 1), minlength=n_hh)
    w.person_attr = make_person_attr(lock, kind, (n_key[hh_of] > 1) & ~in_care)
    if commute:
        w.person_commute = person_commute
        w.city_station_off, w.city_station, w.station_transport_off, w.station_transport = commute_tables
    if leisure:
        w.person_area = area_of_hh[hh_of].astype(np.uint32)
        w.leisure_supergroups = list(LEISURE_ORDER)
        w.leisure_activities = [dict((v[0], v[1]) for v in LEISURE_VENUES)[n_] for n_ in LEISURE_ORDER]
        w.leisure_subgroups = [0] * len(LEISURE_ORDER)
        w.area_venue_off, w.area_venue = area_venue_off, area_venue
        if visit_tables:
            from .world import VISITS_ACTIVITY, VISITS_SUPERGROUP
            w.leisure_supergroups.append(VISITS_SUPERGROUP)
            w.leisure_activities.append(VISITS_ACTIVITY)
            w.leisure_subgroups.append(0)
            w.visit_household_off, w.visit_household, w.visit_care_home_off, w.visit_care_home = visit_tables
    w.validate()
    return DomainDraft(world=w, out_person=out_person, company_first=g_comp0, company_count=n_comp,
                       company_size=csize, rank=rank, world_size=world_size, seed=seed)


def _attach_halo(draft: DomainDraft, inbound_gids: List[np.ndarray]) -> None:
    """Append inbound commuters (one gid array per source rank, rank order) as halo members of
    this domain's companies, and fix the outbound slot order (destination-rank order)."""
    w = draft.world
    n = w.n_people
    rng = np.random.default_rng([draft.seed, draft.rank, 777])
    counts_in = np.array([len(g) for g in inbound_gids], np.int64)
    halo_gid = np.concatenate(inbound_gids).astype(np.uint32) if counts_in.sum() else np.zeros(0, np.uint32)
    H = len(halo_gid)
    out_list, out_counts = [], np.zeros(draft.world_size, np.int64)
    for q in range(draft.world_size):
        sel = draft.out_person.get(q)
        if sel is not None and len(sel):
            out_list.append(sel)
            out_counts[q] = len(sel)
    w.out_person = np.concatenate(out_list).astype(np.uint32) if out_list else np.zeros(0, np.uint32)
    w.out_counts, w.in_counts = out_counts, counts_in
    if H:
        # visitors join a company with probability proportional to its size
        p = draft.company_size / draft.company_size.sum()
        comp = rng.choice(draft.company_count, size=H, p=p) + draft.company_first
        grp_of = np.repeat(np.arange(w.n_groups, dtype=np.int64), np.diff(w.grp_offset.astype(np.int64)))
        g_all = np.concatenate([grp_of, comp.astype(np.int64)])
        l_all = np.concatenate([(w.reg_info & 0xFF).astype(np.int64), np.zeros(H, np.int64)])
        m_all = np.concatenate([w.reg_member.astype(np.int64), n + np.arange(H, dtype=np.int64)])
        a_all = np.concatenate([(w.reg_info >> 8).astype(np.int64), np.full(H, ACT_PRIMARY, np.int64)])
        w.grp_offset, w.reg_member, w.reg_info = _csr_from_rows(g_all, l_all, m_all, a_all, w.n_groups)
    w.n_halo = H
    w.halo_gid = halo_gid
    w.validate()


def wire_halos(drafts: List[DomainDraft]) -> List[WorldSoA]:
    """Single-process wiring of K virtual domains (tests; 1-GPU emulation of K ranks)."""
    K = len(drafts)
    for r, d in enumerate(drafts):
        inbound = []
        for q in range(K):
            sel = drafts[q].out_person.get(r)
            base = drafts[q].world.person_id_base
            inbound.append((sel.astype(np.int64) + base).astype(np.uint32) if sel is not None else np.zeros(0, np.uint32))
        _attach_halo(d, inbound)
    return [d.world for d in drafts]


def wire_halos_distributed(draft: DomainDraft, dist) -> WorldSoA:
    """One process per rank: exchange the commuter gid lists over ``torch.distributed``
    (any backend; objects are small and this runs once at set-up)."""
    K = draft.world_size
    base = draft.world.person_id_base
    send = [((draft.out_person[q].astype(np.int64) + base).astype(np.uint32) if q in draft.out_person
             else np.zeros(0, np.uint32)) for q in range(K)]
    recv: List[Optional[np.ndarray]] = [None] * K
    gathered: List[object] = [None] * K
    dist.all_gather_object(gathered, send)
    for q in range(K):
        recv[q] = gathered[q][draft.rank]
    _attach_halo(draft, recv)
    return draft.world


def generate_world(n_people: int, seed: int = 0, sector_betas: Optional[Dict[str, float]] = None,
                   leisure: bool = False, commute: bool = False, visits: bool = False) -> WorldSoA:
    """Single-domain world (configs 0/1 of BASELINE.json; ``leisure=True`` adds the venues of config 3,
    ``visits=True`` its household and care-home visits)."""
    d = generate_domain(n_people, 0, 1, seed, sector_betas=sector_betas, leisure=leisure, commute=commute, visits=visits)
    _attach_halo(d, [np.zeros(0, np.uint32)])
    return d.world
