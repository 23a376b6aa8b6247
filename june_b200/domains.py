"""Domain decomposition by super area: split ONE world into K domains, one per GPU.

Replaces ``june/domains`` (``DomainSplitter.generate_world_split``,
``domain_decomposition.py:70-151``; ``Domain.from_hdf5``, ``domain.py:35-73``) on the SoA form:

* a person belongs to the domain of their home super area, a group to the domain of its super
  area (``WorldSoA.grp_super_area``);
* domains are contiguous ranges of super areas balanced by the reference's score
  ``5*people + workers + pupils + commuters`` per super area (``domain_decomposition.py:22,107-113``;
  the reference clusters the scored super areas with a library that is not available -- for
  geographies numbered along a space-filling order a balanced prefix split of the same score plays
  that role);
* a registered member that lives in another domain becomes a HALO person of the group's domain
  (the static form of the reference's ``ExternalSubgroup`` + ``MovablePeople`` pair,
  ``groups/group/external.py``, ``mpi_wrapper.py:99-293``): its state arrives every step through
  the people all-to-all and infections found here travel back through the second one.

Requires people ordered by home super area (true for ``june_b200.synth`` and for worlds compiled
area by area), so that every domain owns a contiguous range of global person ids.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .world import ACT_PRIMARY, ACT_RESIDENCE, Supergroup, WorldSoA


#: ``default_weights`` of the reference (``domain_decomposition.py:22``)
SCORE_WEIGHTS = {"population": 5.0, "workers": 1.0, "commuters": 1.0}


def super_area_scores(world: WorldSoA, weights=None) -> np.ndarray:
    """The reference's load score per super area (``DomainSplitter.get_score``, ``domain_decomposition.py:107-113``):
    ``population * n_people + workers * (n_workers + n_pupils) + commuters * n_commuters``.  Workers and pupils are
    counted where they work / learn (the primary-activity registrations of the groups of the super area,
    ``geography_saver.py:111-121``), commuters where they live."""
    w = dict(SCORE_WEIGHTS if weights is None else weights)
    S = world.n_super_areas
    psa = world.super_area.astype(np.int64)
    score = w["population"] * np.bincount(psa, minlength=S).astype(np.float64)
    if world.grp_super_area is not None and len(world.reg_member):
        grp_of = np.repeat(np.arange(world.n_groups, dtype=np.int64), np.diff(world.grp_offset.astype(np.int64)))
        primary = (world.reg_info >> 8) == ACT_PRIMARY
        score += w["workers"] * np.bincount(world.grp_super_area.astype(np.int64)[grp_of[primary]], minlength=S)
    if world.person_commute is not None:
        score += w["commuters"] * np.bincount(psa[world.person_commute >= 0], minlength=S)
    return score


def balanced_split(sa_population: np.ndarray, n_domains: int) -> np.ndarray:
    """First super area of every domain (+ end sentinel): contiguous, balanced in the given per-super-area load."""
    cum = np.cumsum(sa_population)
    total = cum[-1]
    cuts = [0]
    for d in range(1, n_domains):
        cuts.append(int(np.searchsorted(cum, total * d / n_domains, side="left")) + 1)
    cuts.append(len(sa_population))
    cuts = np.maximum.accumulate(np.minimum(np.array(cuts), len(sa_population)))
    return cuts


def partition(world: WorldSoA, n_domains: int) -> List[WorldSoA]:
    assert world.n_halo == 0, "partition a whole world, not a domain"
    assert world.grp_super_area is not None, "WorldSoA.grp_super_area is needed to place groups"
    n, G = world.n_people, world.n_groups
    psa = world.super_area.astype(np.int64)
    assert np.all(np.diff(psa) >= 0), "people must be ordered by home super area"
    cuts = balanced_split(super_area_scores(world), n_domains)
    sa_dom = np.zeros(world.n_super_areas, np.int64)
    for d in range(n_domains):
        sa_dom[cuts[d]:cuts[d + 1]] = d
    p_dom = sa_dom[psa]
    g_dom = sa_dom[world.grp_super_area.astype(np.int64)]
    p_first = np.searchsorted(p_dom, np.arange(n_domains + 1), side="left")  # contiguous person ranges
    grp_of = np.repeat(np.arange(G, dtype=np.int64), np.diff(world.grp_offset.astype(np.int64)))
    m_person = world.reg_member.astype(np.int64)
    m_pdom, m_gdom = p_dom[m_person], g_dom[grp_of]
    sg_of = world.group_supergroup()

    # halo lists: for (group domain d, person domain q != d): the distinct persons of q registered in d
    inbound = [[np.zeros(0, np.int64) for _ in range(n_domains)] for _ in range(n_domains)]
    for d in range(n_domains):
        sel = (m_gdom == d) & (m_pdom != d)
        persons = np.unique(m_person[sel])
        pd = p_dom[persons]
        for q in range(n_domains):
            inbound[d][q] = persons[pd == q]

    out: List[WorldSoA] = []
    for d in range(n_domains):
        lo, hi = int(p_first[d]), int(p_first[d + 1])
        nd = hi - lo
        groups = np.nonzero(g_dom == d)[0]            # ascending == supergroup order preserved
        g_new = np.full(G, -1, np.int64)
        g_new[groups] = np.arange(len(groups))
        halo = np.concatenate(inbound[d]) if n_domains > 1 else np.zeros(0, np.int64)
        halo_idx = {int(p): nd + j for j, p in enumerate(halo)}
        sel = m_gdom == d
        mp, mg = m_person[sel], grp_of[sel]
        local = (mp >= lo) & (mp < hi)
        new_member = np.where(local, mp - lo, 0)
        if (~local).any():
            new_member[~local] = np.array([halo_idx[int(p)] for p in mp[~local]], np.int64)
        info = world.reg_info[sel]
        key = (g_new[mg] << 36) | ((info & 0xFF).astype(np.int64) << 28) | new_member
        order = np.argsort(key, kind="stable")
        counts = np.bincount(g_new[mg], minlength=len(groups))
        off = np.zeros(len(groups) + 1, np.uint32)
        off[1:] = np.cumsum(counts)
        sgs = []
        first = 0
        for k, sg in enumerate(world.supergroups):
            cnt = int((sg_of[groups] == k).sum())
            sgs.append(Supergroup(sg.name, sg.spec, sg.activity, first, cnt, sg.kind, sg.launch_mode, sg.has_dynamic))
            first += cnt
        # closest hospital must live in the same domain (patients are dynamic members; the reference
        # uses ExternalHospital for the cross-domain case, which is out of scope here)
        sa_hosp = world.sa_hospital.astype(np.int64).copy()
        hosp_here = [g for g in groups if world.supergroups[sg_of[g]].spec == "hospital"]
        for sa in range(world.n_super_areas):
            h = sa_hosp[sa]
            if h >= 0 and g_dom[h] != d:
                sa_hosp[sa] = hosp_here[sa % len(hosp_here)] if hosp_here else -1
        sa_hosp = np.where(sa_hosp >= 0, g_new[np.maximum(sa_hosp, 0)], -1).astype(np.int32)
        # outbound: my residents registered in groups of domain q, in q's halo order
        out_list, out_counts = [], np.zeros(n_domains, np.int64)
        for q in range(n_domains):
            if q == d:
                continue
            mine = inbound[q][d]
            out_list.append(mine - lo)
            out_counts[q] = len(mine)
        in_counts = np.array([len(inbound[d][q]) for q in range(n_domains)], np.int64)
        w = WorldSoA(
            age=world.age[lo:hi].copy(), sex=world.sex[lo:hi].copy(),
            care_home_resident=world.care_home_resident[lo:hi].copy(), has_activity=world.has_activity[lo:hi].copy(),
            super_area=world.super_area[lo:hi].copy(), sa_hospital=sa_hosp, sa_region=world.sa_region.copy(),
            region_names=list(world.region_names), grp_offset=off, grp_info=world.grp_info[groups].copy(),
            grp_aux=world.grp_aux[groups].copy(), reg_member=new_member[order].astype(np.uint32),
            reg_info=info[order].copy(), school_years=world.school_years.copy(), supergroups=sgs,
            beta_rules=list(world.beta_rules), beta_scale=world.beta_scale.copy(),
            n_halo=len(halo), person_id_base=world.person_id_base + lo,
            halo_gid=(halo + world.person_id_base).astype(np.uint32),
            out_person=(np.concatenate(out_list) if out_list else np.zeros(0, np.int64)).astype(np.uint32),
            out_counts=out_counts, in_counts=in_counts,
            grp_super_area=world.grp_super_area[groups].copy(),
        )
        # per-person facts the individual policies read travel with the residents
        if world.person_attr is not None:
            w.person_attr = world.person_attr[lo:hi].copy()
        if getattr(world, "person_work_region", None) is not None:
            w.person_work_region = world.person_work_region[lo:hi].copy()
        w.validate()
        out.append(w)
    return out


class LoopbackExchange:
    """K domains in ONE process: the two all-to-all exchanges as plain array copies.  Used by the
    CPU tests and to emulate K ranks on one GPU (no spinning kernels, see B200_PROFILING.md)."""

    def __init__(self, worlds: Sequence[WorldSoA], record_size: int):
        self.worlds, self.rec = list(worlds), record_size
        K = len(worlds)
        self.send = [np.zeros(max(1, int(w.out_counts.sum())) * record_size, np.uint8) for w in worlds]
        self.recv = [np.zeros(max(1, int(w.in_counts.sum())) * record_size, np.uint8) for w in worlds]
        self.ret_send = [np.zeros(max(1, int(w.in_counts.sum())), np.uint8) for w in worlds]
        self.ret_recv = [np.zeros(max(1, int(w.out_counts.sum())), np.uint8) for w in worlds]
        self.K = K

    def _a2a(self, src, dst, src_counts, dst_counts, unit):
        K = self.K
        soff = [np.concatenate([[0], np.cumsum(c)]) * unit for c in src_counts]
        doff = [np.concatenate([[0], np.cumsum(c)]) * unit for c in dst_counts]
        for r in range(K):
            for q in range(K):
                n = int(src_counts[r][q]) * unit
                if n:
                    assert int(dst_counts[q][r]) * unit == n
                    dst[q][doff[q][r]: doff[q][r] + n] = src[r][soff[r][q]: soff[r][q] + n]

    def exchange_people(self):
        self._a2a(self.send, self.recv, [w.out_counts for w in self.worlds], [w.in_counts for w in self.worlds], self.rec)

    def exchange_infections(self):
        self._a2a(self.ret_send, self.ret_recv, [w.in_counts for w in self.worlds], [w.out_counts for w in self.worlds], 1)
