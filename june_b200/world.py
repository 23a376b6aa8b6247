"""Device-facing struct-of-arrays description of a JUNE world (or of one domain of it).

This is the host-side (numpy) form of ``JbWorldDesc`` in ``include/june_b200.h``: people,
groups, and the registered membership CSR group -> (subgroup, member).  It replaces the object
graph the reference hot path walks (``june/world.py:63-98``, ``june/groups/group/group.py:25``,
``june/demography/person.py:35-81``).

Index spaces
------------
* persons ``[0, n_people)`` live in this domain; ``[n_people, n_people + n_halo)`` are halo
  persons (residents of another domain that are registered members of groups here; the device
  form of the visitors in ``june/mpi_wrapper.py:169-190``).
* groups are laid out supergroup after supergroup, in the order the reference iterates them in
  ``Simulator.do_timestep`` (``june/simulator.py:399-455``): activity-hierarchy order
  (``june/activity/activity_manager.py:24-33``), then the order of ``activity_to_super_groups``.
* inside a group, registered members are sorted by (local subgroup index, person index), which
  is the reference's traversal order of ``InteractiveGroup`` (``interactive.py:65-99``) when
  people were appended in ``world.people`` order.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

# activity codes, hierarchy order (activity_manager.py:24-33)
ACT_MEDICAL, ACT_COMMUTE, ACT_PRIMARY, ACT_LEISURE, ACT_RESIDENCE, ACT_NONE = 0, 1, 2, 3, 4, 7
ACTIVITY_CODES = {
    "medical_facility": ACT_MEDICAL,
    "commute": ACT_COMMUTE,
    "primary_activity": ACT_PRIMARY,
    "leisure": ACT_LEISURE,
    "residence": ACT_RESIDENCE,
}
ACTIVITY_NAMES = {v: k for k, v in ACTIVITY_CODES.items()}

PS_ACT_MASK, PS_INFECTED, PS_DEAD, PS_SEVERE, PS_MEDICAL = 0x07, 0x08, 0x10, 0x20, 0x40

KIND_MATRIX, KIND_SCHOOL = 0, 1
# person_attr fields
LOCKDOWN_CODES = {None: 0, "key_worker": 1, "random": 2, "furlough": 3}
PRIMARY_KIND_CODES = {"school": 1, "company": 2, "university": 3}
PA_TWO_KEY_WORKERS, PA_LONG_COMMUTER = 0x20, 0x40


def make_person_attr(lockdown, kind, two_key_workers, long_commuter=None):
    out = (np.asarray(lockdown, np.uint8) | (np.asarray(kind, np.uint8) << 2)
           | np.where(np.asarray(two_key_workers, bool), PA_TWO_KEY_WORKERS, 0).astype(np.uint8))
    if long_commuter is not None:
        out = out | np.where(np.asarray(long_commuter, bool), PA_LONG_COMMUTER, 0).astype(np.uint8)
    return out.astype(np.uint8)
SCHOOL_DIM = 32
MAX_STAGES = 8
N_TPAR = 5


def make_info(lsg, act):
    return (np.asarray(act, dtype=np.uint16) << 8) | np.asarray(lsg, dtype=np.uint16)


def make_ginfo(beta_idx, scale_idx, region):
    return (
        np.asarray(beta_idx, dtype=np.uint32)
        | (np.asarray(scale_idx, dtype=np.uint32) << 8)
        | (np.asarray(region, dtype=np.uint32) << 16)
    )


@dataclass
class Supergroup:
    """One contiguous range of groups of the same spec (one reference ``Supergroup``)."""

    name: str            # attribute name on the reference World ("households", "companies", ...)
    spec: str            # group spec: key into Interaction.contact_matrices / betas
    activity: int        # activity whose presence in timer.activities activates it
    first_group: int
    n_groups: int
    kind: int = KIND_MATRIX
    launch_mode: int = 1  # 0 thread per group, 1 warp per group
    has_dynamic: int = 0

    def to_json(self):
        return dict(self.__dict__)


@dataclass
class BetaRule:
    """How one row of the per-step beta table is computed (``get_processed_beta``)."""

    key: str                       # e.g. "company", "secondary_school", "household"
    fallback: Optional[str] = None  # school.py:495-502: fall back to "school"
    uses_tier: bool = True          # household.py:305-307 has no lockdown-tier term

    def to_json(self):
        return dict(self.__dict__)


VISITS_ACTIVITY = "residence_visits"      # key of the reference's distributor (leisure.py:84-91)
VISITS_SUPERGROUP = "residence_visits"    # entry of WorldSoA.leisure_supergroups for it (not a supergroup)
CARE_HOME_VISITORS_SUBGROUP = 2           # CareHome.SubgroupType.visitors (care_home.py)


@dataclass
class WorldSoA:
    # persons
    age: np.ndarray
    sex: np.ndarray
    care_home_resident: np.ndarray
    has_activity: np.ndarray
    super_area: np.ndarray
    # super areas
    sa_hospital: np.ndarray
    sa_region: np.ndarray
    region_names: List[str]
    # groups
    grp_offset: np.ndarray
    grp_info: np.ndarray
    grp_aux: np.ndarray
    reg_member: np.ndarray
    reg_info: np.ndarray
    school_years: np.ndarray
    supergroups: List[Supergroup]
    beta_rules: List[BetaRule]
    beta_scale: np.ndarray
    # domain decomposition
    n_halo: int = 0
    person_id_base: int = 0
    halo_gid: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    out_person: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    out_counts: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))  # per destination rank
    in_counts: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))   # per source rank
    # bookkeeping for adapters (not used on the device)
    person_ids: Optional[np.ndarray] = None  # reference Person.id per person index
    group_ids: Optional[np.ndarray] = None   # reference Group.id per group index
    grp_super_area: Optional[np.ndarray] = None  # super area of every group (domain decomposition)
    # leisure (leisure.py:230-275): activity k (reference distributor order) sends people to groups
    # of supergroup `leisure_supergroups[k]`; candidates per (area, activity) as a CSR
    person_area: Optional[np.ndarray] = None      # uint32 [n_people]
    leisure_activities: List[str] = field(default_factory=list)     # distributor specs: "pub", "gym", ...
    leisure_supergroups: List[str] = field(default_factory=list)    # supergroup names: "pubs", "gyms", ...
    leisure_subgroups: List[int] = field(default_factory=list)      # leisure_subgroup_type per activity
    area_venue_off: Optional[np.ndarray] = None   # uint32 [n_areas * K + 1]
    area_venue: Optional[np.ndarray] = None       # uint32 group indices
    # residence visits (residence_visits.py:189-345): the leisure activity "residence_visits" (supergroup name
    # VISITS_SUPERGROUP: it has no venues) sends a person to a residence its household is linked to
    # (Household.residences_to_visit); CSRs over the groups of the households supergroup
    visit_household_off: Optional[np.ndarray] = None   # uint32 [n_households + 1]
    visit_household: Optional[np.ndarray] = None       # uint32 group indices (households)
    visit_care_home_off: Optional[np.ndarray] = None   # uint32 [n_households + 1]
    visit_care_home: Optional[np.ndarray] = None       # uint32 group indices (care homes)
    # commute (travel.py:125-131): -1 | city << 1 | (station << 1) | 1 per person; city -> stations and
    # station -> transports (groups) as CSRs
    person_commute: Optional[np.ndarray] = None          # int32 [n_people]
    city_station_off: Optional[np.ndarray] = None
    city_station: Optional[np.ndarray] = None
    station_transport_off: Optional[np.ndarray] = None
    station_transport: Optional[np.ndarray] = None
    # static facts the individual policies read (include/june_b200.h: JB_PA_*): lockdown status,
    # spec of the primary-activity group, ">= 2 key workers among the household's residents"
    person_attr: Optional[np.ndarray] = None      # uint8 [n_people]
    # region of person.work_super_area when it differs from the home super area, else 0xFF
    # (CloseCompaniesLockdownTiers, individual_policies.py:570-599)
    person_work_region: Optional[np.ndarray] = None  # uint8 [n_people]

    _ARRAYS = ("age", "sex", "care_home_resident", "has_activity", "super_area", "sa_hospital", "sa_region",
               "grp_offset", "grp_info", "grp_aux", "reg_member", "reg_info", "school_years", "beta_scale",
               "halo_gid", "out_person", "out_counts", "in_counts", "person_ids", "group_ids", "grp_super_area",
               "person_area", "area_venue_off", "area_venue", "person_attr", "person_work_region",
               "visit_household_off", "visit_household", "visit_care_home_off", "visit_care_home",
               "person_commute", "city_station_off", "city_station", "station_transport_off", "station_transport")

    @property
    def n_people(self) -> int:
        return int(self.age.shape[0])

    @property
    def n_groups(self) -> int:
        return int(self.grp_info.shape[0])

    @property
    def n_members(self) -> int:
        return int(self.reg_member.shape[0])

    @property
    def n_regions(self) -> int:
        return max(1, len(self.region_names))

    @property
    def n_areas(self) -> int:
        return 0 if self.person_area is None or not len(self.person_area) else int(self.person_area.max()) + 1

    @property
    def has_visits(self) -> bool:
        return VISITS_ACTIVITY in self.leisure_activities and self.visit_household_off is not None

    @property
    def has_leisure(self) -> bool:
        return bool(self.leisure_activities) and self.person_area is not None and self.area_venue_off is not None

    @property
    def has_commute(self) -> bool:
        return self.person_commute is not None and self.station_transport_off is not None

    @property
    def n_super_areas(self) -> int:
        return int(self.sa_region.shape[0])

    def supergroup(self, name: str) -> Supergroup:
        for sg in self.supergroups:
            if sg.name == name:
                return sg
        raise KeyError(name)

    def group_supergroup(self) -> np.ndarray:
        out = np.zeros(self.n_groups, np.int32)
        for k, sg in enumerate(self.supergroups):
            out[sg.first_group: sg.first_group + sg.n_groups] = k
        return out

    def validate(self) -> None:
        n, g, m = self.n_people, self.n_groups, self.n_members
        assert self.age.dtype == np.uint8 and self.sex.dtype == np.uint8
        assert self.care_home_resident.shape == (n,) and self.has_activity.shape == (n,)
        assert self.super_area.dtype == np.uint32 and self.super_area.shape == (n,)
        assert self.grp_offset.dtype == np.uint32 and self.grp_offset.shape == (g + 1,)
        assert self.grp_offset[0] == 0 and self.grp_offset[-1] == m
        assert np.all(np.diff(self.grp_offset.astype(np.int64)) >= 0)
        assert self.grp_info.dtype == np.uint32 and self.grp_aux.dtype == np.uint32 and self.grp_aux.shape == (g,)
        assert self.reg_member.dtype == np.uint32 and self.reg_info.dtype == np.uint16 and self.reg_info.shape == (m,)
        if m:
            assert int(self.reg_member.max()) < n + self.n_halo
        if n:
            assert int(self.super_area.max()) < self.n_super_areas
        exp = 0
        for sg in self.supergroups:
            assert sg.first_group == exp, "supergroups must tile the group range"
            exp += sg.n_groups
        assert exp == g
        if g:
            assert int((self.grp_info & 0xFF).max()) < len(self.beta_rules)
            assert int(((self.grp_info >> 8) & 0xFF).max()) < len(self.beta_scale)
            assert int((self.grp_info >> 16).max()) < self.n_regions
        assert self.halo_gid.shape == (self.n_halo,)
        # members sorted by (lsg, person) inside each group
        if m:
            grp_of = np.repeat(np.arange(g, dtype=np.int64), np.diff(self.grp_offset.astype(np.int64)))
            key = (grp_of << 36) | ((self.reg_info & 0xFF).astype(np.int64) << 28) | self.reg_member.astype(np.int64)
            assert np.all(np.diff(key) > 0), "registered members must be sorted by (group, subgroup, person)"

    # ---- persistence (small fixtures under tests/golden/) ------------------------------------
    def save(self, path: str) -> None:
        meta = {
            "region_names": self.region_names,
            "supergroups": [sg.to_json() for sg in self.supergroups],
            "beta_rules": [b.to_json() for b in self.beta_rules],
            "n_halo": int(self.n_halo),
            "person_id_base": int(self.person_id_base),
            "leisure_activities": list(self.leisure_activities),
            "leisure_supergroups": list(self.leisure_supergroups),
            "leisure_subgroups": [int(x) for x in self.leisure_subgroups],
        }
        arrays = {k: getattr(self, k) for k in self._ARRAYS if getattr(self, k) is not None}
        np.savez_compressed(path, _meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)

    @classmethod
    def load(cls, path: str) -> "WorldSoA":
        z = np.load(path)
        return cls.from_npz(z)

    @classmethod
    def from_npz(cls, z, prefix: str = "") -> "WorldSoA":
        meta = json.loads(bytes(z[prefix + "_meta"]).decode())
        kw: Dict[str, object] = {}
        for k in cls._ARRAYS:
            kw[k] = z[prefix + k] if (prefix + k) in z else None
        for k in ("halo_gid", "out_person"):
            if kw[k] is None:
                kw[k] = np.zeros(0, np.uint32)
        for k in ("out_counts", "in_counts"):
            if kw[k] is None:
                kw[k] = np.zeros(0, np.int64)
        return cls(
            region_names=list(meta["region_names"]),
            supergroups=[Supergroup(**d) for d in meta["supergroups"]],
            beta_rules=[BetaRule(**d) for d in meta["beta_rules"]],
            n_halo=int(meta["n_halo"]),
            person_id_base=int(meta["person_id_base"]),
            leisure_activities=list(meta.get("leisure_activities", [])),
            leisure_supergroups=list(meta.get("leisure_supergroups", [])),
            leisure_subgroups=list(meta.get("leisure_subgroups", [])),
            **kw,
        )
