"""Host half of the reference's leisure model (``june/groups/leisure/leisure.py:125-372``,
``social_venue_distributor.py:24-215``): configuration and the per-time-slot probability tables.

The per-person part -- does the person go out, which activity, which of the venues its area
lists (``leisure.py:230-275``, ``social_venue_distributor.py:257-278``) -- runs inside the placement
kernel; this module only produces, per time slot, the two tables the kernel looks up::

    does[region][sex][age]          = 1 - exp(-delta_t * sum_k lambda_k)          (leisure.py:338)
    cum[region][sex][age][k]        = cumulative lambda_k / sum lambda             (leisure.py:342-349)

with the reference's Poisson parameters ``lambda_k = times_per_week / n_days * 24 / hours_per_day``
(``social_venue_distributor.py:143-149``), opening hours (``leisure.py:402-408``), regional
compliance / closed venues (``:183-205``) and leisure-policy reductions (``:410-413``).

Covered: social venues (pubs, cinemas, groceries, gyms: one leisure subgroup per venue).
Not covered yet: residence visits (``residence_visits.py``; their placement has order-dependent side
effects on the visited household) and friend invitations.
"""
from __future__ import annotations

import datetime
from typing import Dict, List, Optional, Sequence

import numpy as np
import yaml

from .interaction import load_defaults

DEFAULT_DAYTYPES = {"weekday": ["Monday", "Tuesday", "Wednesday", "Thursday", "Friday"],
                    "weekend": ["Saturday", "Sunday"]}
_DAY_NAMES = ["Monday", "Tuesday", "Wednesday", "Thursday", "Friday", "Saturday", "Sunday"]
SEXES = ("f", "m")  # device sex byte: 0 = f, 1 = m


def parse_age_probabilities(age_dict: Optional[dict], fill_value: float = 0.0) -> List[float]:
    """``{"a-b": v}`` -> 100 per-age values, ``a <= age < b`` (``utils/parse_probabilities.py:6-34``)."""
    out = [fill_value] * 100
    if not age_dict:
        return out
    for key, v in age_dict.items():
        lo, hi = (int(x) for x in str(key).split("-"))
        for age in range(max(0, lo), min(100, hi)):
            out[age] = v
    return out


def parse_opens(opens: Dict[str, str]) -> Dict[str, List[int]]:
    return {k: [int(x) for x in v.split("-")] for k, v in opens.items()}


class SocialVenueDistributor:
    """Configuration of one leisure activity (``social_venue_distributor.py:24-215``)."""

    spec = "social_venue"
    config_key: Optional[str] = None

    def __init__(self, social_venues=None, times_per_week: Optional[dict] = None, daytypes: dict = DEFAULT_DAYTYPES,
                 hours_per_day: Optional[dict] = None, drags_household_probability: float = 0.0,
                 invites_friends_probability: float = 0.0, neighbours_to_consider: int = 5, maximum_distance: float = 5,
                 leisure_subgroup_type: int = 0, nearest_venues_to_visit: int = 0, open: Optional[dict] = None):
        self.social_venues = social_venues
        self.daytypes = daytypes
        self.drags_household_probability = drags_household_probability
        self.invites_friends_probability = invites_friends_probability
        self.neighbours_to_consider = neighbours_to_consider
        self.maximum_distance = maximum_distance
        self.leisure_subgroup_type = leisure_subgroup_type
        self.nearest_venues_to_visit = nearest_venues_to_visit
        self.open = open or {"weekday": "0-24", "weekend": "0-24"}
        if hours_per_day is None:
            hours_per_day = {d: {"male": {"0-65": 3, "65-100": 11}, "female": {"0-65": 3, "65-100": 11}} if d == "weekday"
                             else {"male": {"0-100": 12}, "female": {"0-100": 12}} for d in ("weekday", "weekend")}
        self.poisson_parameters = self._parse_poisson_parameters(times_per_week or {}, hours_per_day)

    @classmethod
    def from_config(cls, social_venues=None, daytypes: dict = DEFAULT_DAYTYPES, config_filename: Optional[str] = None,
                    config_override: Optional[dict] = None):
        if config_filename is None:
            config = dict(load_defaults()["leisure"][cls.config_key])
        else:
            with open(config_filename) as f:
                config = yaml.safe_load(f)
        for k, v in (config_override or {}).items():
            if v is not None:
                config[k] = v
        return cls(social_venues, daytypes=daytypes, **config)

    def _parse_poisson_parameters(self, times_per_week: dict, hours_per_day: dict):
        ret: Dict[str, Dict[str, List[float]]] = {}
        for day_type in ("weekday", "weekend"):
            ret[day_type] = {}
            ndays = len(self.daytypes[day_type])
            for sex_long, sex in (("male", "m"), ("female", "f")):
                tpw = parse_age_probabilities((times_per_week.get(day_type) or {}).get(sex_long))
                hpd = parse_age_probabilities((hours_per_day.get(day_type) or {}).get(sex_long))
                # social_venue_distributor.py:143-149
                ret[day_type][sex] = [0.0 if t == 0 else (t / ndays) * (24 / h) for t, h in zip(tpw, hpd)]
        return ret

    def get_poisson_parameter(self, sex: str, age: int, day_type: str, working_hours: bool,
                              regional_compliance: Optional[float] = None, closed: bool = False,
                              policy_reduction: Optional[float] = None) -> float:
        """``social_venue_distributor.py:176-205`` (``regional_compliance`` None = no region)."""
        if closed:
            return 0.0
        original = self.poisson_parameters[day_type][sex][age]
        if policy_reduction is None:
            return original
        rc = 1.0 if regional_compliance is None else regional_compliance
        return original * (1 + rc * (policy_reduction - 1))

    def poisson_parameters_by_age(self, sex: str, day_type: str, working_hours: bool,
                                  regional_compliance: Optional[float] = None, closed: bool = False,
                                  policy_reduction=None) -> np.ndarray:
        """``get_poisson_parameter`` for the ages 0..99 at once: the same arithmetic, element by element
        (``policy_reduction``: a value per age, or None).  The tables of a time slot are 9 regions x 2 sexes x 100 ages
        x K activities; the per-age Python calls were the host cost of every new slot."""
        if closed:
            return np.zeros(100)
        original = np.asarray(self.poisson_parameters[day_type][sex], dtype=np.float64)
        if policy_reduction is None:
            return original
        rc = 1.0 if regional_compliance is None else regional_compliance
        red = np.array([policy_reduction[age] for age in range(100)], dtype=np.float64)
        return original * (1 + rc * (red - 1))


class PubDistributor(SocialVenueDistributor):
    spec, config_key = "pub", "pubs"


class CinemaDistributor(SocialVenueDistributor):
    spec, config_key = "cinema", "cinemas"


class GroceryDistributor(SocialVenueDistributor):
    spec, config_key = "grocery", "groceries"


class GymDistributor(SocialVenueDistributor):
    spec, config_key = "gym", "gyms"


class ResidenceVisitsDistributor(SocialVenueDistributor):
    """``ResidenceVisitsDistributor`` (``residence_visits.py:19-50,333-351``): visits between households and to care
    homes.  No venues: the candidates are the residences the visitor's household is linked to
    (``WorldSoA.visit_household`` / ``visit_care_home``).  Nobody visits during working hours."""

    spec, config_key = "residence_visits", "visits"

    def __init__(self, residence_type_probabilities: Optional[dict] = None, times_per_week: Optional[dict] = None,
                 hours_per_day: Optional[dict] = None, daytypes: dict = DEFAULT_DAYTYPES,
                 drags_household_probability: float = 0.0, invites_friends_probability: float = 0.0, social_venues=None):
        self.residence_type_probabilities = dict(residence_type_probabilities or {"household": 0.66, "care_home": 0.34})
        self.policy_reductions: dict = {}   # ChangeVisitsProbability: the probabilities in force while it is active
        super().__init__(None, times_per_week=times_per_week, daytypes=daytypes, hours_per_day=hours_per_day,
                         drags_household_probability=drags_household_probability,
                         invites_friends_probability=invites_friends_probability)

    @classmethod
    def from_config(cls, social_venues=None, daytypes: dict = DEFAULT_DAYTYPES, config_filename: Optional[str] = None,
                    config_override: Optional[dict] = None):
        if config_filename is None:
            config = dict(load_defaults()["leisure"][cls.config_key])
        else:
            with open(config_filename) as f:
                config = yaml.safe_load(f)
        config.update({k: v for k, v in (config_override or {}).items() if v is not None})
        return cls(daytypes=daytypes, **config)

    def get_poisson_parameter(self, sex: str, age: int, day_type: str, working_hours: bool, **kw) -> float:
        if working_hours:
            return 0.0
        return super().get_poisson_parameter(sex=sex, age=age, day_type=day_type, working_hours=working_hours, **kw)

    def poisson_parameters_by_age(self, sex: str, day_type: str, working_hours: bool, **kw) -> np.ndarray:
        if working_hours:
            return np.zeros(100)
        return super().poisson_parameters_by_age(sex=sex, day_type=day_type, working_hours=working_hours, **kw)

    def type_probabilities(self):
        """(household, care_home) as ``get_leisure_group`` reads them (``residence_visits.py:305-309``)."""
        p = self.policy_reductions or self.residence_type_probabilities
        return float(p.get("household", 0.0)), float(p.get("care_home", 0.0))


DISTRIBUTORS = {"pubs": PubDistributor, "cinemas": CinemaDistributor, "groceries": GroceryDistributor,
                "gyms": GymDistributor}


class Leisure:
    """``Leisure`` (``leisure.py:125-372``): ``leisure_distributors`` keeps the reference's insertion
    order (pub, gym, cinema, grocery -- ``generate_leisure_for_world``, ``:33-91``), which is the order
    of the activity axis of the tables and of ``WorldSoA.leisure_activities``."""

    def __init__(self, leisure_distributors: Dict[str, SocialVenueDistributor], region_names: Optional[Sequence[str]] = None):
        self.leisure_distributors = dict(leisure_distributors)
        self.n_activities = len(self.leisure_distributors)
        self.policy_reductions: Dict[str, dict] = {}
        self.region_names = list(region_names or [])
        self.regional_compliance: Dict[str, float] = {}
        self.closed_venues: Dict[str, set] = {}
        self._view_arrays = None

    @property
    def activities(self) -> List[str]:
        return list(self.leisure_distributors)

    def generate_leisure_probabilities_for_timestep(self, delta_time: float, working_hours: bool,
                                                    date: datetime.datetime):
        """Returns ``(does[R][2][100], cum[R][2][100][K])`` (``leisure.py:205-228``); the reference's nested-dict
        view of the same numbers is built on demand (``probabilities_by_region_sex_age``).  The arithmetic is the
        reference's, element by element, carried out on the 100 ages of a (region, sex) at once."""
        regions = self.region_names or [None]
        K = self.n_activities
        does = np.zeros((len(regions), 2, 100))
        cum = np.zeros((len(regions), 2, 100, max(1, K)))
        probs_all = np.zeros((len(regions), 2, 100, max(1, K)))
        day = _DAY_NAMES[date.weekday()]
        per_activity = []  # day type and "open at this hour" do not depend on the person
        for activity, dist in self.leisure_distributors.items():
            day_type = "weekday" if day in dist.daytypes["weekday"] else "weekend"
            ot = parse_opens(dist.open)[day_type]
            is_open = 0 if (ot[1] - ot[0] == 0 or date.hour < ot[0] or date.hour >= ot[1]) else 1
            per_activity.append((activity, dist, day_type, is_open))
        for r, region in enumerate(regions):
            rc = None if region is None else self.regional_compliance.get(region, 1.0)
            for si, sex in enumerate(SEXES):
                lam = []
                for activity, dist, day_type, is_open in per_activity:
                    red = self.policy_reductions[activity][day_type][sex] if activity in self.policy_reductions else None
                    closed = region is not None and dist.spec in self.closed_venues.get(region, ())
                    lam.append(self._poisson_by_age(dist, sex, day_type, working_hours, rc, closed, red) * is_open)
                total = sum(lam) if lam else np.zeros(100)  # 0 + lam[0] + lam[1] + ...: the order of Python's sum
                does[r, si] = 1.0 - np.exp(-delta_time * total)
                if K:
                    with np.errstate(divide="ignore", invalid="ignore"):
                        probs = np.stack([np.where(x == 0, 0.0, x / total) for x in lam], axis=-1)
                    probs_all[r, si, :, :K] = probs
                    cum[r, si, :, :K] = np.cumsum(probs, axis=-1)
        self._view_arrays = (does, probs_all)
        return does, cum

    @staticmethod
    def _poisson_by_age(dist, sex, day_type, working_hours, rc, closed, red) -> np.ndarray:
        """The distributor's Poisson parameters for all ages: its vector method, unless a subclass overrides only the
        per-person ``get_poisson_parameter`` (then that is what counts)."""
        cls = type(dist)
        scalar_owner = next(c for c in cls.__mro__ if "get_poisson_parameter" in vars(c))
        vector_owner = next(c for c in cls.__mro__ if "poisson_parameters_by_age" in vars(c))
        if issubclass(vector_owner, scalar_owner):
            return dist.poisson_parameters_by_age(sex=sex, day_type=day_type, working_hours=working_hours,
                                                  regional_compliance=rc, closed=closed, policy_reduction=red)
        return np.array([dist.get_poisson_parameter(sex=sex, age=age, day_type=day_type, working_hours=working_hours,
                                                    regional_compliance=rc, closed=closed,
                                                    policy_reduction=1 if red is None else red[age])
                         for age in range(100)], dtype=np.float64)

    @property
    def probabilities_by_region_sex_age(self):
        """The reference's view of the last generated tables: ``[region][sex][age] -> {"does_activity",
        "activities": {activity: probability}}`` (without regions: ``[sex][age]``)."""
        if getattr(self, "_view_arrays", None) is None:
            return None
        does, probs = self._view_arrays
        regions = self.region_names or [None]
        names = list(self.leisure_distributors)
        view = {region: {sex: [{"does_activity": does[r, si, age],
                                "activities": dict(zip(names, (float(x) for x in probs[r, si, age, :len(names)])))}
                               for age in range(100)]
                         for si, sex in enumerate(SEXES)}
                for r, region in enumerate(regions)}
        return view if self.region_names else view[None]


def generate_leisure_for_world(list_of_leisure_groups: Sequence[str], world, daytypes: dict = DEFAULT_DAYTYPES) -> Leisure:
    """``generate_leisure_for_world`` (``leisure.py:33-91``) for a ``WorldSoA``: one distributor per
    listed venue supergroup the world has, in the reference's fixed order."""
    dists: Dict[str, SocialVenueDistributor] = {}
    have = {sg.name for sg in world.supergroups}
    for name in ("pubs", "gyms", "cinemas", "groceries"):
        if name in list_of_leisure_groups and name in have:
            d = DISTRIBUTORS[name].from_config(daytypes=daytypes)
            dists[d.spec] = d
    if ("household_visits" in list_of_leisure_groups or "care_home_visits" in list_of_leisure_groups) and getattr(world, "has_visits", False):
        dists["residence_visits"] = ResidenceVisitsDistributor.from_config(daytypes=daytypes)  # leisure.py:84-91
    return Leisure(dists, region_names=world.region_names)
