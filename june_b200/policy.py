"""The policies ``Simulator.do_timestep`` evaluates on every step, with the reference's class
names, constructor arguments and YAML schema (``june/policy/*.py``,
``configs/defaults/policy/policy_<disease>.yaml``):

* interaction policies -> ``interaction.beta_reductions`` (``interaction_policies.py:15-96``),
* regional compliance / tiered lockdown -> per-region scalars (``regional_compliance.py:21-66``),
* ``SevereSymptomsStayHome`` (``individual_policies.py:121-149``) and ``Hospitalisation``
  (``medical_care_policies.py:413-526``) -> step flags evaluated per person on the device,
* ``Quarantine``, ``Shielding``, ``CloseSchools``, ``CloseUniversities``, ``CloseCompanies``
  (``individual_policies.py:152-237,416-548,604-770``) -> one small record each; the placement
  kernel evaluates the active records per person, in the reference's list order.

Per-person work never happens here: a policy contributes either a few host scalars per step or a
flag bit that switches a branch inside the placement / health kernels.
"""
from __future__ import annotations

import datetime
import re
from collections import defaultdict
from typing import Dict, List, Optional, Sequence, Union

import yaml

from . import _lib
from .interaction import load_defaults

DateLike = Union[str, datetime.date, datetime.datetime]


def read_date(date: DateLike) -> datetime.datetime:
    if isinstance(date, datetime.datetime):
        return date
    if isinstance(date, datetime.date):
        return datetime.datetime(date.year, date.month, date.day)
    if isinstance(date, str):
        return datetime.datetime.strptime(date.split(" ")[0], "%Y-%m-%d")
    raise TypeError("date must be a string or a datetime.date object")


class Policy:
    policy_type = "policy"

    def __init__(self, start_time: DateLike = "1900-01-01", end_time: DateLike = "2100-01-01"):
        self.spec = re.sub(r"(?<!^)(?=[A-Z])", "_", self.__class__.__name__).lower()
        self.start_time = read_date(start_time)
        self.end_time = read_date(end_time)

    def is_active(self, date: datetime.datetime) -> bool:
        return self.start_time <= date < self.end_time


# ---- interaction policies -----------------------------------------------------------------------
class SocialDistancing(Policy):
    policy_type = "interaction"

    def __init__(self, start_time: DateLike, end_time: DateLike, beta_factors: Optional[dict] = None):
        super().__init__(start_time, end_time)
        self.beta_factors = beta_factors or {}

    def apply(self) -> Dict[str, float]:
        return self.beta_factors


class MaskWearing(Policy):
    policy_type = "interaction"

    def __init__(self, start_time: DateLike, end_time: DateLike, compliance: float, beta_factor: float,
                 mask_probabilities: Optional[dict] = None):
        super().__init__(start_time, end_time)
        self.compliance, self.beta_factor, self.mask_probabilities = compliance, beta_factor, mask_probabilities or {}

    def apply(self) -> Dict[str, float]:
        # interaction_policies.py:93-96
        return {k: 1 - (v * self.compliance * (1 - self.beta_factor)) for k, v in self.mask_probabilities.items()}


# ---- region-level policies ------------------------------------------------------------------------
class RegionalCompliance(Policy):
    policy_type = "regional_compliance"

    def __init__(self, start_time: DateLike, end_time: DateLike, compliances_per_region: dict):
        super().__init__(start_time, end_time)
        self.compliances_per_region = compliances_per_region


class TieredLockdown(Policy):
    policy_type = "tiered_lockdown"

    def __init__(self, start_time: DateLike, end_time: DateLike, tiers_per_region: dict):
        super().__init__(start_time, end_time)
        self.tiers_per_region = tiers_per_region


# ---- per-person policies evaluated on the device ---------------------------------------------------
class SevereSymptomsStayHome(Policy):
    policy_type = "individual"
    step_flag = _lib.STEP_SEVERE_STAY_HOME


class Hospitalisation(Policy):
    policy_type = "medical_care"
    step_flag = _lib.STEP_HOSPITALISATION

    def __init__(self, disease_config=None, start_time: DateLike = "1900-01-01", end_time: DateLike = "2100-01-01"):
        super().__init__(start_time, end_time)
        self.disease_config = disease_config


class Quarantine(Policy):
    """``individual_policies.py:152-237``.  What reaches the device is what the reference actually
    evaluates: ``infection.time_of_symptoms_onset`` is 0 for every infection (``symptoms.py:77-88``
    stops at the first trajectory entry), hence ``release_day == n_days`` and the household branch
    (``household.py:175-185``) never fires -- ``n_days_household`` and the household compliances are
    accepted and unused, exactly as in the reference."""
    policy_type = "individual"

    def __init__(self, start_time: DateLike = "1900-01-01", end_time: DateLike = "2100-01-01", n_days: int = 7,
                 n_days_household: int = 14, compliance: float = 1.0, household_compliance: float = 1.0,
                 vaccinated_household_compliance: float = 1.0):
        super().__init__(start_time, end_time)
        self.n_days, self.n_days_household, self.compliance = n_days, n_days_household, compliance
        self.household_compliance = household_compliance
        self.vaccinated_household_compliance = vaccinated_household_compliance

    def device_record(self, ctx) -> dict:
        return dict(type=_lib.POL_QUARANTINE, mask=ctx["stay_at_home_mask"], p=[self.n_days, self.compliance])


class Shielding(Policy):
    """``individual_policies.py:416-439``."""
    policy_type = "individual"

    def __init__(self, start_time: DateLike, end_time: DateLike, min_age: int, compliance: Optional[float] = None):
        super().__init__(start_time, end_time)
        self.min_age, self.compliance = min_age, compliance

    def device_record(self, ctx) -> dict:
        return dict(type=_lib.POL_SHIELDING, p=[self.min_age, self.compliance])


class CloseSchools(Policy):
    """``individual_policies.py:480-530``."""
    policy_type = "individual"

    def __init__(self, start_time: DateLike, end_time: DateLike, years_to_close=None, attending_compliance: float = 1.0,
                 full_closure=None):
        super().__init__(start_time, end_time)
        self.full_closure, self.attending_compliance = full_closure, attending_compliance
        self.years_to_close = list(range(20)) if years_to_close == "all" else years_to_close

    def device_record(self, ctx) -> dict:
        mask = 0
        for y in self.years_to_close or []:
            if 0 <= int(y) < 32:
                mask |= 1 << int(y)
        return dict(type=_lib.POL_CLOSE_SCHOOLS, flags=1 if self.full_closure else 0, mask=mask,
                    p=[self.attending_compliance])


class CloseUniversities(Policy):
    """``individual_policies.py:533-548``."""
    policy_type = "individual"

    def device_record(self, ctx) -> dict:
        return dict(type=_lib.POL_CLOSE_UNIVERSITIES)


class CloseCompanies(Policy):
    """``individual_policies.py:604-770``.  ``initialize`` reproduces the reference's ratio
    arithmetic literally, including the running normalisation of ``:641-643``."""
    policy_type = "individual"
    furlough_ratio = key_ratio = random_ratio = None

    def __init__(self, start_time: DateLike, end_time: DateLike, full_closure: bool = False, avoid_work_probability=None,
                 furlough_probability=None, key_probability=None):
        super().__init__(start_time, end_time)
        self.full_closure, self.avoid_work_probability = full_closure, avoid_work_probability
        self.furlough_probability, self.key_probability = furlough_probability, key_probability

    @classmethod
    def initialize(cls, world, date=None, record=None):
        """``world``: a ``WorldSoA`` with ``person_attr`` (lockdown status in bits 0-1)."""
        import numpy as np
        status = (world.person_attr & 3) if getattr(world, "person_attr", None) is not None else np.zeros(0, np.uint8)
        f, k, r = (float(np.count_nonzero(status == c)) for c in (3, 1, 2))
        if f != 0 and k != 0 and r != 0:
            f /= f + k + r
            k /= f + k + r
            r /= f + k + r
        else:
            f = k = r = None
        cls.furlough_ratio, cls.key_ratio, cls.random_ratio = f, k, r

    def device_record(self, ctx) -> dict:
        return dict(type=_lib.POL_CLOSE_COMPANIES, flags=1 if self.full_closure else 0,
                    p=[self.avoid_work_probability, self.furlough_probability, self.key_probability,
                       self.furlough_ratio, self.key_ratio, self.random_ratio])


class CloseCompaniesLockdownTiers(Policy):
    """``CloseCompaniesLockdownTiers`` (``individual_policies.py:551-601``): company workers with lockdown
    status "random" who work outside their home super area skip work (and the commute) with probability
    ``regional_compliance`` of the home region when the work region or the home region is in lockdown tier 3
    or 4 (``Region.policy["lockdown_tier"]``, set by ``TieredLockdown``)."""
    policy_type = "individual"
    TIERS = (3, 4)

    def __init__(self, start_time: DateLike = "1900-01-01", end_time: DateLike = "2100-01-01"):
        super().__init__(start_time, end_time)

    def device_record(self, ctx) -> dict:
        mask = 0
        for r, tier in enumerate(ctx.get("lockdown_tiers") or []):
            if tier in self.TIERS:
                mask |= 1 << r
        return dict(type=_lib.POL_CLOSE_COMPANIES_TIERS, mask=mask)


class LimitLongCommute(Policy):
    """``individual_policies.py:773-824``.  The set of long-distance commuters is a static person
    attribute (``JB_PA_LONG_COMMUTER``, computed by the world compiler from the home-area / work-super-area
    distance).  As in the reference, the activity is skipped when ``random() < going_to_work_probability``."""
    policy_type = "individual"

    def __init__(self, start_time: DateLike = "1000-01-01", end_time: DateLike = "9999-12-31", apply_from_distance: float = 150,
                 going_to_work_probability: float = 0.2):
        super().__init__(start_time, end_time)
        self.apply_from_distance, self.going_to_work_probability = apply_from_distance, going_to_work_probability

    def device_record(self, ctx) -> dict:
        return dict(type=_lib.POL_LIMIT_LONG_COMMUTE, p=[self.going_to_work_probability])


class CloseLeisureVenue(Policy):
    """``CloseLeisureVenue`` (``leisure_policies.py:69-100``): the listed venue types are closed in every
    region (``region.policy["global_closed_venues"]``); a closed venue's Poisson rate is 0
    (``social_venue_distributor.py:195-196``).  Names are the reference's: the venue *spec* ("pub",
    "cinema", ...), which is what ``get_poisson_parameter`` compares against."""

    policy_type = "leisure"
    policy_subtype = "close_venues"

    def __init__(self, start_time: DateLike = "1900-01-01", end_time: DateLike = "2100-01-01",
                 venues_to_close: Sequence[str] = ("cinemas", "groceries")):
        super().__init__(start_time, end_time)
        self.venues_to_close = tuple(venues_to_close)

    def apply(self, leisure) -> None:
        for region in leisure.region_names:
            leisure.closed_venues.setdefault(region, set()).update(self.venues_to_close)


class ChangeLeisureProbability(Policy):
    """``ChangeLeisureProbability`` (``leisure_policies.py:103-180``): per activity, day type, sex and age
    a factor on the Poisson rate, softened by the regional compliance
    (``rate * (1 + compliance * (factor - 1))``, ``social_venue_distributor.py:199-204``)."""

    policy_type = "leisure"
    policy_subtype = "change_leisure_probability"

    def __init__(self, start_time: DateLike = "1900-01-01", end_time: DateLike = "2100-01-01",
                 activity_reductions: Optional[Dict[str, dict]] = None):
        super().__init__(start_time, end_time)
        self.activity_reductions = self._read_activity_reductions(activity_reductions or {})

    @staticmethod
    def _read_activity_reductions(activity_reductions: Dict[str, dict]) -> Dict[str, dict]:
        from .leisure import parse_age_probabilities
        sexes = (("male", "m"), ("female", "f"))
        ret: Dict[str, dict] = {}
        for activity, pp in activity_reductions.items():
            out = ret[activity] = {"weekday": {}, "weekend": {}}
            for first in pp:
                if first in ("weekday", "weekend"):
                    for long, short in sexes:
                        src = pp[first]["both_sexes"] if "both_sexes" in pp[first] else pp[first][long]
                        out[first][short] = parse_age_probabilities(src)
                elif first in ("any", "male", "female"):
                    for long, short in sexes:
                        for day_type in out:
                            out[day_type][short] = parse_age_probabilities(pp[long])
                elif first == "both_sexes":
                    for _, short in sexes:
                        for day_type in out:
                            out[day_type][short] = parse_age_probabilities(pp["both_sexes"])
                else:
                    for day_type in out:
                        for long, short in sexes:
                            out[day_type][short] = parse_age_probabilities(pp[day_type][long])
        return ret

    def apply(self, leisure) -> Dict[str, dict]:
        return self.activity_reductions


class ChangeVisitsProbability(Policy):
    """``ChangeVisitsProbability`` (``leisure_policies.py:180-220``): while active, the split between visits to
    households and to care homes (``residence_type_probabilities``) is ``new_residence_type_probabilities``."""

    policy_type = "leisure"
    policy_subtype = "change_visits_probability"

    def __init__(self, start_time: DateLike = "1900-01-01", end_time: DateLike = "2100-01-01",
                 new_residence_type_probabilities: Optional[Dict[str, float]] = None):
        super().__init__(start_time, end_time)
        self.policy_reductions = dict(new_residence_type_probabilities or {})

    def apply(self, leisure) -> None:
        dist = leisure.leisure_distributors.get("residence_visits")
        if dist is not None:
            dist.policy_reductions = self.policy_reductions


_CLASSES = {c.__name__: c for c in (CloseLeisureVenue, ChangeLeisureProbability, ChangeVisitsProbability, SocialDistancing, MaskWearing, RegionalCompliance, TieredLockdown,
                                    SevereSymptomsStayHome, Hospitalisation, Quarantine, Shielding, CloseSchools,
                                    CloseUniversities, CloseCompanies, CloseCompaniesLockdownTiers, LimitLongCommute)}


def _camel(key: str) -> str:
    return "".join(x.capitalize() for x in key.split("_"))


class Policies:
    """Collection with the reference's category views (``policy/policy.py:56-99``)."""

    def __init__(self, policies: Optional[Sequence[Policy]] = None, apply_tiered_lockdown: bool = False,
                 not_evaluated: Optional[Sequence[tuple]] = None):
        self.policies: List[Policy] = list(policies or [])
        self._edges: Optional[List[datetime.datetime]] = None
        # False = what the running reference does (see ``regional_scalars``)
        self.apply_tiered_lockdown = apply_tiered_lockdown
        # (key, start, end) of the policies of a configuration file this package does not evaluate
        self.not_evaluated: List[tuple] = list(not_evaluated or [])

    def check_window(self, start: datetime.datetime, end: datetime.datetime) -> None:
        """Refuse a run during which a policy this package does not evaluate would be active: its effect would
        silently be missing from the results (the plugin raises for the same reason)."""
        hit = sorted({k for k, s, e in self.not_evaluated if s < end and e > start})
        if hit:
            raise NotImplementedError(
                "policies active between %s and %s that june_b200 does not evaluate: %s (remove them from the "
                "policy file, or pass strict=False to Policies.from_file to ignore them knowingly)" % (start, end, ", ".join(hit)))

    def active_key(self, date: datetime.datetime) -> int:
        """Index of the interval between two consecutive policy start / end times that holds
        ``date``: the set of active policies is constant inside an interval, so this integer keys
        everything derived from it (one bisection per step instead of one comparison per policy)."""
        if self._edges is None or self._edges_n != len(self.policies):
            self._edges = sorted({t for p in self.policies for t in (p.start_time, p.end_time)})
            self._edges_n = len(self.policies)
        import bisect
        return bisect.bisect_right(self._edges, date)

    def _of(self, ptype: str) -> List[Policy]:
        return [p for p in self.policies if p.policy_type == ptype]

    @property
    def interaction_policies(self):
        return self._of("interaction")

    @property
    def individual_policies(self):
        return self._of("individual")

    @property
    def leisure_policies(self):
        return self._of("leisure")

    def apply_leisure_policies(self, date: datetime.datetime, leisure) -> None:
        """``LeisurePolicies.apply`` (``leisure_policies.py:23-66``): reset the closed venues and the
        probability reductions, then apply the active policies (at most one change of probabilities)."""
        leisure.closed_venues = {r: set() for r in leisure.region_names}
        leisure.policy_reductions = {}
        if "residence_visits" in getattr(leisure, "leisure_distributors", {}):  # leisure_policies.py:46-48
            leisure.leisure_distributors["residence_visits"].policy_reductions = {}
        if self.apply_tiered_lockdown:
            # TieredLockdown.apply, ``regional_compliance.py:42-52`` (tier 2's ``update("residence_visits")`` adds the
            # string's letters to the set, i.e. closes nothing)
            _, tiers = self.regional_scalars(date, leisure.region_names)
            closures = {3: {"cinema", "residence_visits"}, 4: {"pub", "cinema", "gym", "residence_visits"}}
            for r, t in zip(leisure.region_names, tiers):
                leisure.closed_venues[r] |= closures.get(t, set())
        n_change = 0
        for p in self.leisure_policies:
            if not p.is_active(date):
                continue
            if p.policy_subtype == "change_leisure_probability":
                n_change += 1
                if n_change > 1:
                    raise ValueError("Having more than one change leisure probability policy active is not supported.")
                leisure.policy_reductions = p.apply(leisure=leisure)
            else:
                p.apply(leisure=leisure)

    @property
    def medical_care_policies(self):
        return self._of("medical_care")

    @classmethod
    def from_file(cls, config_file: Optional[str] = None, disease: str = "covid19", disease_config=None,
                  strict: bool = True) -> "Policies":
        """YAML key -> CamelCase class; numbered sub-entries expand to several instances
        (``policy/policy.py:137-229``).  Policies this package does not evaluate (test-and-trace, school
        class quarantines, ...: outside the hot-path scope) are remembered with their dates: ``Simulator`` refuses a
        run window in which one of them would be active (``check_window``) unless ``strict`` is False."""
        if config_file is None:
            cfg = load_defaults()["policy"][disease]
        else:
            with open(config_file) as f:
                cfg = yaml.safe_load(f)
        out: List[Policy] = []
        skipped: List[tuple] = []
        for key, data in (cfg or {}).items():
            klass = _CLASSES.get(_camel(key))
            entries = [data]
            if isinstance(data, dict) and data and all(str(k).isdigit() for k in data):
                entries = [data[k] for k in sorted(data, key=lambda x: int(x))]
            for e in entries:
                kw = dict(e or {})
                if klass is None:
                    if strict:
                        skipped.append((key, read_date(kw.get("start_time", "1900-01-01")), read_date(kw.get("end_time", "2100-01-01"))))
                    continue
                if klass is Hospitalisation:
                    kw["disease_config"] = disease_config
                out.append(klass(**kw))
        return cls(out, not_evaluated=skipped)

    # ---- what do_timestep needs every step -------------------------------------------------------
    def beta_reductions(self, date: datetime.datetime) -> Dict[str, float]:
        """``InteractionPolicies.apply`` (``interaction_policies.py:15-35``)."""
        red: Dict[str, float] = defaultdict(lambda: 1.0)
        for p in self.interaction_policies:
            if p.is_active(date):
                for group, f in p.apply().items():
                    red[group] *= f
        return red

    def regional_scalars(self, date: datetime.datetime, region_names: Sequence[str], state: Optional[dict] = None):
        """``RegionalCompliances.apply`` + ``TieredLockdowns.apply`` (``regional_compliance.py:21-66``):
        returns ``(compliances, lockdown_tiers)`` per region.  ``state`` keeps values between steps
        exactly like the attributes on the reference's ``Region`` objects do."""
        state = state if state is not None else {}
        comp = state.setdefault("compliance", {r: 1.0 for r in region_names})
        tiers = state.setdefault("tier", {r: None for r in region_names})
        rc = self._of("regional_compliance")
        if rc:
            for r in region_names:
                comp[r] = 1.0
        for p in rc:
            if p.is_active(date):
                for r in region_names:
                    comp[r] = p.compliances_per_region[r]
        # The running reference never applies its TieredLockdowns collection: ``Simulator.do_timestep`` calls
        # ``interaction_policies.apply`` and ``regional_compliance.apply`` only (``simulator.py:359-366``), so
        # ``Region.policy["lockdown_tier"]`` stays None for a whole run.  That is the default here too (the drop-in
        # must equal the reference); ``apply_tiered_lockdown = True`` switches the intended behaviour on (tier-4
        # beta reduction, ``CloseCompaniesLockdownTiers``, the tier 3 / 4 venue closures).
        tl = self._of("tiered_lockdown") if self.apply_tiered_lockdown else []
        if tl:
            for r in region_names:
                tiers[r] = None
        for p in tl:
            if p.is_active(date):
                for r in region_names:
                    tiers[r] = int(p.tiers_per_region[r])
        return [comp[r] for r in region_names], [tiers[r] for r in region_names]

    def device_policies(self, date: datetime.datetime, ctx: dict) -> List[dict]:
        """Records of the active individual policies the placement kernel evaluates per person, in
        list order (``IndividualPolicies.get_active``/``apply``, ``individual_policies.py:30-84``).
        ``SevereSymptomsStayHome`` stays a step flag: it draws nothing, so its position is immaterial."""
        return [p.device_record(ctx) for p in self.individual_policies
                if hasattr(p, "device_record") and p.is_active(date)]

    def init_policies(self, world, date=None, record=None) -> None:
        """``Policies.init_policies`` (``policy/policy.py``): class-level set-up such as the
        lockdown-status ratios of ``CloseCompanies``."""
        for p in self.policies:
            if hasattr(type(p), "initialize"):
                type(p).initialize(world, date, record)

    def step_flags(self, date: datetime.datetime) -> int:
        flags = 0
        for p in self.policies:
            if hasattr(p, "step_flag") and p.is_active(date):
                flags |= p.step_flag
        return flags
