"""Host half of the reference's vaccination model (``june/epidemiology/vaccines/vaccines.py:380-617``,
``vaccination_campaign.py:36-385``): configuration objects with the reference's constructor
arguments and YAML schema, flattened into the two small tables the daily device kernel reads
(``JbVaccine`` / ``JbVaccinationCampaign`` in ``include/june_b200.h``).

The per-person work -- eligibility, the daily Bernoulli draw ``c / (T - d c)``, the piecewise-linear
dose efficacy and ``susceptibility = min(prior, 1 - efficacy)`` -- runs on the device once per
simulated day (``k_vaccinate``); the symptomatic efficacy feeds the outcome table of new infections
(``HealthIndexGenerator.apply_effective_multiplier``, ``health_index.py:183-204``)."""
from __future__ import annotations

import datetime
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from .leisure import parse_age_probabilities
from .policy import read_date

GROUP_BY = {"age": 0, "residence": 1, "primary_activity": 2, "sex": 3}
PRIMARY_KIND = {"school": 1, "company": 2, "university": 3}


class Vaccine:
    """``Vaccine`` (``vaccines.py:380-417``).  Efficacies: one dict per dose,
    ``{infection name: {"a-b": value}}``."""

    def __init__(self, name: str, days_administered_to_effective: List[int], days_effective_to_waning: List[int],
                 days_waning: List[int], sterilisation_efficacies, symptomatic_efficacies,
                 waning_factor: Optional[float] = 1.0):
        self.name = name
        self.days_administered_to_effective = list(days_administered_to_effective)
        self.days_effective_to_waning = list(days_effective_to_waning)
        self.days_waning = list(days_waning)
        self.sterilisation_efficacies = [{k: parse_age_probabilities(v) for k, v in d.items()} for d in sterilisation_efficacies]
        self.symptomatic_efficacies = [{k: parse_age_probabilities(v) for k, v in d.items()} for d in symptomatic_efficacies]
        self.waning_factor = 1.0 if waning_factor is None else waning_factor

    @classmethod
    def from_config_dict(cls, name: str, config: Dict):
        return cls(name=name, **{k: config[k] for k in (
            "days_administered_to_effective", "days_effective_to_waning", "days_waning", "sterilisation_efficacies",
            "symptomatic_efficacies", "waning_factor")})

    def table(self, which: str, variant_names: Sequence[str]) -> np.ndarray:
        src = self.sterilisation_efficacies if which == "sterilisation" else self.symptomatic_efficacies
        out = np.zeros((len(src), len(variant_names), 100))
        for d, per in enumerate(src):
            for v, name in enumerate(variant_names):
                if name in per:
                    out[d, v] = per[name]
        return out


class VaccinationCampaign:
    """``VaccinationCampaign`` (``vaccination_campaign.py:36-119``)."""

    def __init__(self, vaccine: Vaccine, days_to_next_dose: List[int], dose_numbers: List[int] = (0, 1),
                 start_time: str = "2100-01-01", end_time: str = "2100-01-02", group_by: str = "age",
                 group_type: str = "50-100", group_coverage: float = 1.0, last_dose_type: Optional[Sequence[str]] = None):
        self.vaccine, self.days_to_next_dose, self.dose_numbers = vaccine, list(days_to_next_dose), list(dose_numbers)
        self.start_time, self.end_time = read_date(start_time), read_date(end_time)
        self.group_by, self.group_type, self.group_coverage = group_by, group_type, group_coverage
        self.last_dose_type = list(last_dose_type) if last_dose_type else []
        self.total_days = (self.end_time - self.start_time).days

    def is_active(self, date: datetime.datetime) -> bool:
        return self.start_time <= date < self.end_time

    def daily_vaccination_probability(self, days_passed: int) -> float:
        return self.group_coverage * (1 / (self.total_days - days_passed * self.group_coverage))


class VaccinationCampaigns:
    """``VaccinationCampaigns`` (``vaccination_campaign.py:266-385``)."""

    def __init__(self, vaccination_campaigns: List[VaccinationCampaign]):
        self.vaccination_campaigns = list(vaccination_campaigns)

    def __iter__(self):
        return iter(self.vaccination_campaigns)

    def get_active(self, date: datetime.datetime) -> List[VaccinationCampaign]:
        return [vc for vc in self.vaccination_campaigns if vc.is_active(date)]

    def vaccines(self) -> List[Vaccine]:
        out: List[Vaccine] = []
        for vc in self.vaccination_campaigns:
            if vc.vaccine not in out:
                out.append(vc.vaccine)
        return out

    def device_tables(self, first_day: datetime.date, variant_names: Sequence[str]):
        """``(vaccine dicts, campaign dicts)`` for ``Engine.set_vaccination``; days are counted from
        ``first_day`` (the date of the simulation's first step)."""
        vaccines: List[Vaccine] = []
        for vc in self.vaccination_campaigns:
            if vc.vaccine not in vaccines:
                vaccines.append(vc.vaccine)
        vrec = [dict(n_doses=len(v.sterilisation_efficacies), days_administered_to_effective=v.days_administered_to_effective,
                     days_effective_to_waning=v.days_effective_to_waning, days_waning=v.days_waning,
                     waning_factor=v.waning_factor, sterilisation=v.table("sterilisation", variant_names),
                     symptomatic=v.table("symptomatic", variant_names)) for v in vaccines]
        names = [v.name for v in vaccines]
        crec = []
        for vc in self.vaccination_campaigns:
            if vc.group_by == "age":
                lo, hi = (int(x) for x in vc.group_type.split("-"))
            elif vc.group_by == "residence":
                lo, hi = (1 if vc.group_type in ("care_home", "carehome") else 0), 0
            elif vc.group_by == "primary_activity":
                lo, hi = PRIMARY_KIND.get(vc.group_type, 4), 0
            elif vc.group_by == "sex":
                lo, hi = (1 if vc.group_type == "m" else 0), 0
            else:
                raise ValueError(f"vaccination group_by={vc.group_by!r} is not supported on the device")
            mask = 0
            for t in vc.last_dose_type:
                if t in names:
                    mask |= 1 << names.index(t)
            crec.append(dict(vaccine=vaccines.index(vc.vaccine), group_by=GROUP_BY[vc.group_by], lo=lo, hi=hi,
                             dose_numbers=vc.dose_numbers, days_to_next_dose=vc.days_to_next_dose,
                             start_day=(vc.start_time.date() - first_day).days, end_day=(vc.end_time.date() - first_day).days,
                             last_dose_vaccines=mask, coverage=vc.group_coverage))
        return vrec, crec


def campaigns_from_reference(ref_campaigns, infection_names: Dict[int, str]) -> VaccinationCampaigns:
    """The reference's own ``VaccinationCampaigns`` / ``VaccinationCampaign`` / ``Vaccine`` objects
    (``vaccination_campaign.py:36-119,266-300``, ``vaccines.py:380-513``) as this module's classes.  The
    reference keys efficacies by infection id (``_parse_efficacies``); ``infection_names`` maps them back to
    the class names the device tables are ordered by."""
    vaccines: Dict[int, Vaccine] = {}
    out = []
    for rc in ref_campaigns.vaccination_campaigns:
        rv = rc.vaccine
        if id(rv) not in vaccines:
            v = Vaccine.__new__(Vaccine)
            v.name = rv.name
            v.days_administered_to_effective = list(rv.days_administered_to_effective)
            v.days_effective_to_waning = list(rv.days_effective_to_waning)
            v.days_waning = list(rv.days_waning)
            v.waning_factor = 1.0 if rv.waning_factor is None else rv.waning_factor
            conv = lambda doses: [{infection_names[k]: list(x) for k, x in d.items() if k in infection_names} for d in doses]
            v.sterilisation_efficacies, v.symptomatic_efficacies = conv(rv.sterilisation_efficacies), conv(rv.symptomatic_efficacies)
            vaccines[id(rv)] = v
        attr = rc.group_attribute
        group_by = "residence" if attr.startswith("residence") else "primary_activity" if attr.startswith("primary_activity") else attr
        c = VaccinationCampaign.__new__(VaccinationCampaign)
        c.vaccine, c.days_to_next_dose, c.dose_numbers = vaccines[id(rv)], list(rc.days_to_next_dose), list(rc.dose_numbers)
        c.start_time, c.end_time = rc.start_time, rc.end_time
        c.group_by, c.group_type, c.group_coverage = group_by, rc.group_value, rc.group_coverage
        c.last_dose_type = list(rc.last_dose_type or [])
        c.total_days = (c.end_time - c.start_time).days
        out.append(c)
    return VaccinationCampaigns(out)
