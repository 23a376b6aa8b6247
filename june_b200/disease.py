"""Disease description: symptom tags, trajectories, transmission sampler and the outcome
(health-index) table, parsed from the reference's disease YAML schema
(``june/configs/defaults/epidemiology/infection/disease/<name>.yaml``)
and flattened into the plain tables the device consumes (``JbDisease`` in june_b200.h).

Mirrors ``DiseaseConfig`` / ``SymptomManager`` / ``RatesManager``
(``june/epidemiology/infection/disease_config.py:65-522``), ``TrajectoryMakers``
(``trajectory_maker.py:252``) and ``HealthIndexGenerator`` (``health_index/health_index.py:21``).
"""
from __future__ import annotations

import zlib
from typing import Dict, List, Optional

import numpy as np
import yaml

from .interaction import load_defaults
from .world import MAX_STAGES

DIST_CODES = {"constant": 0, "exponential": 1, "beta": 2, "lognormal": 3, "normal": 4, "exponweib": 5}
TRANS_CODES = {"constant": 0, "gamma": 1, "xnexp": 2}

_INFECTION_CLASS = {"covid19": "Covid19", "measles": "Measles", "ev-d68-v": "EVD68V"}


def dist_params(d: dict):
    """(type code, p[4]) for one completion-time dict (``trajectory_maker.py:43-132``)."""
    t = d["type"]
    if t == "constant":
        p = [d["value"]]
    elif t == "exponential":
        p = [d["loc"], d["scale"]]
    elif t == "beta":
        p = [d["a"], d["b"], d.get("loc", 0.0), d.get("scale", 1.0)]
    elif t == "lognormal":
        p = [d["s"], d.get("loc", 0.0), d.get("scale", 1.0)]
    elif t == "normal":
        p = [d["loc"], d["scale"]]
    elif t == "exponweib":
        p = [d["a"], d["c"], d.get("loc", 0.0), d.get("scale", 1.0)]
    else:
        raise AssertionError(f"Unrecognised variation type {t}")
    p = [float(x) for x in p] + [0.0] * (4 - len(p))
    return DIST_CODES[t], p


class DiseaseConfig:
    def __init__(self, disease_name: str = "covid19", disease_yaml: Optional[dict] = None):
        self.disease_name = disease_name
        if disease_yaml is None:
            disease_yaml = load_defaults()["disease"][disease_name]
        self.disease_yaml = disease_yaml
        d = disease_yaml["disease"]
        self.symptom_tags: Dict[str, int] = {t["name"]: int(t["value"]) for t in d["symptom_tags"]}
        settings = d.get("settings", {})
        self.default_lowest_stage = settings.get("default_lowest_stage", "exposed")
        self.max_mild_symptom_tag = settings.get("max_mild_symptom_tag")
        self._settings = settings
        self.infection_outcome_rates = [o["parameter"] for o in d.get("infection_outcome_rates", [])]
        self.rate_to_tag_mapping = dict(d.get("rate_to_tag_mapping", {}))
        self.unrated_tags = list(d.get("unrated_tags", []) or [])
        self.trajectories = d["trajectories"]
        self.transmission = d["transmission"]

    @classmethod
    def from_file(cls, path: str) -> "DiseaseConfig":
        with open(path) as f:
            y = yaml.safe_load(f)
        return cls(disease_name=y["disease"].get("name", "disease"), disease_yaml=y)

    def resolve_tags(self, category: str) -> List[int]:
        """``SymptomManager._resolve_tags``: tag values of a settings category."""
        return [self.symptom_tags[e["name"]] for e in self._settings.get(category, []) or []]

    def tag_mask(self, category: str) -> int:
        m = 0
        for v in self.resolve_tags(category):
            m |= 1 << (v + 2)
        return m

    @property
    def n_tags(self) -> int:
        return max(self.symptom_tags.values()) + 1

    @property
    def infection_id(self) -> int:
        # Infection.infection_id(): adler32 of the class name (infection.py:42-47)
        name = _INFECTION_CLASS.get(self.disease_name.lower(), self.disease_name)
        return zlib.adler32(name.encode("ascii"))

    # ---- trajectories ------------------------------------------------------------------------
    def trajectory_table(self):
        """List of ``(max_tag, [(tag, dist_code, params), ...])``; keyed like
        ``TrajectoryMakers.trajectories`` by the most severe tag (later entries win)."""
        out = []
        for tr in self.trajectories:
            stages = []
            for st in tr["stages"]:
                tagname = st["symptom_tag"]
                tag = self.symptom_tags[tagname] if isinstance(tagname, str) else int(tagname)
                code, p = dist_params(dict(st["completion_time"]))
                stages.append((tag, code, p))
            assert len(stages) <= MAX_STAGES
            out.append((max(s[0] for s in stages), stages))
        return out

    def transmission_table(self):
        t = self.transmission
        ttype = t["type"]
        if ttype == "gamma":
            keys = ["max_infectiousness", "shape", "rate", "shift"]
        elif ttype == "xnexp":
            keys = ["max_probability", "smearing_time_first_infectious", "smearing_peak_position", "alpha", "norm_time"]
        elif ttype == "constant":
            keys = ["probability"]
        else:
            raise NotImplementedError(f"Transmission type '{ttype}' is not implemented.")
        return TRANS_CODES[ttype], [dist_params(dict(t[k])) for k in keys]


# ---- outcome table ---------------------------------------------------------------------------
def outcome_probabilities(dc: DiseaseConfig, rates: Dict[str, np.ndarray], age_bins) -> np.ndarray:
    """``HealthIndexGenerator._get_probabilities`` / ``_set_probability_per_age_bin``
    (``health_index.py:184-305,358-400``).

    ``rates[f"{pop}_{parameter}_{male|female}"]`` = one value per age bin; ``age_bins`` = list of
    closed (lo, hi).  Returns ``prob[pop(gp,ch)][sex(f,m)][age 0..99][n_tags]``."""
    n_tags = dc.n_tags
    out = np.zeros((2, 2, 100, n_tags))
    fatal = dc.resolve_tags("fatality_stage")
    lowest = dc.symptom_tags[dc.default_lowest_stage]
    for pi, pop in enumerate(("gp", "ch")):
        for si, sex in enumerate(("female", "male")):
            for bi, (lo, hi) in enumerate(age_bins):
                probs = [0.0] * n_tags
                for par in dc.infection_outcome_rates:
                    if par in dc.rate_to_tag_mapping:
                        probs[dc.symptom_tags[dc.rate_to_tag_mapping[par]]] = float(rates[f"{pop}_{par}_{sex}"][bi])
                for un in dc.unrated_tags:
                    dep = sum(probs[dc.symptom_tags[d]] for d in un["rate_calc_dependency"] if d in dc.symptom_tags)
                    probs[dc.symptom_tags[un["name"]]] = max(0, 1 - dep)
                for i in range(lowest):
                    probs[i] = 0
                fat = sum(probs[i] for i in fatal)
                nonfat = [probs[i] if i not in fatal else 0 for i in range(n_tags)]
                tot = sum(nonfat)
                if tot > 0:
                    nonfat = [v / tot * (1 - fat) if v > 0 else 0 for v in nonfat]
                final = [nonfat[i] if i not in fatal else probs[i] for i in range(n_tags)]
                for age in range(lo, min(hi, 99) + 1):
                    out[pi, si, age, :] = final
    return out


def health_index_cum(prob: np.ndarray) -> np.ndarray:
    """``np.cumsum(probabilities)`` per (pop, sex, age) (``health_index.py:181``)."""
    return np.cumsum(prob, axis=-1)


def rates_from_dataframe(df):
    """pandas rates table in the reference schema -> (rates dict, age bins)."""
    bins = [(int(iv.left), int(iv.right)) for iv in df.index]
    return {c: df[c].to_numpy(dtype=np.float64) for c in df.columns}, bins


def synthetic_rates():
    """Synthetic, age-banded outcome rates in the reference's column schema
    (``disease_config.py:191``; the real CSV needs the absent ``data/`` tree).  The same table is
    used by the golden fixtures (tests/golden/ref_world.py)."""
    bins = [(0, 9), (10, 19), (20, 29), (30, 39), (40, 49), (50, 59), (60, 69), (70, 79), (80, 89), (90, 99)]
    base = {
        "asymptomatic": [0.45, 0.40, 0.35, 0.30, 0.28, 0.25, 0.22, 0.20, 0.18, 0.15],
        "mild": [0.50, 0.52, 0.52, 0.50, 0.45, 0.40, 0.35, 0.28, 0.22, 0.20],
        "hospital": [0.004, 0.006, 0.02, 0.04, 0.06, 0.10, 0.16, 0.22, 0.26, 0.28],
        "icu": [0.0005, 0.001, 0.004, 0.008, 0.015, 0.03, 0.05, 0.06, 0.05, 0.04],
        "home_ifr": [0.0001, 0.0001, 0.0003, 0.0006, 0.001, 0.003, 0.008, 0.02, 0.04, 0.06],
        "hospital_ifr": [0.0002, 0.0003, 0.001, 0.002, 0.005, 0.012, 0.03, 0.07, 0.12, 0.15],
        "icu_ifr": [0.0001, 0.0002, 0.0008, 0.002, 0.004, 0.01, 0.02, 0.03, 0.03, 0.03],
    }
    rates = {}
    for pop, mult in (("gp", 1.0), ("ch", 1.5)):
        for sex, smult in (("male", 1.1), ("female", 0.9)):
            for name, vals in base.items():
                v = np.array(vals, dtype=np.float64)
                if name not in ("asymptomatic", "mild"):
                    v = v * mult * smult
                rates[f"{pop}_{name}_{sex}"] = v
    return rates, bins
