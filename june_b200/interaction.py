"""Host-side mirror of the reference's ``Interaction`` configuration object.

Same constructor, same YAML schema, same attribute names (``alpha_physical``, ``betas``,
``contact_matrices``, ``beta_reductions``) as ``june/interaction/interaction.py:18-43,370-434``;
the per-group arithmetic (``time_step_for_group``) runs on the GPU.  What stays here is the
tiny per-step scalar work: processed raw contact matrices (once) and the processed beta per
(beta rule, region) (every step).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import yaml

from .world import SCHOOL_DIM, BetaRule

_DEFAULTS = os.path.join(os.path.dirname(__file__), "configs", "defaults.json")


def load_defaults() -> dict:
    with open(_DEFAULTS) as f:
        return json.load(f)


def raw_contact_matrix(contacts, proportion_physical, alpha_physical: float, characteristic_time: float) -> np.ndarray:
    """``InteractiveGroup.get_raw_contact_matrix`` (``groups/group/interactive.py:140-152``):
    ``C * (1 + (alpha-1) * P) * 24 / characteristic_time``."""
    c = np.array(contacts, dtype=np.float64)
    p = np.array(proportion_physical, dtype=np.float64)
    out = c * (1.0 + (alpha_physical - 1.0) * p)
    out = out * (24 / characteristic_time)
    return np.atleast_2d(out)


def raw_school_contact_matrix(contacts, proportion_physical, alpha_physical: float, characteristic_time: float) -> np.ndarray:
    """``InteractiveSchool.get_raw_contact_matrix`` (``groups/school.py:424-460``): 32x32 table,
    index 0 = teachers, 1+k = pupils aged k (k = 0..30); pupil-pupil contacts decay as
    ``xi ** |age difference|`` with xi hard-coded to 0.3 (the YAML ``xi`` key is ignored)."""
    xi = 0.3
    n = SCHOOL_DIM
    ages = np.arange(0, n - 1)
    diff = np.abs(ages[:, None] - ages[None, :])
    cm = np.zeros((n, n))
    cm[0, 0] = contacts[0][0]
    cm[0, 1:] = contacts[0][1]
    cm[1:, 0] = contacts[1][0]
    cm[1:, 1:] = xi ** diff * contacts[1][1]
    ph = np.zeros((n, n))
    ph[0, 0] = proportion_physical[0][0]
    ph[0, 1:] = proportion_physical[0][1]
    ph[1:, 0] = proportion_physical[1][0]
    ph[1:, 1:] = proportion_physical[1][1]
    cm = cm * (1.0 + (alpha_physical - 1.0) * ph)
    cm = cm * (24 / characteristic_time)
    return cm


def processed_school_matrix(raw: np.ndarray, years: Sequence[int]) -> np.ndarray:
    """``InteractiveSchool.get_processed_contact_matrix`` (``groups/school.py:462-485``).  Host
    copy for tests; the device evaluates the same expression element by element."""
    ny = len(years)
    n = ny + 1
    out = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i == j:
                out[i, j] = raw[1, 1] if i != 0 else raw[0, 0]
            elif i == 0:
                out[i, j] = raw[0, 1] / ny
            elif j == 0:
                out[i, j] = raw[1, 0] / ny
            else:
                yi, yj = years[i - 1] + 1, years[j - 1] + 1
                out[i, j] = raw[yi, yj] / 4 if yi == yj else raw[yi, yj]
    return out


class Interaction:
    def __init__(self, alpha_physical: float, betas: Dict[str, float], contact_matrices: Optional[dict]):
        self.alpha_physical = alpha_physical
        self.betas = dict(betas or {})
        self.contact_matrices = self.get_raw_contact_matrices(
            groups=self.betas.keys(), input_contact_matrices=contact_matrices or {}, alpha_physical=alpha_physical)
        self.beta_reductions: Dict[str, float] = {}
        self.current_time = None

    @classmethod
    def from_file(cls, config_filename: Optional[str] = None, disease: str = "covid19") -> "Interaction":
        """``config_filename``: a reference ``interaction_<disease>.yaml``; default: built-in copy
        of the reference defaults."""
        if config_filename is None:
            cfg = load_defaults()["interaction"][disease]
        else:
            with open(config_filename) as f:
                cfg = yaml.safe_load(f)
        return cls(alpha_physical=cfg["alpha_physical"], betas=cfg["betas"], contact_matrices=cfg.get("contact_matrices", {}))

    def get_raw_contact_matrices(self, groups, input_contact_matrices: dict, alpha_physical: float) -> Dict[str, np.ndarray]:
        # interaction.py:392-434 (defaults [[1]], [[0]], 8 h)
        out = {}
        for group in groups:
            data = input_contact_matrices.get(group, {}) or {}
            contacts = data.get("contacts", [[1]])
            phys = data.get("proportion_physical", [[0]])
            ctime = data.get("characteristic_time", 8)
            fn = raw_school_contact_matrix if group == "school" else raw_contact_matrix
            out[group] = fn(contacts, phys, alpha_physical, ctime)
        return out

    # ---- per-step scalars -------------------------------------------------------------------
    def processed_beta(self, rule: BetaRule, regional_compliance: float = 1.0, lockdown_tier=None) -> float:
        """``get_processed_beta`` for one (rule, region): ``interactive.py:154-177``,
        ``household.py:256-309`` (no tier term), ``school.py:487-522`` (key fallback)."""
        key = rule.key
        if key in self.betas:
            beta = self.betas[key]
        elif rule.fallback is not None:
            beta = self.betas[rule.fallback]
        else:
            raise KeyError(f"no beta for '{key}'")
        red = self.beta_reductions
        if key in red:
            beta_reduction = red[key]
        elif rule.fallback is not None:
            beta_reduction = red.get(rule.fallback, 1.0)
        else:
            beta_reduction = 1.0
        if not rule.uses_tier:
            return beta * (1 + regional_compliance * (beta_reduction - 1))
        tier = 1 if lockdown_tier is None else lockdown_tier
        tier_reduction = 0.5 if int(tier) == 4 else 1.0
        return beta * (1 + regional_compliance * tier_reduction * (beta_reduction - 1))

    def beta_table(self, rules: List[BetaRule], regional_compliances: Sequence[float], lockdown_tiers: Sequence) -> np.ndarray:
        """[n_rules, n_regions] table uploaded with every step (``JbStepParams.beta``)."""
        out = np.zeros((len(rules), len(regional_compliances)), np.float64)
        for i, rule in enumerate(rules):
            has = rule.key in self.betas or (rule.fallback is not None and rule.fallback in self.betas)
            for r, (rc, tier) in enumerate(zip(regional_compliances, lockdown_tiers)):
                out[i, r] = self.processed_beta(rule, rc, tier) if has else 0.0
        return out
