"""Drop-in for the reference's time loop: ``Simulator.run`` / ``Simulator.do_timestep``
(``june/simulator.py:346-707``) with the same collaborators -- ``Interaction``, ``Timer``,
``ActivityManager``, ``Epidemiology``, ``Policies`` -- but the world lives in HBM and each
``do_timestep`` is one enqueue of the CUDA kernels behind the C ABI.

Host work per step is O(#regions + #specs): policy scalars, the beta table and one parameter
struct.  Nothing here loops over people or groups.
"""
from __future__ import annotations

import datetime
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import SimulatorError
from .adapter import ACTIVITY_HIERARCHY
from .device import DeviceWorld, Engine
from .disease import DiseaseConfig
from .interaction import Interaction, load_defaults
from .policy import Policies
from .timer import Timer
from .world import ACTIVITY_CODES, WorldSoA


class ActivityManager:
    """Configuration half of the reference ``ActivityManager``
    (``june/activity/activity_manager.py:37-140,172-188``); the per-person loop
    (``move_people_to_active_subgroups``, ``:255-402``) is the placement kernel."""

    def __init__(self, world, policies: Optional[Policies], timer: Timer, all_activities: Sequence[str],
                 activity_to_super_groups: Dict[str, Sequence[str]], record=None, leisure=None, travel=None):
        self.world, self.policies, self.timer = world, policies if policies is not None else Policies([]), timer
        self.all_activities = list(all_activities)
        self.activity_to_super_group_dict = {a: list(activity_to_super_groups.get(a, [])) for a in ACTIVITY_HIERARCHY}
        self.leisure, self.travel = leisure, travel
        for a in self.all_activities:
            if a not in ACTIVITY_CODES:
                raise SimulatorError("Config file contains unsupported activity name.")

    @classmethod
    def from_file(cls, config_filename, world, policies, timer, record=None, leisure=None, travel=None):
        import yaml
        with open(config_filename) as f:
            config = yaml.safe_load(f)
        return cls.from_config(config, world, policies, timer)

    @classmethod
    def from_config(cls, config: dict, world, policies, timer):
        a2sg = config.get("activity_to_super_groups") or config.get("activity_to_groups")  # legacy key, :80-87
        acts = set()
        for day in ("weekday", "weekend"):
            sched = config["time"]["step_activities"][day]
            for a in (sched.values() if isinstance(sched, dict) else sched):
                acts.update(a)
        leisure = None
        if getattr(world, "has_leisure", False):
            from .leisure import generate_leisure_for_world
            leisure = generate_leisure_for_world(a2sg.get("leisure", []), world)
            if leisure.activities != list(world.leisure_activities):
                raise SimulatorError(f"leisure activities of the config {leisure.activities} do not match the "
                                     f"world's {world.leisure_activities}")
        return cls(world=world, policies=policies, timer=timer, all_activities=sorted(acts, key=ACTIVITY_HIERARCHY.index),
                   activity_to_super_groups=a2sg, leisure=leisure)

    @staticmethod
    def apply_activity_hierarchy(activities: List[str]) -> List[str]:
        activities.sort(key=ACTIVITY_HIERARCHY.index)
        return activities

    def activities_to_super_groups(self, activities: Sequence[str]) -> List[str]:
        return [sg for a in activities for sg in self.activity_to_super_group_dict.get(a, [])]

    @property
    def active_super_groups(self) -> List[str]:
        return self.activities_to_super_groups(self.timer.activities)

    def activity_mask(self) -> int:
        acts = self.timer.activities
        cache = self.__dict__.setdefault("_mask_cache", {})
        hit = cache.get(id(acts))
        if hit is None or hit[0] is not acts or hit[1] != len(acts):  # the schedule's own lists: stable objects
            mask = 0
            for a in acts:
                code = ACTIVITY_CODES.get(a)
                if code is not None:
                    mask |= 1 << code
            hit = cache[id(acts)] = (acts, len(acts), mask)
        return hit[2]


class InfectionSeed:
    """Uniform seeding (``InfectionSeed.from_uniform_cases`` + ``unleash_virus_per_day``,
    ``infection_seed.py:146-192,401-475``): on ``date`` a fraction ``cases_per_capita`` of the
    not-yet-infected residents is infected through the device sampler."""

    def __init__(self, cases_per_capita: float, date: str, seed: int = 0):
        self.cases_per_capita = cases_per_capita
        self.date = datetime.datetime.strptime(date.split(" ")[0], "%Y-%m-%d").date()
        self.seed = seed
        self.done = False

    @classmethod
    def from_uniform_cases(cls, world=None, infection_selector=None, cases_per_capita: float = 0.01,
                           date: str = "2020-03-01", seed_past_infections=False, seed: int = 0):
        return cls(cases_per_capita, date, seed)

    def unleash_virus_per_day(self, sim: "Simulator", date: datetime.datetime, time: float) -> int:
        if self.done or date.date() != self.date:
            return 0
        n = sim.world.n_people
        rng = np.random.default_rng([self.seed, sim.rank, 99])
        k = int(round(self.cases_per_capita * n))
        chosen = np.sort(rng.choice(n, size=k, replace=False)).astype(np.uint32)
        sim.device.infect(chosen, time=time, seed=sim.seed ^ 0x5EED, step=0xFFFF0000)
        self.done = True
        if sim.record is not None:
            # the seeds are rows of the `infections` table too (infection_seed.py:196-205, by the person's region):
            # they count in the daily_infected of the step that follows
            w = sim.world
            sim._seeded_by_region = getattr(sim, "_seeded_by_region", 0) + np.bincount(
                np.asarray(w.sa_region)[np.asarray(w.super_area)[chosen]], minlength=w.n_regions)
        return k


class Epidemiology:
    def __init__(self, infection_selectors=None, infection_seeds: Optional[Sequence[InfectionSeed]] = None,
                 immunity_setter=None, medical_care_policies=None, vaccination_campaigns=None):
        self.infection_selectors = infection_selectors
        self.infection_seeds = list(infection_seeds or [])
        self.vaccination_campaigns = vaccination_campaigns
        self.current_date = None

    def infection_seeds_timestep(self, sim: "Simulator", timer: Timer) -> None:
        for s in self.infection_seeds:
            s.unleash_virus_per_day(sim, timer.date, timer.now)


class HaloExchange:
    """The two per-step exchanges of the reference's domain-decomposed mode
    (``activity_manager.py:404-447`` people, ``epidemiology.py:359-401`` infections) as two
    fixed-size ``all_to_all_single`` calls on device buffers (NCCL over NVLink; gloo in CPU tests)."""

    def __init__(self, world: WorldSoA, record_size: int, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.out_counts = [int(x) for x in world.out_counts]
        self.in_counts = [int(x) for x in world.in_counts]
        n_out, n_in = sum(self.out_counts), sum(self.in_counts)
        self.rec = record_size
        dev = device if device is not None else "cpu"
        self.send = torch.zeros(max(1, n_out) * record_size, dtype=torch.uint8, device=dev)
        self.recv = torch.zeros(max(1, n_in) * record_size, dtype=torch.uint8, device=dev)
        self.ret_send = torch.zeros(max(1, n_in), dtype=torch.uint8, device=dev)
        self.ret_recv = torch.zeros(max(1, n_out), dtype=torch.uint8, device=dev)
        self.n_out, self.n_in = n_out, n_in

    def exchange_people(self):
        r = self.rec
        self.dist.all_to_all_single(self.recv[: self.n_in * r], self.send[: self.n_out * r],
                                    output_split_sizes=[c * r for c in self.in_counts],
                                    input_split_sizes=[c * r for c in self.out_counts])

    def exchange_infections(self):
        self.dist.all_to_all_single(self.ret_recv[: self.n_out], self.ret_send[: self.n_in],
                                    output_split_sizes=self.out_counts, input_split_sizes=self.in_counts)


def _read_checkpoint_dates(dates) -> List[datetime.date]:
    """``_read_checkpoint_dates`` (``simulator.py:61-76``): a date, a "YYYY-MM-DD" string or a list of either."""
    if dates is None:
        return []
    if isinstance(dates, (str, datetime.date)):
        dates = [dates]
    out = []
    for d in dates:
        if isinstance(d, str):
            d = datetime.datetime.strptime(d, "%Y-%m-%d").date()
        elif isinstance(d, datetime.datetime):
            d = d.date()
        out.append(d)
    return out


class Simulator:
    ActivityManager = ActivityManager

    def __init__(self, world: WorldSoA, interaction: Interaction, timer: Timer, activity_manager: ActivityManager,
                 epidemiology: Optional[Epidemiology] = None, tracker=None, events=None, record=None,
                 disease_config: Optional[DiseaseConfig] = None, outcome_prob: Optional[np.ndarray] = None,
                 device: int = 0, seed: int = 0, distributed: bool = False,
                 checkpoint_save_dates=None, checkpoint_save_path: Optional[str] = None):
        self.world, self.interaction, self.timer = world, interaction, timer
        self.activity_manager, self.epidemiology, self.record = activity_manager, epidemiology, record
        self.events = events
        if events is not None:
            events.init_events(world=world)
        self.disease_config = disease_config or DiseaseConfig("covid19")
        self.seed = int(seed)
        self.step_ordinal = 0
        self.checkpoint_save_dates = tuple(_read_checkpoint_dates(checkpoint_save_dates))
        self.checkpoint_save_path = checkpoint_save_path or "results/checkpoints"
        self.rank = 0
        self._region_state: dict = {}
        self._policy_cache: dict = {}
        if outcome_prob is None:
            from .disease import outcome_probabilities, synthetic_rates
            rates, bins = synthetic_rates()
            outcome_prob = outcome_probabilities(self.disease_config, rates, bins)
        self.device = self._create_device(world, interaction, outcome_prob, device)
        self.halo: Optional[HaloExchange] = None
        if distributed:
            import torch.distributed as dist
            self.rank = dist.get_rank()
            buffers_on = self._exchange_buffer_device()
            self.halo = HaloExchange(world, self.device.halo_record_size(), device=buffers_on)
        self._ext_stream = None
        self.p2p = False
        self.halo_activity_mask = 0
        if self.halo is not None:
            # activities under which halo persons are registered, OR-ed over the ranks: only steps with one
            # of them exchange anything (a commuter at home is present in no group of another domain)
            import torch
            halo_regs = world.reg_member >= world.n_people
            local = int(np.bitwise_or.reduce(1 << ((world.reg_info[halo_regs].astype(np.int64) >> 8) & 7))) if halo_regs.any() else 0
            masks = [None] * self.halo.dist.get_world_size()
            self.halo.dist.all_gather_object(masks, local)
            self.halo_activity_mask = int(np.bitwise_or.reduce(np.array(masks, np.int64)))
            self.device.set_halo_activities(self.halo_activity_mask)
        if self.halo is not None and buffers_on is not None and os.environ.get("JB_HALO", "p2p") != "nccl":
            self._connect_peer_windows()
        if self.halo is not None and buffers_on is not None:
            import torch
            ptr = self.device._f("stream_handle")(self.device.handle)
            self._ext_stream = torch.cuda.ExternalStream(int(ptr), device=torch.device("cuda", self.device.device_index))
        self._policy_key_on_device = None
        self._vaccination_on_device = False
        self.variant_names = ["Covid19"]  # infection class names, in variant order (vaccines.yaml keys)
        self.activity_manager.policies.init_policies(world)
        self._leisure_tables: dict = {}   # time-slot key -> table index on the device
        self._leisure_stack: List[tuple] = []
        self.infection_events: List[tuple] = []
        self.health_events: List[tuple] = []
        self.vaccine_events: List[tuple] = []
        self.region_history: List[tuple] = []
        self.last_stats: Dict[str, int] = {}
        self.history: List[Dict[str, object]] = []

    def _create_device(self, world, interaction, outcome_prob, device: int) -> Engine:
        """The world on the GPU (``libjune_b200.so``; raises without the library or without a device)."""
        return DeviceWorld(world, interaction, self.disease_config, outcome_prob, device=device)

    def _exchange_buffer_device(self):
        """Where the buffers of the NCCL fallback exchange live: the GPU of this rank."""
        import torch
        return torch.device("cuda", self.device.device_index)

    def _step_with_exchange(self, params, want_stats: bool):
        """A step whose two exchanges are host-issued collectives (``JB_HALO=nccl``): the all-to-alls are
        issued with the library's stream current, so kernels and collectives are ordered on the device
        -- no host synchronisation."""
        dev, h = self.device, self.halo
        with h.torch.cuda.stream(self._ext_stream):
            dev.step_begin(params, h.send.data_ptr())
            h.exchange_people()
            dev.step_interact(h.recv.data_ptr(), h.ret_send.data_ptr())
            h.exchange_infections()
            return dev.step_finish(h.ret_recv.data_ptr(), want_stats=want_stats)

    def _connect_peer_windows(self) -> None:
        """Halo exchange over peer memory (``jb_p2p_window`` / ``jb_p2p_connect``): the ranks swap the
        IPC handles of their exchange windows once; afterwards ``jb_step`` runs the whole halo step as
        one graph whose pack kernels store into the peers' HBM over NVLink -- no host-issued collective.
        ``JB_HALO=nccl`` keeps the ``all_to_all_single`` path."""
        import ctypes as C
        dist = self.halo.dist
        ws, rank = dist.get_world_size(), dist.get_rank()
        dev = self.device
        blob = (C.c_ubyte * 80)()
        dev._check(dev._f("p2p_window")(dev.handle, blob))
        out_counts, in_counts = np.array(self.halo.out_counts, np.uint32), np.array(self.halo.in_counts, np.uint32)
        mine = dict(blob=bytes(blob), in_off=np.concatenate([[0], np.cumsum(in_counts)[:-1]]).tolist(),
                    out_off=np.concatenate([[0], np.cumsum(out_counts)[:-1]]).tolist())
        everyone: List[object] = [None] * ws
        dist.all_gather_object(everyone, mine)
        handles = b"".join(e["blob"][:64] for e in everyone) + b"".join(e["blob"][64:80] for e in everyone)
        hbuf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        recv_base = np.array([everyone[q]["in_off"][rank] for q in range(ws)], np.uint32)
        ret_base = np.array([everyone[q]["out_off"][rank] for q in range(ws)], np.uint32)
        dev._check(dev._f("p2p_connect")(dev.handle, rank, ws, hbuf, out_counts.ctypes.data, in_counts.ctypes.data,
                                         recv_base.ctypes.data, ret_base.ctypes.data))
        dist.barrier()
        self.p2p = True

    @classmethod
    def from_file(cls, world: WorldSoA, interaction: Interaction, policies: Optional[Policies] = None,
                  epidemiology: Optional[Epidemiology] = None, config_filename: Optional[str] = None, **kw):
        if config_filename is None:
            config = load_defaults()["simulation"]
        else:
            import yaml
            with open(config_filename) as f:
                config = yaml.safe_load(f)
        timer = Timer.from_config(config)
        am = cls.ActivityManager.from_config(config, world, policies, timer)
        kw.setdefault("checkpoint_save_dates", config.get("checkpoint_save_dates"))
        return cls(world=world, interaction=interaction, timer=timer, activity_manager=am, epidemiology=epidemiology, **kw)

    def clear_world(self) -> None:
        """``simulator.py:305-343`` empties every subgroup list; here membership is a per-step
        byte per person that the next placement overwrites, so there is nothing to clear."""

    # ---- one step -----------------------------------------------------------------------------------
    def step_params(self, extra_flags: int = 0):
        """Per-step host scalars (``simulator.py:360-366`` + ``get_processed_beta``).  Everything
        but the clock depends only on WHICH policies are active, so the beta table and flags are
        cached per active-policy set."""
        date = self.timer.date
        policies = self.activity_manager.policies
        key = policies.active_key(date)
        hit = self._policy_cache.get(key)
        if hit is None:
            self.interaction.beta_reductions = policies.beta_reductions(date)
            comp, tiers = policies.regional_scalars(date, self.world.region_names, self._region_state)
            beta = np.ascontiguousarray(self.interaction.beta_table(self.world.beta_rules, comp, tiers))
            ctx = dict(stay_at_home_mask=self.disease_config.tag_mask("stay_at_home_stage"), lockdown_tiers=list(tiers))
            hit = (beta, policies.step_flags(date), dict(self.interaction.beta_reductions),
                   policies.device_policies(date, ctx), list(comp))
            self._policy_cache[key] = hit
        beta, flags, self.interaction.beta_reductions, records, comp = hit
        if key != self._policy_key_on_device and self.activity_manager.leisure is not None and getattr(self.world, "has_visits", False):
            # ChangeVisitsProbability: the split between household and care-home visits in force now
            leisure = self.activity_manager.leisure
            policies.apply_leisure_policies(date, leisure)
            probs = leisure.leisure_distributors["residence_visits"].type_probabilities()
            if probs != getattr(self, "_visit_probabilities_on_device", None):
                self.device.set_visit_probabilities(*probs)
                self._visit_probabilities_on_device = probs
        if key != self._policy_key_on_device:
            # the active set changed (a handful of times per run): refresh the device-resident table
            if records or self._policy_key_on_device is not None:
                self.device.set_individual_policies(records, comp)
            self._policy_key_on_device = key
        return self.device.make_params(now=self.timer.now, delta_time=self.timer.duration, step=self.step_ordinal,
                                       activity_mask=self.activity_manager.activity_mask(), seed=self.seed,
                                       flags=flags | extra_flags, beta=beta, leisure_table=self.leisure_table(key))

    def prepare_steps(self, days: int = 7) -> int:
        """Build the device's step graph for every distinct (activities, flags) combination of the next
        ``days`` days under the currently active policy set (``jb_step_prepare``: capture + instantiate,
        nothing runs), so that no ``do_timestep`` pays for a graph capture.  Returns the number of
        distinct step shapes."""
        import copy
        saved, saved_am = self.timer, self.activity_manager.timer
        policies = self.activity_manager.policies
        key0 = policies.active_key(saved.date)
        extra = _lib.STEP_RECORD_EVENTS if self.record is not None else 0
        seen = set()
        try:
            self.timer = self.activity_manager.timer = copy.deepcopy(saved)
            end = self.timer.date + datetime.timedelta(days=days)
            while self.timer.date < end and self.timer.date < self.timer.final_date:
                if self.timer.activities and policies.active_key(self.timer.date) == key0:
                    params = self.step_params(extra)
                    sig = (int(params.activity_mask), int(params.flags), bool(params.leisure_table))
                    if sig not in seen:
                        seen.add(sig)
                        self.device.step_prepare(params)
                next(self.timer)
        finally:
            self.timer, self.activity_manager.timer = saved, saved_am
        return len(seen)

    def prepare_leisure_tables(self, days: int = 7) -> int:
        """Compute and upload the leisure tables of every time slot of the next ``days`` days in one
        go (otherwise each new slot is appended -- and the captured step graphs dropped -- the first
        time the clock reaches it).  Returns the number of distinct slots."""
        if self.activity_manager.leisure is None or not self.world.has_leisure:
            return 0
        import copy
        saved = self.timer
        policies = self.activity_manager.policies
        try:
            self.timer = copy.deepcopy(saved)
            end = self.timer.date + datetime.timedelta(days=days)
            upload, self.device.set_leisure_tables = self.device.set_leisure_tables, lambda *a: None
            while self.timer.date < end and self.timer.date < self.timer.final_date:
                self.leisure_table(policies.active_key(self.timer.date))
                next(self.timer)
        finally:
            self.timer = saved
            del self.device.set_leisure_tables
        if self._leisure_stack:
            self.device.set_leisure_tables(np.stack([t[0] for t in self._leisure_stack]),
                                           np.stack([t[1] for t in self._leisure_stack]))
        return len(self._leisure_stack)

    def leisure_table(self, policy_key: tuple = ()) -> int:
        """``ActivityManager.do_timestep`` regenerates the leisure probabilities on every step
        (``activity_manager.py:200-209``); they only depend on the time slot (duration, working
        hours, day type, hour) and on the active policies, so each distinct slot is computed once,
        appended to the device-resident stack and selected per step by index."""
        leisure = self.activity_manager.leisure
        acts = self.timer.activities
        if leisure is None or not self.world.has_leisure or "leisure" not in acts:
            return 0
        date = self.timer.date
        key = (round(self.timer.duration, 9), "primary_activity" in acts, date.weekday(), date.hour, policy_key)
        idx = self._leisure_tables.get(key)
        if idx is None:
            comp, _ = self.activity_manager.policies.regional_scalars(date, self.world.region_names, self._region_state)
            leisure.regional_compliance = dict(zip(self.world.region_names, comp))
            self.activity_manager.policies.apply_leisure_policies(date, leisure)  # activity_manager.py:200-204
            self._leisure_stack.append(leisure.generate_leisure_probabilities_for_timestep(
                delta_time=self.timer.duration, working_hours="primary_activity" in acts, date=date))
            idx = self._leisure_tables[key] = len(self._leisure_stack) - 1
            self.device.set_leisure_tables(np.stack([t[0] for t in self._leisure_stack]),
                                           np.stack([t[1] for t in self._leisure_stack]))
        return idx + 1

    def do_timestep(self, want_stats: bool = True) -> None:
        activities = self.timer.activities
        if self.events is not None:  # simulator.py:368-376
            self.events.apply(date=self.timer.date, world=self.world, activities=activities,
                              day_type=self.timer.day_type, simulator=self)
        if not activities:
            return  # simulator.py:377-379
        params = self.step_params(_lib.STEP_RECORD_EVENTS if self.record is not None else 0)
        dev = self.device
        if self.halo is None or self.p2p or not (self.activity_manager.activity_mask() & self.halo_activity_mask):
            stats = dev.step(params, want_stats=want_stats)
        else:
            stats = self._step_with_exchange(params, want_stats)
        self.step_ordinal += 1
        self._last_step_date = self.timer.date
        ep = self.epidemiology
        if ep is not None and ep.vaccination_campaigns is not None and (
                ep.current_date is None or self.timer.date.date() != ep.current_date.date()):
            # once per simulated day, behind the day's first health update (epidemiology.py:127-135,295-303)
            ep.current_date = self.timer.date
            if not self._vaccination_on_device:
                first = self.timer.initial_date.date()
                self._vaccination_first_day = first
                dev.set_vaccination(*ep.vaccination_campaigns.device_tables(first, self.variant_names))
                self._vaccination_on_device = True
            dev.vaccinate((self.timer.date.date() - self._vaccination_first_day).days, seed=self.seed)
            if self.record is not None:
                # `vaccines` table (vaccines.py:309-313): a row whenever somebody's dose number changes
                vs = dev.fetch_vaccination()
                prev = getattr(self, "_vaccination_dose_prev", None)
                if prev is None:
                    prev = np.full(self.world.n_people, -1, np.int8)
                changed = np.nonzero(vs["dose"] != prev)[0]
                if len(changed):
                    self.vaccine_events.append((self.timer.date, changed, vs["vaccine"][changed].copy(), vs["dose"][changed].copy()))
                self._vaccination_dose_prev = vs["dose"].copy()
        if stats is not None:
            self.last_stats = stats
            self.history.append(dict(date=self.timer.date, now=self.timer.now, **stats))
            if self.record is not None:
                summ = dev.fetch_summary().astype(np.int64)
                ev = dev.fetch_infection_events()
                # daily_infected of summary.csv: rows of the `infections` table by the region of the group the infection
                # happened in (interaction.py:679, records_writer.py:329-330) + the step's seeds (person's region)
                w = self.world
                by_group = np.bincount((np.asarray(w.grp_info)[ev["group"]] >> 16).astype(np.int64), minlength=w.n_regions) \
                    if len(ev["infected"]) else np.zeros(w.n_regions, np.int64)
                summ[:, 0] = by_group[:w.n_regions] + getattr(self, "_seeded_by_region", 0)
                self._seeded_by_region = 0
                self.region_history.append((self.timer.date, summ))
                if len(ev["infected"]):
                    self.infection_events.append((self.timer.date, ev))
                hev = dev.fetch_health_events()
                if len(hev["person"]):
                    self.health_events.append((self.timer.date, hev))

    def run(self) -> None:
        """``Simulator.run`` (``simulator.py:628-707``), including the checkpoints written after the last
        step of every day listed in ``checkpoint_save_dates`` (``:681-690``)."""
        pol = self.activity_manager.policies
        if pol is not None and hasattr(pol, "check_window"):
            pol.check_window(self.timer.date, self.timer.final_date)  # a policy this package does not evaluate would be active
        # the leisure tables of every time slot of the run, computed and uploaded once (a slot met for the first time
        # in the middle of the run re-uploads the stack and drops the captured step graphs)
        self.prepare_leisure_tables(days=max(1, (self.timer.final_date - self.timer.date).days + 1))
        while self.timer.date < self.timer.final_date:
            if self.epidemiology is not None:
                self.epidemiology.infection_seeds_timestep(self, self.timer)
            self.do_timestep()
            if (self.timer.date.date() in self.checkpoint_save_dates
                    and float(self.timer.now + self.timer.duration).is_integer()):
                self.save_checkpoint(self.timer.date.date())
            next(self.timer)

    def save_checkpoint(self, saving_date, path: Optional[str] = None) -> str:
        """``Simulator.save_checkpoint`` (``simulator.py:709-735``): ``checkpoint_<date>.<rank>`` under
        ``checkpoint_save_path`` (NumPy container, see ``june_b200.checkpoint``)."""
        import os
        from .checkpoint import save_checkpoint
        if path is None:
            os.makedirs(self.checkpoint_save_path, exist_ok=True)
            path = os.path.join(self.checkpoint_save_path, f"checkpoint_{saving_date}.{self.rank}.npz")
        save_checkpoint(self, path)
        return path

    def restore_checkpoint(self, path: str, reset_infections: bool = False, exact: bool = True) -> None:
        """``restore_simulator_to_checkpoint`` (``hdf5_savers/checkpoint_saver.py:162-218``)."""
        from .checkpoint import restore_checkpoint
        restore_checkpoint(self, path, reset_infections=reset_infections, exact=exact)

    SUMMARY_HEADER = ("time_stamp", "region", "current_infected", "daily_infected", "current_hospitalised",
                      "daily_hospitalised", "current_intensive_care", "daily_intensive_care", "daily_hospital_deaths",
                      "daily_deaths")  # records_writer.py:151-164

    def summary_rows(self) -> List[Dict[str, object]]:
        """Rows of the reference's ``summary.csv`` (``records_writer.py:151-164,693-790``): the reference writes one
        row per TIME STEP and region (``Epidemiology.do_timestep`` calls ``summarise_time_step`` every step,
        ``epidemiology.py:164-166``; the ``daily_`` columns hold the events of that step), with the day as time stamp,
        and skips rows whose eight numbers are all zero.  Same here; ``daily_recovered`` is an extra key of the
        dictionaries that ``write_summary_csv`` does not write.  Needs ``record`` to be set."""
        cols = self.device.SUMMARY_COLUMNS
        out = []
        for date, summ in self.region_history:
            for r, name in enumerate(self.world.region_names):
                row = dict(time_stamp=date.date(), region=name, **{c: int(summ[r, k]) for k, c in enumerate(cols)})
                if sum(row[c] for c in self.SUMMARY_HEADER[2:]) > 0:
                    out.append(row)
        return out

    def infection_rows(self) -> List[Dict[str, object]]:
        """Rows of the reference's ``infections`` record table (``Record.accumulate(table_name=
        "infections", ...)``, ``interaction.py:664-683``; ``records/event_records_writer.py:60-84``):
        ``time_stamp, location_specs, location_ids, region_names, infector_ids, infected_ids,
        infection_ids``.  Ids are the caller's (``world.person_ids`` / ``world.group_ids`` when the
        world was compiled from reference objects, else indices + ``person_id_base``)."""
        w = self.world
        grp_sg = w.group_supergroup()
        pid = w.person_ids if w.person_ids is not None else np.arange(w.n_people, dtype=np.int64) + w.person_id_base
        if w.n_halo:
            pid = np.concatenate([pid, w.halo_gid.astype(np.int64)])
        gid = w.group_ids if w.group_ids is not None else np.arange(w.n_groups, dtype=np.int64)
        rows = []
        for date, ev in self.infection_events:
            for a, b, g, v in zip(ev["infected"], ev["infector"], ev["group"], ev["variant"]):
                rows.append(dict(time_stamp=date, location_specs=w.supergroups[grp_sg[g]].spec, location_ids=int(gid[g]),
                                 region_names=w.region_names[int(w.grp_info[g] >> 16)] if w.region_names else "",
                                 infector_ids=int(pid[b]) if b != 0xFFFFFFFF else -1, infected_ids=int(pid[a]),
                                 infection_ids=int(v)))
        return rows

    def health_event_rows(self) -> Dict[str, List[Dict[str, object]]]:
        """Rows of the reference's ``deaths`` (``epidemiology.py:176-186``), ``recoveries`` (``:209-214``),
        ``hospital_admissions`` / ``icu_admissions`` / ``discharges`` (``medical_care_policies.py:463-515``)
        record tables, keyed by table name; columns follow ``records/event_records_writer.py``."""
        w = self.world
        grp_sg = w.group_supergroup()
        pid = w.person_ids if w.person_ids is not None else np.arange(w.n_people, dtype=np.int64) + w.person_id_base
        gid = w.group_ids if w.group_ids is not None else np.arange(w.n_groups, dtype=np.int64)
        names = self.device.HEALTH_EVENT_KINDS
        out: Dict[str, List[Dict[str, object]]] = {n: [] for n in names.values()}
        for date, ev in self.health_events:
            for p, k, g in zip(ev["person"], ev["kind"], ev["group"]):
                table = names[int(k)]
                if table == "recoveries":
                    out[table].append(dict(time_stamp=date, recovered_person_ids=int(pid[p]), infection_ids=-1))
                elif table == "symptoms":
                    out[table].append(dict(time_stamp=date, infected_ids=int(pid[p]), new_symptoms=int(g), infection_ids=-1))
                elif table == "deaths":
                    out[table].append(dict(time_stamp=date, location_specs=w.supergroups[grp_sg[g]].spec if g >= 0 else "",
                                           location_ids=int(gid[g]) if g >= 0 else -1, dead_person_ids=int(pid[p])))
                else:
                    out[table].append(dict(time_stamp=date, hospital_ids=int(gid[g]), patient_ids=int(pid[p])))
        return out

    def vaccine_rows(self) -> List[Dict[str, object]]:
        """Rows of the reference's ``vaccines`` record table (``vaccines.py:309-313``,
        ``event_records_writer.py:180-190``): ``time_stamp, vaccinated_ids, vaccine_names, dose_numbers``."""
        w = self.world
        pid = w.person_ids if w.person_ids is not None else np.arange(w.n_people, dtype=np.int64) + w.person_id_base
        names = [v.name for v in self.epidemiology.vaccination_campaigns.vaccines()]
        return [dict(time_stamp=date, vaccinated_ids=int(pid[p]), vaccine_names=names[int(v)], dose_numbers=int(d))
                for date, persons, vacs, doses in self.vaccine_events for p, v, d in zip(persons, vacs, doses)]

    def write_summary_csv(self, path: str) -> None:
        import csv
        rows = self.summary_rows()
        header = list(self.SUMMARY_HEADER)
        with open(path, "w", newline="") as f:
            wr = csv.DictWriter(f, fieldnames=header)
            wr.writeheader()
            for r in rows:
                wr.writerow({k: (r[k].strftime("%Y-%m-%d") if k == "time_stamp" else r[k]) for k in header})

    def daily_summary(self) -> List[Dict[str, object]]:
        """Daily curves in the spirit of ``summary.csv`` (``records_writer.py:151-164``), whole domain."""
        days: Dict[datetime.date, Dict[str, object]] = {}
        for h in self.history:
            d = h["date"].date()
            row = days.setdefault(d, dict(time_stamp=d, daily_infected=0, daily_deaths=0, daily_recovered=0))
            row["daily_infected"] += h["n_new_infections"]
            row["daily_deaths"] += h["n_deaths_step"]
            row["daily_recovered"] += h["n_recovered_step"]
            row["current_infected"] = h["n_infected"]
            row["current_hospitalised"] = h["n_hospitalised"]
            row["current_intensive_care"] = h["n_icu"]
        return [days[k] for k in sorted(days)]
