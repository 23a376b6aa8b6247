#!/usr/bin/env python
"""Benchmark of the JUNE per-timestep hot path (``Simulator.do_timestep``) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--people P]

Metric (BASELINE.json): person-timesteps / second at 56 M agents on 1 / 2 / 4 / 8 B200.  A "step" is one
``do_timestep`` of the default schedule (config_simulation.yaml: 5 weekday / 4 weekend steps) over a
synthetic world built by ``june_b200.synth`` (SURVEY.md section 8d recipe).  The workload is BASELINE.json
configs[2]: the 56 M-person England world, STRONG-scaled: N GPUs hold N domains of 56 M / N people, 5 % of
the workers commute to a company of a neighbouring domain and are exchanged over peer memory (NVLink) on
the steps with the primary activity.  ``--people P`` sets another total (2 700 000 = configs[1]);
``--people-per-gpu P`` fixes the people per GPU instead (weak scaling, the configs[4] sweep).

Prints ONE JSON line on rank 0 (contract in the task statement):
  value     device-resident throughput: K steps timed with CUDA events on the library stream,
            L2 flushed (256 MiB memset) between steps, max over ranks;
  e2e       the same K steps through the public API ``Simulator.do_timestep()``: per-step host
            parameters copied h2d from pinned memory, per-step statistics copied d2h, wall clock;
  roofline  interaction kernel of the busiest supergroup (households, k_tiles): SURVEY.md 8(d)
            algorithmic bytes / CUDA-event time against the measured HBM copy bandwidth
            (MEASURED_PEAKS.json); `traffic` = DRAM bytes per launch from the committed ncu capture
            (profiles/traffic.json);
  cpu_baseline  the CPU oracle (oracle/june_oracle.c, a port of the reference algorithm) on the
            host cores, on a bounded sample of the same workload.
``--impl reference`` times that CPU port alone with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from june_b200.disease import DiseaseConfig, outcome_probabilities, synthetic_rates  # noqa: E402
from june_b200.interaction import Interaction, load_defaults  # noqa: E402
from june_b200.policy import Policies  # noqa: E402
from june_b200.simulator import Simulator  # noqa: E402
from june_b200.synth import generate_domain, generate_world, wire_halos_distributed  # noqa: E402
from june_b200.timer import Timer  # noqa: E402
from june_b200.world import ACT_RESIDENCE  # noqa: E402

METRIC = "person-timesteps/sec"
PEOPLE_TOTAL = 56_000_000  # BASELINE.json: the metric is quoted at 56 M agents
HBM_FALLBACK_GBS = 6650.0


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons DURING the timed region: NVML polled every 5 ms from a
    thread (the timed legs last tens of milliseconds, `nvidia-smi -lms 100` would see one sample);
    falls back to one `nvidia-smi` query if NVML is unavailable."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index, self.sm, self.mask, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self.thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def poll():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                    except Exception:
                        pass
                    time.sleep(0.005)

            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=1)
        if self.sm:
            return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(k for k, bit in self.REASONS.items() if self.mask & bit), "samples": len(self.sm)}
        try:
            q = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout.split(",")
            return {"sm_mhz": float(q[0]), "sm_max_mhz": float(q[1]), "reasons": [], "samples": 1,
                    "note": "NVML unavailable: single nvidia-smi sample right after the timed legs"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}


def lockdown_policies(dc):
    """The reference's default policy file (policy_covid19.yaml) with its dates moved from the years 3020 / 3021, where
    the shipped file parks them, to 2020 / 2021: the UK's spring-2020 measures (quarantine, shielding, school /
    company / venue closures, social distancing, reduced leisure)."""
    import copy
    import tempfile
    import yaml

    def shift(x):
        if isinstance(x, dict):
            return {k: shift(v) for k, v in x.items()}
        if isinstance(x, str) and x[:4] in ("3020", "3021"):
            return "2" + x[1:]
        return x
    cfg = shift(copy.deepcopy(load_defaults()["policy"]["covid19"]))
    with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
        yaml.safe_dump(cfg, f)
    try:
        return Policies.from_file(f.name, disease_config=dc)
    finally:
        os.unlink(f.name)


def vaccination_campaigns():
    """Two campaigns running over the bench window: the over-50s (two doses), then everybody between 16 and 50."""
    from june_b200.vaccination import Vaccine, VaccinationCampaign, VaccinationCampaigns
    v = Vaccine.from_config_dict("Pfizer", dict(
        days_administered_to_effective=[14, 7], days_effective_to_waning=[60, 180], days_waning=[30, 90], waning_factor=0.8,
        sterilisation_efficacies=[{"Covid19": {"0-100": 0.52}}, {"Covid19": {"0-100": 0.90}}],
        symptomatic_efficacies=[{"Covid19": {"0-100": 0.60}}, {"Covid19": {"0-100": 0.95}}]))
    return VaccinationCampaigns([
        VaccinationCampaign(vaccine=v, days_to_next_dose=[0, 21], dose_numbers=[0, 1], start_time="2020-03-20",
                            end_time="2020-06-01", group_by="age", group_type="50-100", group_coverage=0.8),
        VaccinationCampaign(vaccine=v, days_to_next_dose=[0, 21], dose_numbers=[0, 1], start_time="2020-03-27",
                            end_time="2020-07-01", group_by="age", group_type="16-50", group_coverage=0.6),
    ])


def build_setup(n_people: int, rank: int, world_size: int, dist=None, leisure: bool = False, config4: bool = False):
    leisure = leisure or config4
    defaults = load_defaults()
    sb = defaults["company_sector_betas"]
    if world_size == 1:
        world = generate_world(n_people, seed=0, sector_betas=sb, leisure=leisure, visits=config4)
    else:
        draft = generate_domain(n_people, rank, world_size, seed=0, sector_betas=sb, leisure=leisure, visits=config4)
        world = wire_halos_distributed(draft, dist)
    inter = Interaction.from_file()
    dc = DiseaseConfig("covid19")
    rates, bins = synthetic_rates()
    prob = outcome_probabilities(dc, rates, bins)
    cfg = defaults["simulation"]
    cfg["time"]["total_days"] = 10_000
    cfg["time"]["initial_day"] = "2020-03-02 8:00"  # a Monday
    policies = Policies.from_file(disease_config=dc)
    if config4:
        cfg["time"]["initial_day"] = "2020-03-30 8:00"  # a Monday inside the spring-2020 lockdown
        policies = lockdown_policies(dc)
    return world, inter, dc, prob, cfg, policies


def seed_epidemic(sim: Simulator, rank: int):
    """~3 % of the residents infected over the 12 days before the start, so that the timed steps
    see infections in every stage (exposed ... hospitalised) instead of a fresh seed."""
    n = sim.world.n_people
    rng = np.random.default_rng([17, rank])
    chosen = rng.choice(n, size=int(0.03 * n), replace=False).astype(np.uint32)
    for k, part in enumerate(np.array_split(chosen, 12)):
        sim.device.infect(np.sort(part), time=-(12.0 - k), seed=1234, step=0xFFFF0000 + k)


def interaction_bytes(world, sg, n_present: int) -> int:
    """Algorithmic bytes of one launch of the interaction kernel over supergroup ``sg``: SURVEY.md
    section 8(d), ``A_inf = 14 N_present + 8 S_active + 8 G_active`` -- per present person the member
    index (4), flags (1), susceptibility or transmission probability (8) and the result flag (1);
    per non-empty subgroup its CSR offset and group id (8); per group its beta/spec/region word (8).
    The figure is a property of the problem, not of this library's encoding (which reads fewer
    bytes: DESIGN.md section 4).  Subgroups and groups are counted from the registrations and scaled
    by the fraction of registered members present in the step."""
    lo, hi = int(world.grp_offset[sg.first_group]), int(world.grp_offset[sg.first_group + sg.n_groups])
    n_reg = max(1, hi - lo)
    grp_of = np.repeat(np.arange(sg.n_groups, dtype=np.int64),
                       np.diff(world.grp_offset[sg.first_group: sg.first_group + sg.n_groups + 1].astype(np.int64)))
    lsg = (world.reg_info[lo:hi] & 0xFF).astype(np.int64)
    n_sub = int(np.unique(grp_of * 256 + lsg).size)
    n_grp = int(np.unique(grp_of).size)
    f = min(1.0, n_present / n_reg)
    return int(14 * n_present + 8 * n_sub * f + 8 * n_grp * f)


def kernels_by_supergroup(described: str) -> dict:
    """Which interaction kernels the library launches for every supergroup, from its own layout summary
    (``jb_world_describe``): ``k_span`` for a mirrored single-subgroup supergroup packed into work items, ``k_tiles``
    for tiled small groups, ``k_groups`` for the warp / CTA lists."""
    import re
    out = {}
    for m in re.finditer(r"supergroup (\d+): .*?tiles (\d+)( \(none valid\))?, span items (\d+), warp-list (\d+), cta-list (\d+)", described):
        k, tiles, none_valid, span, n_w, n_c = int(m.group(1)), int(m.group(2)), m.group(3), int(m.group(4)), int(m.group(5)), int(m.group(6))
        names = []
        if span:
            names.append("k_span")
        elif tiles and not none_valid:
            names.append("k_tiles")
        if n_w or n_c:
            names.append("k_groups")
        out[k] = " + ".join(names) if names else "k_groups"
    return out


def workload_name(n_total: int, leisure: bool, config4: bool = False) -> str:
    """The workload both arms run (same generator, schedule and seeding)."""
    if config4:
        return (f"synthetic {n_total / 1e6:.1f}M-person world (households, schools, companies, care homes, hospitals, pubs, "
                "cinemas, groceries, gyms, household and care-home visits with leisure), default 5/4-step day schedule from Monday 2020-03-30 08:00, the "
                "reference's default policy file with its dates in 2020 (spring lockdown: quarantine, shielding, school / "
                "company / venue closures, social distancing, reduced leisure), two vaccination campaigns, ~3% seeded prevalence")
    return (f"synthetic {n_total / 1e6:.1f}M-person world (households, schools, companies, care homes, hospitals"
            + (", pubs, cinemas, groceries, gyms with leisure" if leisure else "")
            + "), default 5/4-step day schedule from Monday 2020-03-02 08:00, default policy file, ~3% seeded prevalence")


def config_dict(n_total: int, leisure: bool, config4: bool = False) -> dict:
    """Identical in both arms (the driver compares the dicts): what is simulated, not how."""
    return {"workload": workload_name(n_total, leisure, config4), "people_total": n_total,
            "l2": "inputs larger than the cache (>= 14 GB at 56 M people against 126 MB of L2) and the L2 flushed between "
                  "timed steps"}


def sizes(args, world_size: int):
    """(people per GPU, total, scaling).  Default: 56 M people in total, strong-scaled over the GPUs."""
    if args.people_per_gpu:
        return args.people_per_gpu, args.people_per_gpu * world_size, "weak"
    total = args.people or PEOPLE_TOTAL
    return total // world_size, (total // world_size) * world_size, "strong"


def run_gpu(args):
    import torch
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world_size > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_people, n_total, scaling = sizes(args, world_size)
    world, inter, dc, prob, cfg, policies = build_setup(n_people, rank, world_size, dist, leisure=args.leisure, config4=args.config4)
    sim = Simulator.from_file(world, inter, policies=policies, disease_config=dc, outcome_prob=prob, device=local_rank,
                              seed=2020, distributed=world_size > 1)
    sim.timer = Timer.from_config(cfg)
    sim.activity_manager.timer = sim.timer
    if args.config4:
        from june_b200.simulator import Epidemiology
        sim.epidemiology = Epidemiology(vaccination_campaigns=vaccination_campaigns())
    seed_epidemic(sim, rank)
    sim.prepare_leisure_tables(days=45)  # every slot of every leg (warm-up, three timed legs, the rewinds between them)
    # every step graph the schedule needs is captured and instantiated here, before anything is timed
    n_shapes = sim.prepare_steps(days=8)
    dev = sim.device
    if rank == 0:
        print(dev.describe(), file=sys.stderr, flush=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        dev.sync()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def flush_l2():
        flush.zero_()
        torch.cuda.synchronize()

    # ---- warm-up ------------------------------------------------------------------------------
    for _ in range(args.warmup):
        sim.do_timestep()
        next(sim.timer)
    barrier()

    # ---- leg 1: device-resident (CUDA events on the library stream, L2 flushed between steps) --
    # The K steps are QUEUED: the L2 flush is a memset on the library's stream in front of each
    # step's event bracket, nothing waits on the host between steps (statistics are read once, with
    # the last step), so on N > 1 GPUs the ranks are kept in step by the halo exchange itself and not
    # by host jitter.  person-steps use the alive count at the END of the leg (a lower bound).
    phase0 = []

    def rewind():
        """Every timed leg walks the SAME sequence of time slots (the cost of a step depends on its
        activities: one weekday step in five has the primary activity): the clock is advanced, without
        stepping the device, to the next slot with the weekday and shift the first leg started on."""
        if not phase0:
            phase0.append((sim.timer.date.weekday(), sim.timer.shift))
            return
        while (sim.timer.date.weekday(), sim.timer.shift) != phase0[0]:
            next(sim.timer)

    rewind()
    dev.timing_enable(1)
    dev.timing_read(reset=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    align = None
    if dist is not None and sim._ext_stream is not None:
        # Only one step in five exchanges halo persons, so queued ranks drift apart between exchanges and a
        # rank that reaches the next exchange early would count its peer's L2 flush (which is outside every
        # bracket) as step time.  A one-word all-reduce on the library's stream, between the flush and the
        # step's bracket, lines the ranks up on the device -- no host synchronisation, not timed.
        align = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", local_rank))
    for k in range(args.steps):
        dev.flush_l2()
        if align is not None:
            with torch.cuda.stream(sim._ext_stream):
                dist.all_reduce(align)
        sim.do_timestep(want_stats=(k == args.steps - 1))
        next(sim.timer)
    barrier()
    t_wall1 = time.perf_counter()
    n_alive_last = n_people - sim.last_stats["n_dead_total"]
    person_steps = n_alive_last * args.steps
    tim = dev.timing_read()
    step_ms = tim["step_ms"] / max(1, tim["n_steps"])
    launches = int(tim["n_launches"])

    # ---- leg 1b: roofline of the interaction kernels.  In the normal step the supergroups'
    # kernels overlap on side streams; here they run one after the other on the main stream so
    # that the CUDA-event bracket around each supergroup's launches measures that kernel alone.
    rewind()
    dev.set_concurrency(False)
    dev.timing_enable(2)
    dev.timing_read(reset=True)
    n_roof = min(args.steps, 27)
    act_res = []
    for _ in range(n_roof):
        flush_l2()
        sim.do_timestep(want_stats=True)
        act_res.append(sim.last_stats["n_by_activity"][ACT_RESIDENCE])
        next(sim.timer)
    barrier()
    tim_serial = dev.timing_read()
    per_sg = dev.timing_read_supergroups()
    dev.set_concurrency(True)
    dev.timing_enable(0)

    # ---- leg 2: end to end through the public API (host params h2d, stats d2h, wall clock) -----
    rewind()
    barrier()
    e2e_steps_s = 0.0
    e2e_person_steps = 0
    per_step = {"n_new_infections": 0, "n_infected": 0, "n_draws": 0}
    for _ in range(args.steps):
        flush_l2()
        t0 = time.perf_counter()
        sim.do_timestep(want_stats=True)
        e2e_steps_s += time.perf_counter() - t0
        e2e_person_steps += n_people - sim.last_stats["n_dead_total"]
        for k_ in per_step:
            per_step[k_] += sim.last_stats[k_]
        next(sim.timer)
    barrier()
    per_step = {k_: round(v / args.steps, 1) for k_, v in per_step.items()}

    clocks = sampler.stop()  # sampled across the three timed legs

    # ---- reduce over ranks -----------------------------------------------------------------------
    vals = torch.tensor([step_ms * args.steps, e2e_steps_s * 1e3, float(person_steps), float(e2e_person_steps)],
                        dtype=torch.float64, device=torch.device("cuda", local_rank))
    if dist is not None:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms, ps, e2e_ps = float(mx[0]), float(mx[1]), float(sm[2]), float(sm[3])
    else:
        total_ms, e2e_ms, ps, e2e_ps = (float(x) for x in vals)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    # Roofline of every supergroup's interaction kernel(s): SURVEY.md 8(d) algorithmic bytes of one launch over
    # the mean CUDA-event time of the supergroup's launches (serialised leg).  The headline object is the
    # supergroup with the most registered slots (households: it also runs in every step).
    slots = [int(world.grp_offset[g.first_group + g.n_groups]) - int(world.grp_offset[g.first_group]) for g in world.supergroups]
    k_top = int(np.argmax(slots))
    care = int((world.care_home_resident != 0).sum())
    by_sg = {}
    kernels_of = kernels_by_supergroup(dev.describe())
    for k, sg in enumerate(world.supergroups):
        n_launch = per_sg[k]["launches"]
        if not n_launch or not slots[k]:
            continue
        present = max(0, int(np.mean(act_res)) - care) if sg.name == "households" else slots[k]
        a_bytes = interaction_bytes(world, sg, present)
        t_s = per_sg[k]["ms"] / n_launch * 1e-3
        gbs = a_bytes / t_s / 1e9 if t_s > 0 else 0.0
        by_sg[sg.name] = {"kernel": kernels_of.get(k, "k_tiles" if sg.launch_mode == 0 else "k_groups"), "bytes_per_launch": a_bytes,
                          "us_per_launch": round(t_s * 1e6, 2), "achieved": round(gbs, 1), "frac": round(gbs / peak, 4),
                          "launches": int(n_launch)}
    top = by_sg[world.supergroups[k_top].name]
    roofline = {"bound": "hbm", "kernel": top["kernel"], "supergroup": world.supergroups[k_top].name,
                "achieved": top["achieved"], "peak": peak, "peak_source": peak_src,
                "unit": "GB/s", "frac": top["frac"], "traffic": None,
                "bytes_per_launch": top["bytes_per_launch"], "us_per_launch": top["us_per_launch"],
                "by_supergroup": by_sg,
                "interaction_ms_per_step_serial": round(tim_serial["interaction_ms"] / max(1, tim_serial["n_steps"]), 4),
                "timing": "CUDA events on the launch stream, supergroup kernels serialised, L2 flushed between steps"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            tf = json.load(open(traffic_file))
            roofline["traffic"] = tf.get(f"{roofline['kernel']}@{n_people}")
            for name, o in by_sg.items():
                o["traffic"] = tf.get(f"{name}@{n_people}")
        except Exception:
            pass

    cpu = (cpu_baseline(world, inter, dc, prob, cfg, policies, budget_s=args.cpu_seconds, config4=args.config4)
           if world_size == 1 and args.cpu_seconds > 0 else None)
    # per step: the parameter block (StepDev header of 64 B + the beta table [rule][region] of doubles) is pulled
    # from pinned host memory by k_step_begin; the counter block (20 step counters + 9 per region) is pushed
    # into pinned host memory by k_step_end
    h2d = int(64 + world.n_regions * len(world.beta_rules) * 8)
    d2h = int((20 + 9 * world.n_regions) * 4)
    out = {
        "metric": METRIC, "value": ps / (total_ms * 1e-3), "unit": "person-timesteps/s", "n_gpus": world_size,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(n_total, args.leisure, args.config4),
        "people_per_gpu": n_people,
        "parallelism": "1 domain per GPU" + (", halo exchange over peer memory on primary-activity steps; ranks aligned by a "
                                              "one-word all-reduce before each timed bracket" if world_size > 1 else ""),
        "e2e": {"value": e2e_ps / (e2e_ms * 1e-3), "unit": "person-timesteps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches,
        "step_graphs_prepared": n_shapes,
        "per_step_mean": per_step,
        "clocks": clocks,
        "roofline": roofline,
        "wall_s_leg1": round(t_wall1 - t_wall0, 3),
        "kernel_ms_by_supergroup": {s["name"]: round(s["ms"] / max(1, s["launches"]), 4) for s in per_sg},
    }
    if cpu is not None:
        out["cpu_baseline"] = cpu
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def oracle_engine(world, inter, dc, prob):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_engine  # the oracle is loaded here only as the timed CPU baseline
    eng = make_engine("oracle", world, prob, interaction=inter, disease=dc)
    # all host cores, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)
    eng._lib.jo_set_num_threads(int(os.cpu_count() or 1))
    return eng


def cpu_run(world, inter, dc, prob, cfg, policies, steps: int, warmup: int, budget_s: float = 0.0, config4: bool = False):
    """Time `steps` steps of the CPU port after `warmup` steps; with a budget, as many steps as fit in
    ~budget_s seconds (at least 3, at most `steps`), decided from the duration of the warm-up."""
    eng = oracle_engine(world, inter, dc, prob)
    from helpers import simulator_over  # (tests/ is on the path: oracle_engine) the product's host loop over the CPU port
    sim = simulator_over(eng).from_file(world, inter, policies=policies, disease_config=dc, outcome_prob=prob, seed=2020)
    sim.timer = Timer.from_config(cfg)
    sim.activity_manager.timer = sim.timer
    if config4:
        from june_b200.simulator import Epidemiology
        sim.epidemiology = Epidemiology(vaccination_campaigns=vaccination_campaigns())
    seed_epidemic(sim, 0)
    t0 = time.perf_counter()
    for _ in range(warmup):
        sim.do_timestep()
        next(sim.timer)
    if budget_s > 0 and warmup:
        per = (time.perf_counter() - t0) / warmup
        steps = int(max(3, min(steps, budget_s / max(per, 1e-6))))
    ps, t0 = 0, time.perf_counter()
    for _ in range(steps):
        sim.do_timestep()
        ps += world.n_people - sim.last_stats["n_dead_total"]
        next(sim.timer)
    dt = time.perf_counter() - t0
    threads = int(eng._lib.jo_num_threads())
    eng.close()
    return ps, dt, threads, steps


def cpu_baseline(world, inter, dc, prob, cfg, policies, budget_s: float = 15.0, config4: bool = False):
    """Oracle port on the host cores, bounded sample: as many steps of the same world as fit in
    ~budget_s seconds (at least 3)."""
    ps, dt, threads, steps = cpu_run(world, inter, dc, prob, cfg, policies, steps=40, warmup=2, budget_s=budget_s, config4=config4)
    return {"value": ps / dt, "unit": "person-timesteps/s", "cores": threads, "kind": "port",
            "sample": f"{steps} steps of the same {world.n_people}-person world (oracle/june_oracle.c, OpenMP, {threads} threads)"}


def run_reference(args):
    """The reference arm: the CPU port of the reference algorithm (the reference itself is pure Python at
    ~9e4 person-steps/s per core: DESIGN.md) on ALL host cores of the box, on the same total population as
    the GPU arm (one undivided world on rank 0 -- the GPU arm's N domains are a partition of the same
    recipe), same schedule, seeding, steps and warm-up."""
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    _, n_total, scaling = sizes(args, world_size)
    world, inter, dc, prob, cfg, policies = build_setup(n_total, 0, 1, leisure=args.leisure, config4=args.config4)
    ps, dt, threads, steps = cpu_run(world, inter, dc, prob, cfg, policies, steps=args.steps, warmup=args.warmup, config4=args.config4)
    v = ps / dt
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "person-timesteps/s", "n_gpus": world_size,
           "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
           "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_dict(n_total, args.leisure, args.config4),
           "parallelism": f"rank 0 only: the undivided {n_total}-person world on {threads} OpenMP threads (all host cores)",
           "cpu_baseline": {"value": v, "unit": "person-timesteps/s", "cores": threads, "kind": "port",
                            "sample": f"{steps} steps of the {n_total}-person world on rank 0, {threads} threads"},
           "e2e": {"value": v, "unit": "person-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--people", type=int, default=0,
                    help="TOTAL people over all GPUs (default 56 M = the metric's configuration; 2700000 = BASELINE.json configs[1])")
    ap.add_argument("--people-per-gpu", type=int, default=0, help="fix the people per GPU instead (weak scaling)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--leisure", action="store_true",
                    help="add pubs / cinemas / groceries / gyms and the leisure placement")
    ap.add_argument("--config4", action="store_true",
                    help="BASELINE.json configs[3]: leisure + the lockdown / closure policies of spring 2020 + vaccination campaigns")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
