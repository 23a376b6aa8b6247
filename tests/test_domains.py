"""Domain decomposition on the CPU (oracle engine): the N>1 path without a GPU.

* K virtual domains of ONE world, exchanged through a loopback, reproduce the K=1 run: same
  infection events, same final state (Philox draws are keyed by the global person id, so the
  decomposition must not change who gets infected).
* The same two-domain run over ``torch.distributed`` (gloo, world_size 2) through the
  ``Simulator``'s HaloExchange: one process per rank, exactly the code path the GPU ranks use.
"""
import os
import sys

import numpy as np
import pytest

from june_b200 import _lib
from june_b200.disease import DiseaseConfig, outcome_probabilities, synthetic_rates
from june_b200.domains import LoopbackExchange, partition
from june_b200.interaction import Interaction, load_defaults
from june_b200.synth import generate_world

from helpers import default_beta_table, make_engine, simulator_over

MASKS = [0b10001, 0b10101, 0b10001, 0b10001, 0b10101]
DTS = [1 / 24, 8 / 24, 1 / 24, 3 / 24, 11 / 24]
FLAGS = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION
N_STEPS = 25


def common():
    inter = Interaction.from_file()
    dc = DiseaseConfig("covid19")
    rates, bins = synthetic_rates()
    return inter, dc, outcome_probabilities(dc, rates, bins)


def aligned_world(n_people, K):
    """Synthetic world whose hospital catchments do not cross the K-domain cuts (patients are
    dynamic members; cross-domain hospitalisation is out of scope, see june_b200/domains.py)."""
    from june_b200.domains import balanced_split, super_area_scores
    w = generate_world(n_people, seed=4, sector_betas=load_defaults()["company_sector_betas"])
    hosp = w.supergroup("hospitals")
    assert hosp.n_groups >= K
    # the cuts follow the reference's load score, which counts workers where they work: moving a hospital moves its
    # staff's weight, so place the hospitals and recompute until the cuts stay put
    cuts = balanced_split(super_area_scores(w), K)
    for _ in range(8):
        for k in range(hosp.n_groups):  # put hospital k into domain k % K
            w.grp_super_area[hosp.first_group + k] = cuts[k % K]
        again = balanced_split(super_area_scores(w), K)
        if np.array_equal(again, cuts):
            break
        cuts = again
    else:
        raise AssertionError("hospital placement does not settle")
    hsa = w.grp_super_area[hosp.first_group: hosp.first_group + hosp.n_groups].astype(np.int64)
    for d in range(K):
        inside = [hosp.first_group + k for k in range(hosp.n_groups) if cuts[d] <= hsa[k] < cuts[d + 1]]
        assert inside, "every domain needs a hospital for this test"
        for sa in range(cuts[d], cuts[d + 1]):
            w.sa_hospital[sa] = inside[sa % len(inside)]
    return w


def seed_list(n):
    return np.sort(np.random.default_rng(2).choice(n, n // 40, replace=False)).astype(np.uint32)


def run_single(world):
    inter, dc, prob = common()
    e = make_engine("oracle", world, prob, interaction=inter, disease=dc)
    e.infect(seed_list(world.n_people), time=-4.0, seed=9, step=999)
    beta = default_beta_table(world, inter)
    now, events = 0.0, []
    for i in range(N_STEPS):
        st = e.step(e.make_params(now, DTS[i % 5], i, MASKS[i % 5], seed=13, flags=FLAGS, beta=beta))
        events.append((st["n_new_infections"], st["n_infected"], st["n_dead_total"], st["n_hospitalised"]))
        now += DTS[i % 5]
    s = e.fetch_person_state()
    e.close()
    return events, s


def test_partition_is_consistent():
    w = generate_world(30_000, seed=4)
    parts = partition(w, 3)
    assert sum(p.n_people for p in parts) == w.n_people
    assert sum(p.n_groups for p in parts) == w.n_groups
    assert sum(p.n_members for p in parts) == w.n_members
    assert sum(p.n_halo for p in parts) == sum(len(p.out_person) for p in parts) > 0
    for r, p in enumerate(parts):
        for q, other in enumerate(parts):
            assert p.out_counts[q] == other.in_counts[r]


@pytest.mark.parametrize("K", [2, 3])
def test_virtual_domains_reproduce_single_domain_run(K):
    w = aligned_world(160_000, K)
    ref_events, ref_state = run_single(w)
    inter, dc, prob = common()
    parts = partition(w, K)
    engines = [make_engine("oracle", p, prob, interaction=inter, disease=dc) for p in parts]
    seeds = seed_list(w.n_people)
    for p, e in zip(parts, engines):
        lo = p.person_id_base
        mine = seeds[(seeds >= lo) & (seeds < lo + p.n_people)] - lo
        e.infect(mine.astype(np.uint32), time=-4.0, seed=9, step=999)
    ex = LoopbackExchange(parts, engines[0].halo_record_size())
    betas = [default_beta_table(p, inter) for p in parts]
    now = 0.0
    for i in range(N_STEPS):
        ps = [e.make_params(now, DTS[i % 5], i, MASKS[i % 5], seed=13, flags=FLAGS, beta=b) for e, b in zip(engines, betas)]
        for r, e in enumerate(engines):
            e.step_begin(ps[r], ex.send[r].ctypes.data)
        ex.exchange_people()
        for r, e in enumerate(engines):
            e.step_interact(ex.recv[r].ctypes.data, ex.ret_send[r].ctypes.data)
        ex.exchange_infections()
        stats = [e.step_finish(ex.ret_recv[r].ctypes.data) for r, e in enumerate(engines)]
        # a resident infected abroad is counted where it was infected (n_new_infections) ...
        tot = (sum(s["n_new_infections"] for s in stats), sum(s["n_infected"] for s in stats),
               sum(s["n_dead_total"] for s in stats), sum(s["n_hospitalised"] for s in stats))
        assert tot == ref_events[i], f"step {i}: {tot} vs {ref_events[i]}"
        now += DTS[i % 5]
    state = [e.fetch_person_state() for e in engines]
    for key in ("pstate", "tag"):
        merged = np.concatenate([s[key] for s in state])
        # the chosen-activity bits of commuters are identical too: placement is per resident
        assert np.array_equal(merged, ref_state[key]), key
    np.testing.assert_allclose(np.concatenate([s["trans_prob"] for s in state]), ref_state["trans_prob"], rtol=1e-12)
    assert ref_events[-1][1] > 2000
    for e in engines:
        e.close()


def _gloo_worker(rank, world_size, port, n_people, out_dir):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    from june_b200.policy import Policies
    from june_b200.simulator import Simulator
    from june_b200.timer import Timer
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world_size)
    w = aligned_world(n_people, world_size)
    part = partition(w, world_size)[rank]
    inter, dc, prob = common()
    eng = make_engine("oracle", part, prob, interaction=inter, disease=dc)
    cfg = load_defaults()["simulation"]
    cfg["time"]["initial_day"] = "2020-03-02 8:00"
    cfg["time"]["total_days"] = 4
    sim = simulator_over(eng).from_file(part, inter, policies=Policies.from_file(disease_config=dc), disease_config=dc,
                                        outcome_prob=prob, seed=13, distributed=True)
    sim.timer = Timer.from_config(cfg)
    sim.activity_manager.timer = sim.timer
    seeds = seed_list(w.n_people)
    lo = part.person_id_base
    eng.infect((seeds[(seeds >= lo) & (seeds < lo + part.n_people)] - lo).astype(np.uint32), time=-4.0, seed=9, step=999)
    sim.run()
    st = eng.fetch_person_state()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), pstate=st["pstate"], tag=st["tag"],
             new=np.array([h["n_new_infections"] for h in sim.history]),
             halo=np.array([h["n_halo_infected"] for h in sim.history]))
    dist.destroy_process_group()


def test_two_ranks_over_gloo_match_single_domain(tmp_path):
    import torch.multiprocessing as mp
    from june_b200.policy import Policies
    from june_b200.simulator import Simulator
    from june_b200.timer import Timer
    n_people, port = 110_000, 29517 + os.getpid() % 200
    mp.spawn(_gloo_worker, args=(2, port, n_people, str(tmp_path)), nprocs=2, join=True)
    # single-domain run of the same world through the same Simulator
    w = aligned_world(n_people, 2)
    inter, dc, prob = common()
    eng = make_engine("oracle", w, prob, interaction=inter, disease=dc)
    cfg = load_defaults()["simulation"]
    cfg["time"]["initial_day"] = "2020-03-02 8:00"
    cfg["time"]["total_days"] = 4
    sim = simulator_over(eng).from_file(w, inter, policies=Policies.from_file(disease_config=dc), disease_config=dc,
                                        outcome_prob=prob, seed=13)
    sim.timer = Timer.from_config(cfg)
    sim.activity_manager.timer = sim.timer
    eng.infect(seed_list(n_people), time=-4.0, seed=9, step=999)
    sim.run()
    ref = eng.fetch_person_state()
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(np.concatenate([r0["pstate"], r1["pstate"]]), ref["pstate"])
    assert np.array_equal(np.concatenate([r0["tag"], r1["tag"]]), ref["tag"])
    assert np.array_equal(r0["new"] + r1["new"], np.array([h["n_new_infections"] for h in sim.history]))
    assert (r0["halo"] + r1["halo"]).sum() > 0, "no resident was infected abroad: the return exchange was not exercised"


def test_domain_score_follows_the_reference():
    """domain_decomposition.py:22,107-113: 5 * people + workers + pupils + commuters per super area; the cuts balance it."""
    from june_b200.domains import balanced_split, super_area_scores
    from june_b200.world import ACT_PRIMARY
    w = generate_world(40_000, seed=9, sector_betas=load_defaults()["company_sector_betas"])
    score = super_area_scores(w)
    S = w.n_super_areas
    people = np.bincount(w.super_area, minlength=S)
    grp_of = np.repeat(np.arange(w.n_groups), np.diff(w.grp_offset.astype(np.int64)))
    prim = (w.reg_info >> 8) == ACT_PRIMARY
    at_work = np.bincount(w.grp_super_area[grp_of[prim]], minlength=S)
    expect = 5.0 * people + at_work
    if w.person_commute is not None:
        expect = expect + np.bincount(w.super_area[w.person_commute >= 0], minlength=S)
    assert np.array_equal(score, expect) and at_work.sum() > 0
    K = 4
    cuts = balanced_split(score, K)
    loads = np.array([score[cuts[d]:cuts[d + 1]].sum() for d in range(K)])
    assert loads.max() - loads.min() <= 2 * score.max(), "prefix split balanced to within a super area"
