"""Multi-GPU check, run under torchrun on N B200s (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/mgpu_check.py

Every rank runs its domain of ONE synthetic world on its GPU (CUDA path, NCCL all-to-all halo
exchange); rank 0 additionally runs the whole, undivided world on the CPU oracle.  The
decomposed GPU run must reproduce the single-domain oracle run: same infection counts per step,
same final person state.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)

import torch
import torch.distributed as dist

from june_b200.device import DeviceWorld
from june_b200.domains import partition
from june_b200.interaction import load_defaults
from june_b200.policy import Policies
from june_b200.simulator import Simulator
from june_b200.timer import Timer

from helpers import make_engine, simulator_over
from test_domains import aligned_world, common, seed_list


def make_sim(world, engine, distributed, days):
    inter, dc, prob = common()
    cfg = load_defaults()["simulation"]
    cfg["time"]["initial_day"] = "2020-03-02 8:00"
    cfg["time"]["total_days"] = days
    sim = simulator_over(engine).from_file(world, inter, policies=Policies.from_file(disease_config=dc), disease_config=dc,
                                           outcome_prob=prob, seed=13, distributed=distributed)
    sim.timer = Timer.from_config(cfg)
    sim.activity_manager.timer = sim.timer
    return sim


def main():
    rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    n_people, days = 60_000 * ws, 6
    w = aligned_world(n_people, ws)
    part = partition(w, ws)[rank]
    inter, dc, prob = common()
    eng = DeviceWorld(part, inter, dc, prob, device=lr)
    sim = make_sim(part, eng, True, days)
    seeds = seed_list(n_people)
    lo = part.person_id_base
    eng.infect((seeds[(seeds >= lo) & (seeds < lo + part.n_people)] - lo).astype(np.uint32), time=-4.0, seed=9, step=999)
    sim.run()
    st = eng.fetch_person_state()
    mine = dict(pstate=st["pstate"], tag=st["tag"], new=[h["n_new_infections"] for h in sim.history],
                halo=[h["n_halo_infected"] for h in sim.history])
    gathered = [None] * ws
    dist.all_gather_object(gathered, mine)
    ok = True
    if rank == 0:
        oracle = make_engine("oracle", w, prob, interaction=inter, disease=dc)
        ref = make_sim(w, oracle, False, days)
        oracle.infect(seeds, time=-4.0, seed=9, step=999)
        ref.run()
        rs = oracle.fetch_person_state()
        pstate = np.concatenate([g["pstate"] for g in gathered])
        tag = np.concatenate([g["tag"] for g in gathered])
        new = np.sum([g["new"] for g in gathered], axis=0)
        ref_new = np.array([h["n_new_infections"] for h in ref.history])
        halo = int(np.sum([g["halo"] for g in gathered]))
        ok = bool(np.array_equal(pstate, rs["pstate"]) and np.array_equal(tag, rs["tag"]) and np.array_equal(new, ref_new))
        print(f"mgpu_check: ranks={ws} people={n_people} steps={len(ref_new)} infections={int(ref_new.sum())} "
              f"infected_abroad={halo} halo_per_rank={[int(p.n_halo) for p in partition(w, ws)]} -> {'OK' if ok else 'MISMATCH'}",
              flush=True)
        if not ok:
            print("new per step (gpu)", new.tolist(), "\nnew per step (ref)", ref_new.tolist(), flush=True)
    flag = torch.tensor([1 if ok else 0], device=torch.device("cuda", lr))
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
