"""Shared test plumbing: loads the golden fixtures, the CPU oracle (oracle/libjune_oracle.so, test
infrastructure) and builds engines over either ABI library."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from june_b200 import _lib
from june_b200.device import Engine
from june_b200.disease import DiseaseConfig
from june_b200.interaction import Interaction
from june_b200.world import MAX_STAGES, WorldSoA

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libjune_oracle.so")

_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        src = os.path.join(ORACLE_DIR, "june_oracle.c")
        if not os.path.exists(ORACLE_LIB) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
        _oracle = _lib.bind(ctypes.CDLL(ORACLE_LIB), "jo_")
    return _oracle


def have_gpu() -> bool:
    try:
        return os.path.exists(_lib.LIB_PATH) and _lib.load().jb_device_count() > 0
    except Exception:
        return False


class Golden:
    def __init__(self, name: str):
        self.world = WorldSoA.load(os.path.join(GOLDEN, name + "_world.npz"))
        self.z = np.load(os.path.join(GOLDEN, name + "_trace.npz"))
        self.n_steps = int(self.z["n_steps"])
        self.outcome_prob = self.z["health_index_prob"]

    def step(self, i: int) -> dict:
        pre = f"s{i}_"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}


def rows_to_dicts(rows: np.ndarray):
    out = []
    for r in rows:
        n = int(r[2])
        out.append(dict(person=int(r[0]), start_time=float(r[1]), max_tag=int(r[3]), probability=float(r[4]),
                        tpar=[float(x) for x in r[5:10]], traj_time=[float(x) for x in r[10:10 + n]],
                        traj_tag=[int(x) for x in r[10 + MAX_STAGES:10 + MAX_STAGES + n]], stage=0))
    return out


def make_engine(kind: str, world: WorldSoA, outcome_prob: np.ndarray, interaction=None, disease=None, **kw) -> Engine:
    """kind = 'oracle' (CPU restatement) or 'cuda' (the product library)."""
    interaction = interaction or Interaction.from_file()
    disease = disease or DiseaseConfig("covid19")
    if kind == "oracle":
        return Engine(oracle_lib(), "jo_", world, interaction, disease, outcome_prob, **kw)
    return Engine(_lib.load(), "jb_", world, interaction, disease, outcome_prob, **kw)


def default_beta_table(world: WorldSoA, interaction: Interaction) -> np.ndarray:
    r = world.n_regions
    return interaction.beta_table(world.beta_rules, [1.0] * r, [None] * r)


def simulator_over(engine, base=None):
    """A ``Simulator`` class whose device side is ``engine`` (the CPU oracle, or a CUDA engine the test built itself)
    instead of a ``DeviceWorld`` it creates: the host loop under test is the product's, only the hooks that create the
    device and run the exchange of a multi-domain step are replaced -- with host buffers over gloo for a CPU engine."""
    from june_b200.device import DeviceWorld
    from june_b200.simulator import Simulator
    base = base or Simulator

    class _Over(base):
        def _create_device(self, world, interaction, outcome_prob, device):
            return engine

        def _exchange_buffer_device(self):
            return super()._exchange_buffer_device() if isinstance(engine, DeviceWorld) else None

        def _step_with_exchange(self, params, want_stats):
            if isinstance(engine, DeviceWorld):
                return super()._step_with_exchange(params, want_stats)
            dev, h = self.device, self.halo
            dev.step_begin(params, h.send.data_ptr())
            h.exchange_people()
            dev.step_interact(h.recv.data_ptr(), h.ret_send.data_ptr())
            h.exchange_infections()
            return dev.step_finish(h.ret_recv.data_ptr(), want_stats=want_stats)

    _Over.__name__ = base.__name__
    return _Over
