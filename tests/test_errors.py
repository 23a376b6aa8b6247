"""Error behaviour at the boundary: what the reference raises, the library reports.

* "Some people do not have an activity" (activity_manager.py:400-402 -> ``SimulatorError``, a ``BaseException``
  as in ``june/exc.py:13``) -> ``JB_ENOACT``;
* a capacity of the device tables exceeded (infection slots) -> ``JB_ECAP``;
* argument errors -> ``JB_EINVAL`` with a message naming the argument."""
import numpy as np
import pytest

from june_b200 import _lib
from june_b200._lib import SimulatorError
from june_b200.disease import DiseaseConfig, outcome_probabilities, synthetic_rates
from june_b200.interaction import Interaction, load_defaults
from june_b200.synth import generate_world

from helpers import default_beta_table, have_gpu, make_engine

FLAGS = _lib.STEP_SEVERE_STAY_HOME | _lib.STEP_HOSPITALISATION


def world_and_engine(kind, n=8000, **kw):
    w = generate_world(n, seed=1, sector_betas=load_defaults()["company_sector_betas"])
    inter, dc = Interaction.from_file(), DiseaseConfig("covid19")
    e = make_engine(kind, w, outcome_probabilities(dc, *synthetic_rates()), interaction=inter, disease=dc, **kw)
    return w, e, default_beta_table(w, inter)


def no_activity_case(kind):
    w, e, beta = world_and_engine(kind)
    assert issubclass(SimulatorError, BaseException) and not issubclass(SimulatorError, Exception)
    # a step whose only activity is the primary one leaves everybody without a job or school nowhere
    with pytest.raises(SimulatorError, match="Some people do not have an activity in this timestep"):
        e.step(e.make_params(0.0, 0.3, 0, 0b00100, seed=1, flags=FLAGS, beta=beta))
    st = e.step(e.make_params(0.3, 0.3, 1, 0b10101, seed=1, flags=FLAGS, beta=beta))
    assert st["n_no_activity"] == 0 and sum(st["n_by_activity"]) == w.n_people
    e.close()


def capacity_case(kind):
    w, e, beta = world_and_engine(kind, slot_capacity=64)
    with pytest.raises(BaseException, match="capacity"):
        e.infect(np.arange(200, dtype=np.uint32), time=0.0, seed=1, step=1)
    # the table holds what fitted, and the world steps on
    st = e.step(e.make_params(0.0, 0.3, 0, 0b10001, seed=1, flags=FLAGS, beta=beta))
    assert st["n_infected"] + st["n_recovered_step"] + st["n_deaths_step"] >= 64
    e.close()


def argument_case(kind):
    w, e, beta = world_and_engine(kind)
    with pytest.raises(BaseException, match="out of range"):
        e.infect(np.array([w.n_people + 5], np.uint32), time=0.0, seed=1, step=1)
    with pytest.raises(BaseException, match="delta_time"):
        e.step(e.make_params(0.0, 0.0, 0, 0b10001, seed=1, flags=FLAGS, beta=beta))
    e.close()


def test_oracle_reports_people_without_an_activity():
    no_activity_case("oracle")


def test_oracle_reports_bad_arguments():
    argument_case("oracle")


@pytest.mark.gpu
def test_cuda_reports_people_without_an_activity():
    assert have_gpu()
    no_activity_case("cuda")


@pytest.mark.gpu
def test_cuda_reports_exceeded_capacity_and_bad_arguments():
    assert have_gpu()
    capacity_case("cuda")
    argument_case("cuda")
