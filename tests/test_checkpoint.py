"""Checkpoint / resume (hdf5_savers/checkpoint_saver.py:39-218,
simulator.py:676-735): a run resumed from a checkpoint continues bit for bit."""
import datetime
import os

import numpy as np
import pytest

from june_b200.interaction import load_defaults
from june_b200.policy import Policies
from june_b200.synth import generate_world

from helpers import have_gpu
from test_leisure_policies import make_sim


def trace(sim, n_steps=None):
    out = []
    while sim.timer.date < sim.timer.final_date and (n_steps is None or len(out) < n_steps):
        if sim.epidemiology is not None:
            sim.epidemiology.infection_seeds_timestep(sim, sim.timer)
        sim.do_timestep()
        h = sim.history[-1]
        out.append((h["n_new_infections"], h["n_infected"], h["n_hospitalised"], h["n_icu"], h["n_dead_total"],
                    h["n_recovered_step"], tuple(h["n_by_activity"])))
        next(sim.timer)
    return out


def resume_case(kind, tmp_path):
    w = generate_world(10_000, seed=8, leisure=True, sector_betas=load_defaults()["company_sector_betas"])
    pol = lambda: Policies.from_file()
    a = make_sim(kind, w, pol(), total_days=18, seed=5)
    a.epidemiology.infection_seeds[0].cases_per_capita = 0.05
    first = trace(a, 57)          # stops in the middle of a day
    path = a.save_checkpoint(None, path=os.path.join(tmp_path, "cp.npz"))
    state_a = a.device.fetch_person_state()
    rest_a = trace(a)
    b = make_sim(kind, w, pol(), total_days=18, seed=5)
    b.restore_checkpoint(path)
    state_b = b.device.fetch_person_state()
    for k in state_a:  # bits 0-2 of the state byte are the activity of the last step, not state
        assert np.array_equal(state_a[k] & 0x78 if k == "pstate" else state_a[k], state_b[k] & 0x78 if k == "pstate" else state_b[k]), k
    assert b.timer.date == datetime.datetime(2020, 3, 2) + datetime.timedelta(days=b.timer.now) and b.step_ordinal == 57
    rest_b = trace(b)
    assert rest_a == rest_b, "the resumed run must continue exactly"
    assert first[-1][4] + first[-1][2] > 0 and rest_a[-1][4] > first[-1][4], "deaths / patients on both sides of the checkpoint"
    fa, fb = a.device.fetch_person_state(), b.device.fetch_person_state()
    for k in fa:
        assert np.array_equal(fa[k], fb[k]), k
    return path, w


def test_oracle_resumes_bit_for_bit(tmp_path):
    path, w = resume_case("oracle", str(tmp_path))
    z = np.load(path)
    # the reference's dataset names (checkpoint_saver.py:56-82, infection_savers/*.py)
    for key in ("people_data/people_id", "people_data/infected_id", "people_data/dead_id", "infections/transmissions/shape",
                "infections/transmissions/shift", "infections/transmissions/scale", "infections/transmissions/norm",
                "infections/symptoms/max_tag", "infections/symptoms/tag", "infections/symptoms/stage",
                "infections/symptoms/trajectory_times", "infections/symptoms/trajectory_symptoms",
                "infections/symptoms/trajectory_lengths", "immunities/susceptibilities", "time/date"):
        assert key in z.files, key
    assert len(z["people_data/infected_id"]) == len(z["infections/start_times"]) > 100
    assert int(z["infections/symptoms/trajectory_lengths"].sum()) == len(z["infections/symptoms/trajectory_times"])


def test_reference_style_restore_restarts_the_next_day_and_can_drop_infections(tmp_path):
    w = generate_world(8_000, seed=8, sector_betas=load_defaults()["company_sector_betas"], leisure=True)
    a = make_sim("oracle", w, Policies.from_file(), total_days=12, seed=5)
    a.checkpoint_save_dates = (datetime.date(2020, 3, 5),)
    a.checkpoint_save_path = str(tmp_path)
    a.epidemiology.infection_seeds[0].cases_per_capita = 0.05
    a.run()
    path = os.path.join(str(tmp_path), "checkpoint_2020-03-05.0.npz")
    assert os.path.exists(path), "Simulator.run writes the checkpoint after the day's last step"
    b = make_sim("oracle", w, Policies.from_file(), total_days=12, seed=5)
    b.restore_checkpoint(path, exact=False)
    assert b.timer.date == datetime.datetime(2020, 3, 6) and b.timer.shift == 0
    n_inf = int(((b.device.fetch_person_state()["pstate"] & 0x08) != 0).sum())
    assert n_inf == len(np.load(path)["people_data/infected_id"]) > 0
    c = make_sim("oracle", w, Policies.from_file(), total_days=12, seed=5)
    c.restore_checkpoint(path, reset_infections=True)
    st = c.device.fetch_person_state()
    assert not (st["pstate"] & 0x08).any() and (st["medical"] < 0).all()
    assert ((st["pstate"] & 0x10) != 0).sum() == len(np.load(path)["people_data/dead_id"])
    assert (st["susceptibility"][0] == 0).sum() >= n_inf, "immunities survive a reset of the infections"


def run_written_case(kind, tmp_path, vaccination):
    """The checkpoint ``Simulator.run`` writes itself (after the day's last step, before the clock advances),
    restored with the defaults: the continuation must equal the uninterrupted run step for step."""
    from test_vaccination import campaigns
    w = generate_world(8_000, seed=8, sector_betas=load_defaults()["company_sector_betas"], leisure=True)

    def build():
        s = make_sim(kind, w, Policies.from_file(), total_days=12, seed=5)
        s.epidemiology.infection_seeds[0].cases_per_capita = 0.05
        if vaccination:
            s.epidemiology.vaccination_campaigns = campaigns()
        return s
    a = build()
    a.checkpoint_save_dates = (datetime.date(2020, 3, 6),)
    a.checkpoint_save_path = str(tmp_path)
    a.run()
    path = os.path.join(str(tmp_path), "checkpoint_2020-03-06.0.npz")
    z = np.load(path)
    assert int(z["time/step_executed"]) == 1
    b = build()
    b.restore_checkpoint(path)
    assert b.timer.date == datetime.datetime(2020, 3, 7) and b.timer.shift == 0, "continues with the step AFTER the saved one"
    n_done = int(z["time/step_ordinal"])
    assert b.step_ordinal == n_done
    b.run()
    key = lambda h: (h["date"], h["n_new_infections"], h["n_infected"], h["n_hospitalised"], h["n_dead_total"], tuple(h["n_by_activity"]))
    assert [key(h) for h in a.history[n_done:]] == [key(h) for h in b.history]
    fa, fb = a.device.fetch_person_state(), b.device.fetch_person_state()
    for k in fa:
        assert np.array_equal(fa[k], fb[k]), k
    if vaccination:
        va, vb = a.device.fetch_vaccination_state(), b.device.fetch_vaccination_state()
        for k in va:
            assert np.array_equal(va[k], vb[k]), k
        assert (va["dose"] >= 0).sum() > 500 and (np.load(path)["vaccination/dose"] >= 0).sum() > 100


@pytest.mark.parametrize("vaccination", [False, True])
def test_oracle_resumes_the_checkpoint_run_writes(tmp_path, vaccination):
    run_written_case("oracle", tmp_path, vaccination)


@pytest.mark.gpu
def test_cuda_resumes_the_checkpoint_run_writes_with_vaccination(tmp_path):
    assert have_gpu()
    run_written_case("cuda", tmp_path, True)


@pytest.mark.gpu
def test_cuda_resumes_bit_for_bit(tmp_path):
    assert have_gpu()
    resume_case("cuda", str(tmp_path))
