"""Checkpoint / resume of the device-resident epidemic state.

Mirrors ``hdf5_savers/checkpoint_saver.py`` (``save_checkpoint_to_hdf5`` :39-82,
``restore_simulator_to_checkpoint`` :162-218): ``people_data/{people_id, infected_id, dead_id}``, the
infections (``infection_savers/transmission_saver.py``: gamma ``shape, shift, scale, norm`` / xnexp
``n, alpha, time_first_infectious, norm, norm_time`` / constant ``probability``;
``symptoms_saver.py``: ``max_tag, tag, max_severity, stage, time_of_symptoms_onset`` and the variable
length trajectories as ``trajectory_times / trajectory_symptoms / trajectory_lengths``) and the
immunities (``immunity_saver.py``: susceptibilities).  h5py is not available in this environment, so the
container is a NumPy ``.npz`` with the reference's dataset names as keys ("group/dataset").

Additions that make a resumed run continue bit for bit instead of "from the next day" as the reference does:
the clock (``time/shift``, ``time/now``, ``time/step_ordinal`` and ``time/step_executed``: ``Simulator.run`` writes
its checkpoints after the day's last ``do_timestep`` and BEFORE the clock advances, so a restore has to advance it),
``people_data/medical`` (``person.subgroups.medical_facility``, which the reference drops and re-creates by
re-admission) and the per-person vaccination state (``vaccination/*``: dose, vaccine, running trajectory, prior
efficacies, effective multipliers -- the reference pickles them with the person)."""
from __future__ import annotations

import datetime
from typing import Dict

import numpy as np

from .world import MAX_STAGES

PS_DEAD = 0x10
TRANSMISSION_COLUMNS = {"gamma": ("shape", "shift", "scale", "norm"),
                        "xnexp": ("n", "alpha", "time_first_infectious", "norm", "norm_time"),
                        "constant": ("probability",)}


def checkpoint_arrays(sim) -> Dict[str, np.ndarray]:
    dev, w = sim.device, sim.world
    st = dev.fetch_person_state()
    rows = sorted(dev.fetch_infections(), key=lambda r: r["person"])
    pid = w.person_ids if w.person_ids is not None else np.arange(w.n_people, dtype=np.int64) + w.person_id_base
    n = len(rows)
    ttype = sim.disease_config.transmission["type"]
    out: Dict[str, np.ndarray] = {
        "time/date": np.array(sim.timer.date.strftime("%Y-%m-%d %H:%M:%S")),
        "time/shift": np.array(sim.timer.shift), "time/now": np.array(sim.timer.now),
        "time/step_ordinal": np.array(sim.step_ordinal),
        # the step of the saved clock has already been executed (checkpoint written between do_timestep and next(timer))
        "time/step_executed": np.array(int(getattr(sim, "_last_step_date", None) == sim.timer.date)),
        "people_data/people_id": np.asarray(pid, np.int64),
        "people_data/infected_id": np.asarray([pid[r["person"]] for r in rows], np.int64),
        "people_data/dead_id": np.asarray(pid[(st["pstate"] & PS_DEAD) != 0], np.int64),
        "people_data/medical": st["medical"].astype(np.int32),
        "immunities/susceptibilities": st["susceptibility"],
        "infections/infection_ids": np.asarray([r["variant"] for r in rows], np.int64),
        "infections/start_times": np.asarray([r["start_time"] for r in rows], np.float64),
        "infections/transmissions/transmission_type": np.array(ttype),
        "infections/transmissions/probability": np.asarray([r["probability"] for r in rows], np.float64),
        "infections/symptoms/max_tag": np.asarray([r["max_tag"] for r in rows], np.int64),
        "infections/symptoms/tag": np.asarray([r["tag"] for r in rows], np.int64),
        "infections/symptoms/stage": np.asarray([r["stage"] for r in rows], np.int64),
        "infections/symptoms/trajectory_lengths": np.asarray([len(r["traj_time"]) for r in rows], np.int64),
        "infections/symptoms/trajectory_times": np.asarray([t for r in rows for t in r["traj_time"]], np.float64),
        "infections/symptoms/trajectory_symptoms": np.asarray([t for r in rows for t in r["traj_tag"]], np.int64),
    }
    for k, name in enumerate(TRANSMISSION_COLUMNS[ttype]):
        out[f"infections/transmissions/{name}"] = np.asarray([r["tpar"][k] for r in rows], np.float64).reshape(n)
    ep = getattr(sim, "epidemiology", None)
    if ep is not None and getattr(ep, "vaccination_campaigns", None) is not None and getattr(sim, "_vaccination_on_device", False):
        for k, v in dev.fetch_vaccination_state().items():
            out[f"vaccination/{k}"] = v
        out["vaccination/campaign_day0"] = np.array(sim._vaccination_first_day.isoformat())
        out["vaccination/current_date"] = np.array("" if ep.current_date is None else ep.current_date.strftime("%Y-%m-%d %H:%M:%S"))
    return out


def save_checkpoint(sim, path: str) -> None:
    """``save_checkpoint_to_hdf5``; called by ``Simulator.run`` on ``checkpoint_save_dates``
    (``simulator.py:676-690``) or by hand."""
    np.savez_compressed(path, **checkpoint_arrays(sim))


def restore_checkpoint(sim, path: str, reset_infections: bool = False, exact: bool = True) -> None:
    """``restore_simulator_to_checkpoint`` onto a freshly built simulator of the same world.
    ``exact=False`` reproduces the reference's clock handling (restart at 00:00 of the day after the
    checkpoint, ``checkpoint_saver.py:211-215``); ``exact=True`` continues at the saved step."""
    z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
    dev, w = sim.device, sim.world
    pid = w.person_ids if w.person_ids is not None else np.arange(w.n_people, dtype=np.int64) + w.person_id_base
    index = {int(v): i for i, v in enumerate(pid)}
    ours = np.array([index.get(int(v), -1) for v in z["people_data/people_id"]])
    # dead people and immunities (people of other domains are skipped, as in the reference)
    pstate = np.zeros(w.n_people, np.uint8)
    for v in z["people_data/dead_id"]:
        if int(v) in index:
            pstate[index[int(v)]] = PS_DEAD
    susc = dev.fetch_person_state()["susceptibility"]
    keep = ours >= 0
    susc[:, ours[keep]] = z["immunities/susceptibilities"][:, keep]
    dev.set_person_state(pstate, susc)
    medical = np.full(w.n_people, -1, np.int32)
    medical[ours[keep]] = z["people_data/medical"][keep]
    if reset_infections:
        medical[:] = -1
    dev.set_medical(medical)
    if not reset_infections:
        ttype = str(z["infections/transmissions/transmission_type"])
        cols = TRANSMISSION_COLUMNS[ttype]
        lens = z["infections/symptoms/trajectory_lengths"]
        offs = np.concatenate([[0], np.cumsum(lens)])
        rows = []
        for i, v in enumerate(z["people_data/infected_id"]):
            if int(v) not in index:
                continue
            lo, hi = int(offs[i]), int(offs[i + 1])
            assert hi - lo <= MAX_STAGES
            rows.append(dict(person=index[int(v)], variant=int(z["infections/infection_ids"][i]),
                             start_time=float(z["infections/start_times"][i]),
                             tpar=[float(z[f"infections/transmissions/{c}"][i]) for c in cols],
                             probability=float(z["infections/transmissions/probability"][i]),
                             stage=int(z["infections/symptoms/stage"][i]), tag=int(z["infections/symptoms/tag"][i]),
                             max_tag=int(z["infections/symptoms/max_tag"][i]),
                             traj_time=z["infections/symptoms/trajectory_times"][lo:hi].tolist(),
                             traj_tag=z["infections/symptoms/trajectory_symptoms"][lo:hi].tolist()))
        if rows:
            dev.upload_infections(rows)
    if "vaccination/dose" in z.files and not reset_infections:
        ep = sim.epidemiology
        if ep is None or ep.vaccination_campaigns is None:
            raise ValueError("the checkpoint holds a vaccination state, the simulator has no vaccination campaigns")
        first = datetime.date.fromisoformat(str(z["vaccination/campaign_day0"]))
        sim._vaccination_first_day = first
        dev.set_vaccination(*ep.vaccination_campaigns.device_tables(first, sim.variant_names))
        sim._vaccination_on_device = True
        state = {}
        for k in dev.fetch_vaccination_state():
            a = z[f"vaccination/{k}"]
            cur = dev.fetch_vaccination_state()[k] if not keep.all() else None
            if a.ndim == 2:
                full = cur if cur is not None else np.zeros((a.shape[0], w.n_people), a.dtype)
                full[:, ours[keep]] = a[:, keep]
            else:
                full = cur if cur is not None else np.zeros(w.n_people, a.dtype)
                full[ours[keep]] = a[keep]
            state[k] = full
        dev.set_vaccination_state(state)
        cd = str(z["vaccination/current_date"])
        ep.current_date = datetime.datetime.strptime(cd, "%Y-%m-%d %H:%M:%S") if cd else None
        if getattr(sim, "record", None) is not None:
            sim._vaccination_dose_prev = state["dose"].copy()
    date = datetime.datetime.strptime(str(z["time/date"]), "%Y-%m-%d %H:%M:%S")
    if exact:
        sim.timer.date = date
        sim.timer.previous_date = date
        sim.timer.shift = int(z["time/shift"])
        sim.timer.delta_time = datetime.timedelta(hours=sim.timer.shift_duration)
        sim.step_ordinal = int(z["time/step_ordinal"])
        if "time/step_executed" in z.files and int(z["time/step_executed"]):
            next(sim.timer)  # the saved step has run: continue with the one after it
    else:
        sim.timer.reset_to_new_date(datetime.datetime(date.year, date.month, date.day) + datetime.timedelta(days=1))
        sim.step_ordinal = int(z["time/step_ordinal"])
