"""Simulation clock with the public surface of the reference ``Timer``
(``june/time.py:10-161``): ``date``, ``now`` (days since start), ``duration`` (days),
``activities``, ``shift``, ``day_type``, ``is_weekend``, ``final_date``, ``next(timer)``.

The schedule is expanded once into per-day-type tables instead of being looked up through
``getattr`` on every access."""
from __future__ import annotations

import calendar
import datetime
from typing import Dict, List, Optional, Sequence

import yaml

SECONDS_PER_DAY = 24 * 60 * 60


def _as_list(x) -> list:
    # YAML gives {0: ..., 1: ...} (config_simulation.yaml:17-41) or plain lists
    if isinstance(x, dict):
        return [x[k] for k in sorted(x)]
    return list(x)


class Timer:
    def __init__(self, initial_day: str = "2020-03-01 9:00", total_days: int = 10,
                 weekday_step_duration: Sequence[int] = (12, 12), weekend_step_duration: Sequence[int] = (24,),
                 weekday_activities: Sequence[Sequence[str]] = (("primary_activity", "residence"), ("residence",)),
                 weekend_activities: Sequence[Sequence[str]] = (("residence",),),
                 day_types: Optional[Dict[str, List[str]]] = None):
        parts = initial_day.split(" ")
        y, m, d = (int(v) for v in parts[0].split("-"))
        hour = int(parts[1].split(":")[0]) if len(parts) > 1 else 0
        self.initial_date = datetime.datetime(y, m, d) + datetime.timedelta(hours=hour)
        self.day_types = day_types or {
            "weekday": ["Monday", "Tuesday", "Wednesday", "Thursday", "Friday"],
            "weekend": ["Saturday", "Sunday"],
        }
        self.total_days = total_days
        self._dur = {"weekday": [float(x) for x in _as_list(weekday_step_duration)],
                     "weekend": [float(x) for x in _as_list(weekend_step_duration)]}
        self._act = {"weekday": [list(a) for a in _as_list(weekday_activities)],
                     "weekend": [list(a) for a in _as_list(weekend_activities)]}
        self.final_date = self.initial_date + datetime.timedelta(days=total_days)
        self.reset()

    # reference attribute names
    @property
    def weekday_step_duration(self):
        return self._dur["weekday"]

    @property
    def weekend_step_duration(self):
        return self._dur["weekend"]

    @property
    def weekday_activities(self):
        return self._act["weekday"]

    @property
    def weekend_activities(self):
        return self._act["weekend"]

    @classmethod
    def from_file(cls, config_filename: str) -> "Timer":
        with open(config_filename) as f:
            config = yaml.safe_load(f)
        return cls.from_config(config)

    @classmethod
    def from_config(cls, config: dict) -> "Timer":
        t = config["time"]
        day_types = None
        if "weekday" in config and "weekend" in config:
            day_types = {"weekday": config["weekday"], "weekend": config["weekend"]}
        return cls(initial_day=t["initial_day"], total_days=t["total_days"],
                   weekday_step_duration=t["step_duration"]["weekday"],
                   weekend_step_duration=t["step_duration"]["weekend"],
                   weekday_activities=t["step_activities"]["weekday"],
                   weekend_activities=t["step_activities"]["weekend"], day_types=day_types)

    _DAY_NAMES = ("Monday", "Tuesday", "Wednesday", "Thursday", "Friday", "Saturday", "Sunday")

    # `date` is assigned by reset / __next__ / a checkpoint restore: everything derived from the calendar
    # day is refreshed there once instead of on each of the several reads a time step makes
    @property
    def date(self) -> datetime.datetime:
        return self._date

    @date.setter
    def date(self, value: datetime.datetime) -> None:
        self._date = value
        self._day_of_week = self._DAY_NAMES[value.weekday()]  # calendar.day_name goes through strftime
        self._is_weekend = self._day_of_week in self.day_types["weekend"]
        self._day_type = "weekend" if self._is_weekend else "weekday"

    @property
    def day_of_week(self) -> str:
        return self._day_of_week

    @property
    def is_weekend(self) -> bool:
        return self._is_weekend

    @property
    def day_type(self) -> str:
        return self._day_type

    @property
    def now(self) -> float:
        return (self.date - self.initial_date).total_seconds() / SECONDS_PER_DAY

    @property
    def date_str(self) -> str:
        return self.date.date().strftime("%Y-%m-%d")

    @property
    def duration(self) -> float:
        return self.delta_time.total_seconds() / SECONDS_PER_DAY

    @property
    def day(self) -> int:
        return int(self.now)

    @property
    def activities(self) -> List[str]:
        return self._act[self.day_type][self.shift]

    @property
    def shift_duration(self) -> float:
        return self._dur[self.day_type][self.shift]

    def reset(self) -> None:
        self.date = self.initial_date
        self.previous_date = self.initial_date
        self.shift = 0
        self.delta_time = datetime.timedelta(hours=self.shift_duration)

    def reset_to_new_date(self, date) -> None:
        self.date = date
        self.shift = 0
        self.delta_time = datetime.timedelta(hours=self.shift_duration)
        self.previous_date = self.initial_date

    def __next__(self):
        self.previous_date = self.date
        self.date = self.date + self.delta_time
        self.shift = 0 if self.previous_date.day != self.date.day else self.shift + 1
        self.delta_time = datetime.timedelta(hours=self.shift_duration)
        return self.date
