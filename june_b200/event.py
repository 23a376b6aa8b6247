"""Events (``june/event/event.py``, ``mutation.py``): things that happen to the world at given dates, applied
at the beginning of a time step (``simulator.py:380-386``)."""
from __future__ import annotations

import datetime
from typing import Dict, List, Optional, Sequence

from .policy import read_date


class Event:
    """``Event`` (``event/event.py:14-49``): active for ``start_time <= date < end_time``."""

    def __init__(self, start_time="1900-01-01", end_time="2100-01-01"):
        self.start_time, self.end_time = read_date(start_time), read_date(end_time)

    def is_active(self, date: datetime.datetime) -> bool:
        return self.start_time <= date < self.end_time

    def initialise(self, world=None) -> None:
        pass

    def apply(self, world, simulator, activities=None, day_type=None) -> None:
        raise NotImplementedError


class Mutation(Event):
    """``Mutation`` (``event/mutation.py:9-66``): while active, at every step a fraction
    ``regional_probabilities[region]`` of the current infections is converted to variant ``mutation``
    (a variant name of ``Simulator.variant_names`` or its index): new transmission, same symptoms."""

    def __init__(self, start_time, end_time, regional_probabilities: Dict[str, float], mutation="B117"):
        super().__init__(start_time, end_time)
        self.regional_probabilities, self.mutation = dict(regional_probabilities), mutation

    def apply(self, world, simulator, activities=None, day_type=None) -> None:
        v = self.mutation if isinstance(self.mutation, int) else simulator.variant_names.index(self.mutation)
        probs = [float(self.regional_probabilities.get(r, 0.0)) for r in world.region_names]
        simulator.device.mutate(v, probs, seed=simulator.seed, step=simulator.step_ordinal)


class Events:
    """``Events`` (``event/event.py:52-104``)."""

    def __init__(self, events: Optional[Sequence[Event]] = None):
        self.events: List[Event] = list(events or [])

    def init_events(self, world=None) -> None:
        for e in self.events:
            e.initialise(world=world)

    def apply(self, date, world, simulator, activities=None, day_type=None) -> None:
        for e in self.events:
            if e.is_active(date):
                e.apply(world=world, simulator=simulator, activities=activities, day_type=day_type)
