"""Time granularity of a temporal graph.  Mirrors tgm/core/timedelta.py:9-112 (same fields,
predicates and error behaviour); only what the loader needs for time-unit batching."""
from __future__ import annotations

from dataclasses import dataclass

from tgm_b200.exceptions import EventOrderedConversionError

_NS = {'ns': 1, 'us': 10**3, 'ms': 10**6, 's': 10**9, 'm': 60 * 10**9, 'h': 3600 * 10**9,
       'D': 86400 * 10**9, 'W': 7 * 86400 * 10**9, 'M': 30 * 86400 * 10**9,
       'Y': 365 * 86400 * 10**9}
_EVENT = 'r'


@dataclass(frozen=True)
class TimeDeltaDG:
    unit: str
    value: int = 1

    def __post_init__(self) -> None:
        if not isinstance(self.value, int) or isinstance(self.value, bool) or self.value <= 0:
            raise ValueError(f'Value must be a positive integer, got: {self.value}')
        if self.unit == _EVENT:
            if self.value != 1:
                raise ValueError('Only value=1 is supported for event-ordered TimeDeltaDG')
        elif self.unit not in _NS:
            raise ValueError(f'Unknown unit: {self.unit}, expected one of {[_EVENT, *_NS]}')

    @property
    def is_event_ordered(self) -> bool:
        return self.unit == _EVENT

    @property
    def is_time_ordered(self) -> bool:
        return self.unit != _EVENT

    def convert(self, other: 'str | TimeDeltaDG') -> float:
        """How many `other` ticks fit in one tick of self."""
        if isinstance(other, str):
            other = TimeDeltaDG(other)
        if self.is_event_ordered or other.is_event_ordered:
            raise EventOrderedConversionError(
                'Cannot compare granularity for event-ordered TimeDeltaDG')
        a, b = _NS[self.unit], _NS[other.unit]
        ratio = self.value / other.value
        return ratio * (a // b) if a > b else ratio / (b // a)

    def is_coarser_than(self, other: 'str | TimeDeltaDG') -> bool:
        return self.convert(other) > 1
