"""Running-mean metric meter (reference: shallow_encoders/word2vec/utils/meter.py:17-83)."""
from collections import defaultdict
from typing import Iterable, Tuple


class UnknownMetricException(KeyError):
    """No value was pushed under this name."""


class MetricMeter:
    def __init__(self):
        self._sum, self._count = defaultdict(float), defaultdict(int)

    @property
    def is_empty(self) -> bool:
        return not self._count

    def push(self, name: str, value) -> None:
        self._sum[name] += float(value)
        self._count[name] += 1

    def get(self, name: str) -> float:
        if name not in self._count:
            raise UnknownMetricException(f'Metric name "{name}" not found. Known metrics: {list(self._count)}.')
        return self._sum[name] / self._count[name]

    def get_all(self, flush: bool = True) -> Iterable[Tuple[str, float]]:
        items = [(name, self.get(name)) for name in self._count]
        if flush:
            self.flush()
        return items

    def flush(self) -> None:
        self._sum.clear()
        self._count.clear()
