"""Train/test split used by the downstream node-classification yardstick (reference: shallow_encoders/split/core.py:48-78).
CPU / sklearn evaluation code, not part of the accelerated path."""
from typing import Dict, Optional

import numpy as np
from sklearn.model_selection import train_test_split


class TrainTestRatioSplit:
    def __init__(self, train_ratio: float, stratify: bool = False, test_all: bool = False, random_state: Optional[int] = None):
        self.random_state = 42 if random_state is None else random_state
        self._train_ratio, self._stratify, self._test_all = train_ratio, stratify, test_all

    def split(self, X: np.ndarray, y: np.ndarray) -> Dict[str, np.ndarray]:
        x_tr, x_te, y_tr, y_te = train_test_split(X, y, test_size=1 - self._train_ratio,
                                                  stratify=y if self._stratify else None, random_state=self.random_state)
        if self._test_all:
            x_te, y_te = X, y
        return {'X_train': x_tr.copy(), 'y_train': y_tr.copy(), 'X_test': x_te.copy(), 'y_test': y_te.copy()}

    __call__ = split
