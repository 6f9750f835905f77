"""Train / validation / test splits used by the downstream node- and edge-classification yardstick
(reference: shallow_encoders/split/core.py:11-199; YAML `_target_: shallow_encoders.split.TrainTestRatioSplit`).
Host-side evaluation plumbing around sklearn, not part of the accelerated path.  Every algorithm maps (X, y) to a dict of copies
keyed 'X_train', 'y_train', ['X_val', 'y_val',] 'X_test', 'y_test' and reproduces the reference's draws for a given `random_state`
(default 42): sklearn's `train_test_split` for the ratio splits, numpy's legacy global-seed shuffle for the per-class sample split."""
from typing import Dict, Optional

import numpy as np
from sklearn.model_selection import train_test_split


def _copies(**parts) -> Dict[str, np.ndarray]:
    return {name: np.array(value, copy=True) for name, value in parts.items()}


class SplitAlgorithm:
    """Base: a seed (`random_state`, settable between experiments as the downstream tool does) and `algo(X, y) == algo.split(X, y)`."""

    def __init__(self, random_state: Optional[int] = None):
        self.random_state = 42 if random_state is None else random_state

    def split(self, X: np.ndarray, y: np.ndarray) -> Dict[str, np.ndarray]:
        raise NotImplementedError

    def __call__(self, X: np.ndarray, y: np.ndarray) -> Dict[str, np.ndarray]:
        return self.split(X, y)

    def _two_way(self, X, y, held_out: float, stratify: bool):
        return train_test_split(X, y, test_size=held_out, stratify=y if stratify else None, random_state=self.random_state)


class TrainTestRatioSplit(SplitAlgorithm):
    """`train_ratio` of the samples for training (optionally stratified); `test_all` evaluates on the WHOLE data set (core.py:48-78)."""

    def __init__(self, train_ratio: float, stratify: bool = False, test_all: bool = False, random_state: Optional[int] = None):
        super().__init__(random_state)
        self._train_ratio, self._stratify, self._test_all = train_ratio, stratify, test_all

    def split(self, X: np.ndarray, y: np.ndarray) -> Dict[str, np.ndarray]:
        x_tr, x_te, y_tr, y_te = self._two_way(X, y, 1 - self._train_ratio, self._stratify)
        if self._test_all:
            x_te, y_te = X, y
        return _copies(X_train=x_tr, y_train=y_tr, X_test=x_te, y_test=y_te)


class TrainValTestRatioSplit(SplitAlgorithm):
    """Two nested ratio splits with the same seed: train vs rest, then the rest into validation and test, where the test share of the
    rest is (1 - val_ratio) / (1 - train_ratio) exactly as the reference computes it (core.py:81-120)."""

    def __init__(self, train_ratio: float, val_ratio: float, stratify: bool = False, random_state: Optional[int] = None):
        super().__init__(random_state)
        self._train_ratio, self._val_ratio, self._stratify = train_ratio, val_ratio, stratify

    def split(self, X: np.ndarray, y: np.ndarray) -> Dict[str, np.ndarray]:
        x_tr, x_rest, y_tr, y_rest = self._two_way(X, y, 1 - self._train_ratio, self._stratify)
        x_va, x_te, y_va, y_te = self._two_way(x_rest, y_rest, (1 - self._val_ratio) / (1 - self._train_ratio), self._stratify)
        return _copies(X_train=x_tr, y_train=y_tr, X_val=x_va, y_val=y_va, X_test=x_te, y_test=y_te)


class TrainValTestStratifiedNSamplesSplit(SplitAlgorithm):
    """Per class (ascending label order): shuffle the class's indices, take `train_samples`, then `val_samples`, then `test_samples`
    (or all that is left) -- the Planetoid-style split (core.py:123-199).  The shuffles consume numpy's legacy generator seeded once with
    `random_state`, so the picks equal the reference's `np.random.seed` + `np.random.shuffle` sequence."""

    def __init__(self, train_samples: int, val_samples: int, test_samples: Optional[int] = None, random_state: Optional[int] = None):
        super().__init__(random_state)
        self._n_train, self._n_val, self._n_test = train_samples, val_samples, test_samples

    def split(self, X: np.ndarray, y: np.ndarray) -> Dict[str, np.ndarray]:
        rng = np.random.RandomState(self.random_state)
        picks = {'train': [], 'val': [], 'test': []}
        classes = np.unique(y)
        for label in classes:
            members = np.flatnonzero(y == label)
            rng.shuffle(members)
            a, b = self._n_train, self._n_train + self._n_val
            c = None if self._n_test is None else b + self._n_test
            picks['train'] += members[:a].tolist()
            picks['val'] += members[a:b].tolist()
            picks['test'] += members[b:c].tolist()
        wanted = {'train': self._n_train, 'val': self._n_val, 'test': self._n_test}
        for part, per_class in wanted.items():
            if per_class is not None:
                assert len(picks[part]) == len(classes) * per_class, f'{len(picks[part])} != {len(classes) * per_class}'
        return _copies(X_train=X[picks['train']], y_train=y[picks['train']], X_val=X[picks['val']], y_val=y[picks['val']],
                       X_test=X[picks['test']], y_test=y[picks['test']])
