"""
Downstream node classification -- the accuracy yardstick of the north star.  CPU / sklearn evaluation, restating
tools/graph_model_downstream_classification.py:94-148 of the reference (that module imports hydra + matplotlib and cannot be
imported here): X = input embedding rows 1.. ('<unk>' skipped, :116), optional feature concat (:120-123), labels mapped to
integers, `n_experiments` splits with random_state = i (:134-136), LogisticRegression(**classifier_params) fit on the train
split and scored on the test split (:85-91); mean and best accuracy (:146-148).
"""
from typing import Dict, List, Optional, Tuple

import numpy as np
from sklearn.linear_model import LogisticRegression


def node_classification(embedding: np.ndarray, itos: List[str], labels: Dict[str, str], split_algorithm,
                        n_experiments: int, classifier_params: Optional[dict] = None,
                        features: Optional[Dict[str, np.ndarray]] = None) -> Tuple[float, float]:
    """Returns (mean accuracy, best accuracy) over `n_experiments` splits."""
    X = np.asarray(embedding)[1:, :]
    vertices = itos[1:]
    if features is not None:
        X = np.concatenate([X, np.stack([features[v] for v in vertices])], axis=1)
    classes = {c: i for i, c in enumerate(sorted(set(labels.values())))}
    y = np.array([classes[labels[v]] for v in vertices], dtype=np.float32)
    total, best = 0.0, 0.0
    for i in range(n_experiments):
        split_algorithm.random_state = i
        split = split_algorithm(X, y)
        clf = LogisticRegression(**(classifier_params or {}))
        clf.fit(split['X_train'], split['y_train'])
        acc = float(np.equal(clf.predict(split['X_test']), split['y_test']).astype(np.float32).mean())
        total += acc
        best = max(best, acc)
    return total / n_experiments, best
