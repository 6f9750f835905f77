"""Name -> dataset class registry; same contract as the reference's
shallow_encoders/word2vec/dataloader/registry.py:6-26 (`DATASET_REGISTRY[name](**additional_parameters)`)."""
from typing import Callable, Dict, Type

DATASET_REGISTRY: Dict[str, Type] = {}


def register_dataset(name: str) -> Callable[[Type], Type]:
    """Class decorator adding the dataset to DATASET_REGISTRY; registering a name twice is an error."""
    assert name not in DATASET_REGISTRY, f'Already registered "{name}"!'

    def wrap(cls: Type) -> Type:
        DATASET_REGISTRY[name] = cls
        return cls

    return wrap
