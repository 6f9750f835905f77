"""YAML experiment configs -> dataclasses and object factories (see core.py)."""
from .core import GlobalConfig, load_config  # noqa: F401
