"""Project paths (mirrors shallow_encoders/common/path.py:8-11 of the reference: ROOT / CONFIG / RUNS / ASSETS)."""
import os
from pathlib import Path

ROOT_PATH = os.environ.get('SE_ROOT_PATH', str(Path(__file__).resolve().parent.parent.parent))
CONFIG_PATH = os.path.join(ROOT_PATH, 'configs')
RUNS_PATH = os.path.join(ROOT_PATH, 'runs')
ASSETS_PATH = os.path.join(ROOT_PATH, 'assets')
