import os
from pathlib import Path

# same environment contract as reference src/core/__init__.py:8-14
PROJECT_ROOT = Path.cwd().parents[0] if len(Path.cwd().parents) else Path.cwd()
LOGGING_DIR = Path(PROJECT_ROOT / "storage")
os.environ.setdefault("PROJECT_ROOT", str(PROJECT_ROOT))
os.environ.setdefault("LOGGING_DIR", str(LOGGING_DIR))
os.environ.setdefault("HYDRA_FULL_ERROR", "1")
