"""Drop-in `vsrlab` package for the Real-BasicVSR / BasicVSR hot path.

Mirrors the class paths, constructor arguments, parameter names and forward
signatures of santurini/vsrlab (`src/` is installed as package `vsrlab` by the
reference's setup.py:6-7) so that Hydra `_target_` strings such as
`vsrlab.vsr.models.RealBasicVSR.realbasicvsr.RealBasicVSR`
(reference conf/train/model/basicvsr.yaml:1) resolve to the B200-native
implementation.  All compute goes through the C-ABI in `vsrlab_b200`; there is
no CPU path.
"""
import os
from pathlib import Path

# reference src/core/__init__.py:8-14 exports these on import of vsrlab.core;
# kept so scripts that read them keep working.
PROJECT_ROOT = Path(os.environ.get("PROJECT_ROOT", Path.cwd().parents[0] if len(Path.cwd().parents) else Path.cwd()))

# The reference's caller modules (training runtime, datasets, losses, scripts' helpers) resolve underneath the drop-in
# when VSRLAB_REFERENCE_SRC points at a tree with the reference's src/ layout (see _overlay.py).
if os.environ.get("VSRLAB_REFERENCE_SRC"):
    from ._overlay import install as _install_reference_overlay
    _install_reference_overlay(os.environ["VSRLAB_REFERENCE_SRC"])
