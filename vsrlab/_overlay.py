"""Overlay of the reference's caller modules underneath the drop-in package.

The drop-in `vsrlab` replaces the HOT PATH modules of santurini/vsrlab only.  The reference's training /
evaluation runtime (`vsrlab.core.utils`, `.losses`, `.metrics`, `.loggers`, `vsrlab.vsr.dataset`, `vsrlab.train`, ...) is
the caller of that path and must keep working unchanged; it is not re-implemented here.  When the environment variable
`VSRLAB_REFERENCE_SRC` names a tree with the reference's `src/` layout, every `vsrlab.*` module the drop-in does not
provide itself is loaded from that tree - from `module.py` sources or from byte-compiled `module.bc` files - so that
`python train.py +experiment=basic` runs the reference's own loop on the B200-native model.  Modules the drop-in does
provide always win."""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import sys
from pathlib import Path


class _ReferenceFinder(importlib.abc.MetaPathFinder):
    def __init__(self, root: Path):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith("vsrlab."):
            return None
        rel = Path(*fullname.split(".")[1:])
        for suffix, loader_cls in ((".py", importlib.machinery.SourceFileLoader), (".bc", importlib.machinery.SourcelessFileLoader)):
            pkg, mod = self.root / rel / ("__init__" + suffix), (self.root / rel).with_suffix(suffix)
            if pkg.exists():
                return importlib.util.spec_from_file_location(fullname, str(pkg), loader=loader_cls(fullname, str(pkg)),
                                                              submodule_search_locations=[str(pkg.parent)])
            if mod.exists():
                return importlib.util.spec_from_file_location(fullname, str(mod), loader=loader_cls(fullname, str(mod)))
        return None


def _compat_patches() -> None:
    """API drift between the reference's (unpinned, requirements.txt) dependencies and this image, patched in the running
    process only - the reference's files stay untouched:
    * core/loggers.py:45,48,56 call `torchvision.utils.make_grid(..., ncol=1)`; current torchvision dropped the `**kwargs`
      that used to swallow the unknown `ncol`."""
    import functools
    import inspect

    import torchvision.utils as tvu
    if "ncol" not in inspect.signature(tvu.make_grid).parameters and not getattr(tvu.make_grid, "_vsrlab_compat", False):
        orig = tvu.make_grid

        @functools.wraps(orig)
        def make_grid(*args, ncol=None, **kwargs):
            return orig(*args, **kwargs)
        make_grid._vsrlab_compat = True
        tvu.make_grid = make_grid


def install(reference_src: str) -> None:
    root = Path(reference_src)
    if not root.is_dir():
        raise ImportError(f"VSRLAB_REFERENCE_SRC={reference_src!r} is not a directory")
    if not any(isinstance(f, _ReferenceFinder) and f.root == root for f in sys.meta_path):
        sys.meta_path.append(_ReferenceFinder(root))        # last: the drop-in's own modules are found first
    _compat_patches()
