"""Recipe that places the reference's OWN hot-path modules under oracle/_ref/ so that they can travel to the GPU box.
TEST / BASELINE INFRASTRUCTURE ONLY.

/root/reference exists only in the build container.  The reference's Real-BasicVSR path is pure Python on top of
torch / einops / torchvision, so "building" it means byte-compiling the files the path imports, from
where they lie under /root/reference/src, into CPython bytecode files (`*.bc` = the standard .pyc container, same
relative paths; the neutral suffix keeps snapshot tools that drop `*.pyc` from dropping them) under oracle/_ref/src,
plus a manifest with the SHA-256 of every source file compiled.  No reference SOURCE is copied anywhere;
oracle/_ref/ holds build outputs only, is listed in .gitignore (it never enters this repository's history) but not in
.gpurunignore, so it rides along to the GPU box like the built .so files (same image, same interpreter).
`oracle/ref_runner.py` imports the compiled modules under the package name the reference uses for itself (`vsrlab`) in a
process that never imports the drop-in.

    python -m oracle.make_ref          # run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import hashlib
import json
import py_compile
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference/src")
DST = HERE / "_ref" / "src"

# exactly what `RealBasicVSR` imports (reference realbasicvsr.py:1-3, basicvsr.py:1-7, spynet.py:1-9, upsampling.py:1-2)
FILES = [
    "__init__.py",
    "core/__init__.py",
    "core/modules/__init__.py",
    "core/modules/conv.py",
    "core/modules/upsampling.py",
    "vsr/__init__.py",
    "vsr/models/__init__.py",
    "vsr/models/RealBasicVSR/__init__.py",
    "vsr/models/RealBasicVSR/realbasicvsr.py",
    "vsr/models/RealBasicVSR/modules/__init__.py",
    "vsr/models/RealBasicVSR/modules/basicvsr.py",
    "vsr/models/RealBasicVSR/modules/spynet.py",
]


# The callers of the hot path, for running the reference's own scripts unchanged on top of the drop-in package
# (tests/test_gpu_scripts.py, tools/run_reference_script.py): the entry scripts and the training runtime they import.
CALLER_FILES = [
    "train.py",
    "train_gan.py",
    "test.py",
    "core/utils.py",
    "core/losses.py",
    "core/metrics.py",
    "core/loggers.py",
    "vsr/dataset.py",
    "optical_flow/__init__.py",
    "optical_flow/models/__init__.py",
    "optical_flow/models/raft/__init__.py",
    "optical_flow/models/raft/raft.py",
    "optical_flow/models/raft/corr.py",
    "optical_flow/models/raft/extractor.py",
    "optical_flow/models/raft/update.py",
    "optical_flow/models/raft/utils.py",
]
REF_CONF = Path("/root/reference/conf")
CONF_DST = HERE / "_ref" / "conf"


def build_conf() -> None:
    """The reference's Hydra config tree as JSON (the `# @package` directive, a YAML comment, is kept as a key): parsed
    data, not a copy of the files.  The hydra shim of this repository reads .json configs like .yaml ones."""
    import re
    import yaml
    from omegaconf import _Loader            # YAML 1.2 floats (1e-4), like OmegaConf reads them
    for src in sorted(REF_CONF.rglob("*.yaml")):
        text = src.read_text()
        obj = yaml.load(text, Loader=_Loader) or {}
        m = re.search(r"^\s*#\s*@package\s+(\S+)", "\n".join(text.splitlines()[:5]), re.M)
        if m:
            obj = {"__package__": m.group(1), **obj}
        dst = (CONF_DST / src.relative_to(REF_CONF)).with_suffix(".json")
        dst.parent.mkdir(parents=True, exist_ok=True)
        dst.write_text(json.dumps(obj, indent=1))


def available() -> bool:
    m = DST.parent / "MANIFEST.json"
    if not (DST / "vsr/models/RealBasicVSR/realbasicvsr.bc").exists() or not m.exists():
        return False
    return json.loads(m.read_text()).get("python") == list(sys.version_info[:2])     # bytecode is interpreter-specific


def build(force: bool = False) -> bool:
    """Byte-compile the files; returns True when oracle/_ref is usable afterwards."""
    if not REF_SRC.exists():
        return available()
    if available() and not force:
        return True
    manifest = {}
    sys.path.insert(0, str(HERE.parent / "shims"))
    try:
        build_conf()
    finally:
        sys.path.pop(0)
    for rel in FILES + CALLER_FILES:
        src, dst = REF_SRC / rel, (DST / rel).with_suffix(".bc")
        dst.parent.mkdir(parents=True, exist_ok=True)
        py_compile.compile(str(src), cfile=str(dst), dfile=f"<reference>/src/{rel}", doraise=True)
        manifest[rel] = hashlib.sha256(src.read_bytes()).hexdigest()
    (DST.parent / "MANIFEST.json").write_text(json.dumps({"source": str(REF_SRC), "python": list(sys.version_info[:2]),
                                                           "sha256_of_sources": manifest}, indent=1))
    return True


if __name__ == "__main__":
    print("oracle/_ref:", "ready" if build(force=True) else "reference not present, nothing built")
