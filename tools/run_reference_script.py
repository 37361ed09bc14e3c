"""Run one of the reference's OWN entry scripts (train.py / train_gan.py / test.py), unchanged, on top of the drop-in `vsrlab` package.

    python tools/run_reference_script.py train --project /tmp/proj -- +experiment=basic train.model.pretrained_flow=false ...

What it arranges (nothing in the reference is edited):
  * the script and the caller modules it imports come from the reference tree: `/root/reference/src` where that exists
    (the build container), else the byte-compiled copies `oracle/make_ref.py` placed under `oracle/_ref/src`;
    the drop-in package loads them underneath itself through `VSRLAB_REFERENCE_SRC` (vsrlab/_overlay.py);
  * a project directory `<project>/{conf,src,storage}`: `conf/` = the reference's Hydra tree plus the data group it lacks
    (`conf/train/data/*.yaml` of this repository); the script runs with cwd `<project>/src`, because the reference derives
    PROJECT_ROOT from `Path.cwd().parents[0]` (src/core/__init__.py:8);
  * `hydra`, `omegaconf`, `kornia`, `piqa` resolve to the minimal shims under `shims/` (not installable here: no network);
    W&B runs disabled; the single-process rendezvous variables `torchrun` would set are provided.
"""
from __future__ import annotations

import argparse
import os
import shutil
import socket
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF_SRC, REF_CONF = Path("/root/reference/src"), Path("/root/reference/conf")
BC_SRC, BC_CONF = ROOT / "oracle" / "_ref" / "src", ROOT / "oracle" / "_ref" / "conf"

_RUN_BC = ("import sys, importlib.machinery as m; p = sys.argv[1]; sys.argv = [p] + sys.argv[2:]; "
           "code = m.SourcelessFileLoader('__main__', p).get_code('__main__'); "
           "exec(code, {'__name__': '__main__', '__file__': p, '__builtins__': __builtins__})")


def reference_tree():
    if (REF_SRC / "train.py").exists():
        return REF_SRC, REF_CONF, ".py"
    if (BC_SRC / "train.bc").exists():
        return BC_SRC, BC_CONF, ".bc"
    raise FileNotFoundError("neither /root/reference nor oracle/_ref (python -m oracle.make_ref) is available")


def make_project(project: Path) -> Path:
    _, conf, _ = reference_tree()
    (project / "src").mkdir(parents=True, exist_ok=True)
    (project / "storage").mkdir(exist_ok=True)
    if (project / "conf").exists():
        shutil.rmtree(project / "conf")
    shutil.copytree(conf, project / "conf")
    shutil.copytree(ROOT / "conf" / "train" / "data", project / "conf" / "train" / "data", dirs_exist_ok=True)
    return project


def run(script: str, overrides, project: Path, extra_env=None, timeout=None) -> subprocess.CompletedProcess:
    src, _, suffix = reference_tree()
    make_project(project)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(ROOT), str(ROOT / "shims")] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    env["VSRLAB_REFERENCE_SRC"] = str(src)
    env.setdefault("WANDB_MODE", "disabled")
    env["LOGGING_DIR"] = str(project / "storage")      # (an importing parent process may have exported its own: vsrlab.core sets it)
    env.pop("PROJECT_ROOT", None)
    if "MASTER_PORT" not in env:
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        env["MASTER_PORT"] = str(s.getsockname()[1])
        s.close()
    env.setdefault("MASTER_ADDR", "127.0.0.1")
    for k, v in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")):
        env.setdefault(k, v)
    env.update(extra_env or {})
    path = str(src / (script + suffix))
    cmd = [sys.executable, path, *overrides] if suffix == ".py" else [sys.executable, "-c", _RUN_BC, path, *overrides]
    return subprocess.run(cmd, cwd=str(project / "src"), env=env, capture_output=True, text=True, timeout=timeout)


def main():
    argv = sys.argv[1:]
    overrides = []
    if "--" in argv:                                    # everything after `--` goes to the script (Hydra overrides start with + / ~)
        i = argv.index("--")
        argv, overrides = argv[:i], argv[i + 1:]
    ap = argparse.ArgumentParser()
    ap.add_argument("script", choices=["train", "train_gan", "test"])
    ap.add_argument("--project", default="")
    a = ap.parse_args(argv)
    a.overrides = overrides
    project = Path(a.project) if a.project else Path(tempfile.mkdtemp(prefix="vsrlab_proj_"))
    r = run(a.script, a.overrides, project)
    sys.stdout.write(r.stdout)
    sys.stderr.write(r.stderr)
    print(f"[project dir: {project}]", file=sys.stderr)
    sys.exit(r.returncode)


if __name__ == "__main__":
    main()
