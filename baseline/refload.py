"""Imports the unmodified reference (nomadoor/ComfyUI-Video-Stabilizer) from baseline/_ref/ under ComfyUI stubs.

baseline/_ref/ is a byte-for-byte copy of /root/reference/{nodes,scripts,__init__.py,pyproject.toml,LICENSE} made by
`__graft_entry__.build()` in the build container (the reference is pure Python: "installing" it is copying it; it is
not a pip-installable distribution -- pyproject.toml has no build backend section for its flat layout).  The copy is
git-ignored and travels to the GPU box with gpurun, where /root/reference does not exist.  Nothing here is imported by
the product; bench.py's `--impl reference` arm and its `cpu_baseline` leg are the only users.

The stubs follow the reference's own scripts (scripts/check_crop_aspect_ratio.py:30-55,
scripts/compare_refactor_behavior.py:75-109): `comfy_api.latest` with `io.*` socket kinds and a `NodeOutput` that keeps
its arguments, `comfy.utils.ProgressBar`, and a `comfy` package without `model_management` (the nodes treat the
interrupt hook as optional).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SOURCE_DIR = "/root/reference"
COPIED = ("nodes", "scripts", "__init__.py", "pyproject.toml", "LICENSE")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "nodes", "video_stabilizer_flow.py"))


def sync_from_source() -> bool:
    """Build container only: refresh baseline/_ref/ from /root/reference.  Returns True when the copy exists."""
    import shutil

    if not os.path.isdir(os.path.join(SOURCE_DIR, "nodes")):
        return available()
    os.makedirs(REF_DIR, exist_ok=True)
    for name in COPIED:
        src, dst = os.path.join(SOURCE_DIR, name), os.path.join(REF_DIR, name)
        if os.path.isdir(src):
            shutil.copytree(src, dst, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.isfile(src):
            shutil.copyfile(src, dst)
    return available()


class NodeOutput(tuple):
    """io.NodeOutput stand-in that keeps what the node returned: (frames, mask, meta)."""

    def __new__(cls, *args, **kwargs):
        return super().__new__(cls, args)


def install_stubs() -> None:
    if getattr(sys.modules.get("comfy_api.latest"), "_vstab_bench_stub", False):
        return

    class _Socket:
        def __init__(self, *a, **k):
            self.args, self.kwargs = a, k

    class _Kind:
        Input = Output = _Socket

    class _ControlAfterGenerate:
        fixed = increment = decrement = randomize = "fixed"

    io = types.SimpleNamespace(
        ComfyNode=type("ComfyNode", (), {}), Custom=lambda name: _Kind, Image=_Kind, Mask=_Kind, Float=_Kind, Combo=_Kind,
        Boolean=_Kind, Color=_Kind, Int=_Kind, String=_Kind, NumberDisplay=types.SimpleNamespace(slider="slider", number="number"),
        ControlAfterGenerate=_ControlAfterGenerate,
        Schema=type("Schema", (), {"__init__": lambda self, **k: self.__dict__.update(k)}),
        NodeOutput=NodeOutput,
    )
    latest = types.ModuleType("comfy_api.latest")
    latest.ComfyExtension = type("ComfyExtension", (), {})
    latest.io = io
    latest._vstab_bench_stub = True
    sys.modules.setdefault("comfy_api", types.ModuleType("comfy_api"))
    sys.modules["comfy_api.latest"] = latest
    comfy = types.ModuleType("comfy")
    comfy.__path__ = []  # a package without comfy.model_management: the interrupt hook stays off
    utils = types.ModuleType("comfy.utils")
    utils.ProgressBar = type("ProgressBar", (), {"__init__": lambda self, total: None, "update": lambda self, amount: None,
                                                 "update_absolute": lambda self, value, total=None, preview=None: None})
    sys.modules["comfy"] = comfy
    sys.modules["comfy.utils"] = utils


def load(root: str | None = None):
    """-> namespace of the reference's node modules, imported from `root` (default baseline/_ref)."""
    root = root or REF_DIR
    if not os.path.isfile(os.path.join(root, "nodes", "video_stabilizer_flow.py")):
        raise FileNotFoundError(f"no reference copy under {root}: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "in the build container (it copies /root/reference there)")
    sys.dont_write_bytecode = True
    install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    names = ["stabilizer_utils", "motion_meta", "motion_apply", "shake_noise", "video_stabilizer_flow", "video_stabilizer_classic",
             "video_stabilizer_motion_apply"]
    return types.SimpleNamespace(**{n: importlib.import_module(f"nodes.{n}") for n in names}, root=root)
