"""Imports the UNMODIFIED reference from /root/reference under ComfyUI stubs.

Only available in the build container; used by ``-m "not gpu"`` tests marked ``reference`` and by
scripts/make_golden.py to generate the committed fixtures.  Stub shape follows the recipe in the
reference's own scripts/check_crop_aspect_ratio.py:30-55.
"""
from __future__ import annotations

import importlib
import sys
import types

REFERENCE_DIR = "/root/reference"


def _install_stubs() -> None:
    if "comfy_api.latest" in sys.modules:
        return

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

    class _IO(types.SimpleNamespace):
        def __getattr__(self, name):
            return _Anything()

    comfy_api = types.ModuleType("comfy_api")
    latest = types.ModuleType("comfy_api.latest")
    latest.ComfyExtension = type("ComfyExtension", (), {})
    latest.io = _IO(ComfyNode=type("ComfyNode", (), {}), Custom=lambda name: name)
    sys.modules["comfy_api"] = comfy_api
    sys.modules["comfy_api.latest"] = latest
    comfy = types.ModuleType("comfy")
    comfy.__path__ = []
    utils = types.ModuleType("comfy.utils")
    utils.ProgressBar = type(
        "ProgressBar",
        (),
        {
            "__init__": lambda self, total: None,
            "update": lambda self, amount: None,
            "update_absolute": lambda self, value, total=None, preview=None: None,
        },
    )
    sys.modules["comfy"] = comfy
    sys.modules["comfy.utils"] = utils


def load_reference():
    sys.dont_write_bytecode = True
    _install_stubs()
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    names = ["stabilizer_utils", "motion_meta", "motion_apply", "shake_noise", "video_stabilizer_flow", "video_stabilizer_classic"]
    return types.SimpleNamespace(**{n: importlib.import_module(f"nodes.{n}") for n in names})
