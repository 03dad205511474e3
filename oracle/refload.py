"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (JeremyLinky/YouTube-VLN).

Used only by ``oracle/make_golden.py`` (run in the build container, where ``/root/reference`` is
mounted read-only) to import ``vilbert.vilbert`` / ``lily`` / ``utils.utils_init`` exactly as they are
and record golden vectors.  Nothing under ``tests/``, ``bench.py`` or the product package may import this
module at run time on the GPU box: ``/root/reference`` does not exist there.

The reference needs a handful of third-party modules that are absent here:
``boto3``/``botocore`` (vilbert/file_utils.py:20-21), ``lmdb``, ``tensorboardX``, ``colorama``,
``termcolor``, ``pyfiglet``, ``argtyped`` (import chain of utils/utils_init.py:5-6).  They are replaced by
empty stub modules in ``sys.modules``; no reference file is modified or copied.
"""
import importlib
import sys
import types

import torch  # noqa: F401  (imported before the stubs go in)

REFERENCE_ROOT = "/root/reference"

_STUBS = {
    "boto3": {},
    "botocore": {},
    "botocore.exceptions": {"ClientError": type("ClientError", (Exception,), {})},
    "lmdb": {},
    "tensorboardX": {"SummaryWriter": object},
    "colorama": {"Fore": types.SimpleNamespace(), "Style": types.SimpleNamespace(), "init": lambda *a, **k: None},
    "termcolor": {"colored": lambda s, *a, **k: s, "cprint": lambda *a, **k: None},
    "pyfiglet": {"Figlet": object},
    "argtyped": {"Arguments": type("Arguments", (), {"__init_subclass__": classmethod(lambda cls, **kw: None)}),
                 "Switch": bool},
}


def _install_stubs():
    for name, attrs in _STUBS.items():
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        # any other attribute resolves to an inert callable/class so ``from x import y`` succeeds
        def _inert(attr, _n=name):
            if attr.startswith("__"):
                raise AttributeError(attr)
            return type(attr, (), {"__init__": lambda self, *a, **k: None,
                                   "__call__": lambda self, *a, **k: None})
        mod.__getattr__ = _inert
        mod.__yvb200_stub__ = True
        sys.modules[name] = mod


def load_reference(root: str = REFERENCE_ROOT):
    """Import the reference packages from ``root`` and return (vilbert_module, lily_module, utils_init)."""
    _install_stubs()
    # the reference must win over the drop-in package of the same name
    for k in [k for k in sys.modules if k == "vilbert" or k.startswith("vilbert.") or k == "lily"
              or k == "utils" or k.startswith("utils.")]:
        del sys.modules[k]
    sys.path.insert(0, root)
    try:
        vb = importlib.import_module("vilbert.vilbert")
        lily = importlib.import_module("lily")
        try:
            ui = importlib.import_module("utils.utils_init")
        except Exception as e:  # losses are restated in the oracle if the loop module cannot import
            ui = None
            print("warning: utils.utils_init not importable:", repr(e))
    finally:
        sys.path.remove(root)
    assert vb.__file__.startswith(root), vb.__file__
    return vb, lily, ui
