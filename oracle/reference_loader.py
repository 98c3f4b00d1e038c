"""Imports the UNMODIFIED reference (frgfm/Holocron) from /root/reference as ``holocron.*`` modules.

Only usable in the build container (the reference tree does not exist on the GPU box). Used by
``tests/golden/make_golden.py`` to generate the committed golden fixtures, and by the optional
``tests/test_oracle_vs_reference.py`` cross-check, which skips itself when the tree is absent.

``import holocron`` itself fails in this image (its __init__ pulls matplotlib/fastprogress and a generated
version.py), so a stub parent package whose __path__ points at the reference is pre-seeded and the needed
sub-packages are imported directly.
"""
import importlib
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")


def available() -> bool:
    return (REFERENCE_ROOT / "holocron" / "nn" / "functional.py").exists()


def load():
    """Returns the stub ``holocron`` package with nn, ops, optim and models imported from the reference."""
    if not available():
        raise RuntimeError("reference tree not available at /root/reference")
    existing = sys.modules.get("holocron")
    if existing is not None and getattr(existing, "__holocron_reference__", False):
        return existing
    pkg = types.ModuleType("holocron")
    pkg.__path__ = [str(REFERENCE_ROOT / "holocron")]
    pkg.__holocron_reference__ = True
    sys.modules["holocron"] = pkg
    for sub in ("nn", "ops", "optim", "models"):
        setattr(pkg, sub, importlib.import_module(f"holocron.{sub}"))
    return pkg
