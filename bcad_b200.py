"""Importable name for the package directory ``vision-xai-breast-cancer-cad_b200/`` (hyphens are not
legal in a Python module name): ``import bcad_b200`` / ``from bcad_b200 import CNNModel``."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vision-xai-breast-cancer-cad_b200")
_spec = importlib.util.spec_from_file_location(
    "bcad_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["bcad_b200"] = _mod
_spec.loader.exec_module(_mod)
