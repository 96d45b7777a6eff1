"""scripts/train_dp_harness.py as an importable module (scripts/ is not a package)."""
import importlib.util
import os

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("train_dp_harness", os.path.join(_ROOT, "scripts", "train_dp_harness.py"))
harness = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(harness)
