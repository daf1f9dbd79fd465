"""Import alias: `import heist_b200` loads the package that lives in the (non-identifier) directory
rl-project-heist-architect-adversarial-reinforcement-learning-framework-cse4019_b200/."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_ROOT, "rl-project-heist-architect-adversarial-reinforcement-learning-framework-cse4019_b200")
_spec = importlib.util.spec_from_file_location(
    "heist_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["heist_b200"] = _mod
_spec.loader.exec_module(_mod)
