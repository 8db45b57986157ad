"""Import alias: the package directory is named `roboken-fmskf-robot-controller_b200/` (a
hyphen is not importable), so this module loads it under an importable name:

    import roboken_fmskf_robot_controller_b200 as rk
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "roboken-fmskf-robot-controller_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
