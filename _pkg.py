"""Import helper: the package directory is `ros-turtlebot-navigation_b200/` (hyphens), which the
import statement cannot spell.  load() registers it as `ros_turtlebot_navigation_b200`."""
import importlib.util
import os
import sys

NAME = "ros_turtlebot_navigation_b200"
ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "ros-turtlebot-navigation_b200")


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod


def load_build():
    """The build script alone (does not need libb2nav.so to exist)."""
    spec = importlib.util.spec_from_file_location(NAME + "_build", os.path.join(PKG_DIR, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
