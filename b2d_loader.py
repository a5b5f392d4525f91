"""Loads the product package `deflate-library-java_b200/` (hyphenated directory) as module `b2deflate`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "deflate-library-java_b200")


def load():
    if "b2deflate" in sys.modules:
        return sys.modules["b2deflate"]
    spec = importlib.util.spec_from_file_location("b2deflate", os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["b2deflate"] = mod
    spec.loader.exec_module(mod)
    return mod
