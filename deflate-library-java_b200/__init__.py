"""b2deflate -- B200-native DEFLATE codec behind the io.nayuki.deflate stream API.

The directory name carries a hyphen (it mirrors the reference repository's name), so it is loaded by path:
    import importlib.util, sys
    spec = importlib.util.spec_from_file_location("b2deflate", "<repo>/deflate-library-java_b200/__init__.py",
                                                  submodule_search_locations=["<repo>/deflate-library-java_b200"])
    b2deflate = importlib.util.module_from_spec(spec); sys.modules["b2deflate"] = b2deflate
    spec.loader.exec_module(b2deflate)
(`b2d_loader.load()` at the repository root does exactly this.)

Contents: csrc/ (CUDA kernels + C-ABI host runtime -> libb2deflate.so), binding.py (ctypes over the C ABI),
sharding.py (unit ranges and the gather to rank 0 for torch.distributed runs), host/ (C++ mirror of InflaterInputStream /
DeflaterOutputStream / Gzip*Stream / Zlib*Stream and of the gzip / gunzip CLIs -> bin/), java/ (Panama FFM sources of the
same classes; uncompiled here -- no JDK in the image).
"""
from . import binding  # noqa: F401
from .binding import *  # noqa: F401,F403
