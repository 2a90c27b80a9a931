"""Import alias: the package sources live in ``open3d-ml_b200/`` (a directory
name that Python cannot import directly); this stub makes them importable as
``open3d_ml_b200``."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "open3d-ml_b200"))
from ._pkg import *  # noqa: F401,F403,E402
from ._pkg import __version__, __getattr__  # noqa: F401,E402
