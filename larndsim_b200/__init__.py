"""Import alias: the package directory is ``larnd-sim_b200/`` (not a valid Python identifier);
``import larndsim_b200`` loads it from there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "larnd-sim_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _os, _f, _real
