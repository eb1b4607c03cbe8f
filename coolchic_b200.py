"""Import alias: ``import coolchic_b200`` loads the package stored in ``cool-chic_b200/``
(a hyphen cannot appear in a Python module name).  Sub-modules are then importable as
``coolchic_b200.bitstream.decode`` etc."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "cool-chic_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
