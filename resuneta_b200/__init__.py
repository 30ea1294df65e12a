"""Import shim: the product package lives in ``resunet-a_mltsk_keras_b200/`` (a directory name
Python cannot import directly); this makes it importable as ``resuneta_b200``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "resunet-a_mltsk_keras_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
