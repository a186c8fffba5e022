"""Import alias for the package directory ``gradient-accumulation-tf-estimator_b200/``.

The package is named after the reference repository, and a hyphen cannot appear in a Python
``import`` statement; ``import gaccum_b200`` (and ``gaccum_b200.optimization`` ...) resolves to
the files in that directory.  Nothing lives here.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "gradient-accumulation-tf-estimator_b200")
__path__ = [_REAL]
__file__ = _os.path.join(_REAL, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f, _os
