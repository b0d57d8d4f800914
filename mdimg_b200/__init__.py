"""Import shim: the product package lives in ``medical-image-enhancer_b200/`` (a directory name
that is not a valid Python identifier); this module makes it importable as ``mdimg_b200``."""

import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "medical-image-enhancer_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _os, _f
