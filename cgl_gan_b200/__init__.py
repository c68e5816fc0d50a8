"""Import shim: the product package lives in the directory `cgl-gan_b200/` (the name the build
contract fixes); a hyphen is not importable, so `import cgl_gan_b200` resolves here and forwards
its package path there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "cgl-gan_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
