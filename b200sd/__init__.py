"""Import alias: the package sources live in ``ml-stable-diffusion_b200/`` (a directory name
Python cannot import directly).  ``import b200sd`` executes that package's ``__init__`` with
``__path__`` pointing there, so ``b200sd.unet`` is ``ml-stable-diffusion_b200/unet.py``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "ml-stable-diffusion_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
