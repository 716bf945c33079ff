"""Import shim: the product package lives in ``minimax-speech_b200/`` (a directory name that is
not a valid Python identifier).  This module makes it importable as ``minimax_speech_b200``."""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
__path__ = [_os.path.join(_os.path.dirname(_here), "minimax-speech_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
