"""Import alias.  The product package lives in the directory ``binaural-audio-synthesis_b200/``
(the project's name); a hyphen cannot appear in a Python identifier, so this one-file package
points its search path at that directory and runs its ``__init__``:

    import binaural_audio_synthesis_b200 as bas
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
_real = _os.path.join(_os.path.dirname(_here), 'binaural-audio-synthesis_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _os, _here, _real, _f
