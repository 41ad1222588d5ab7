"""`ovdet` - importable name of the B200-native open-vocabulary head + post-processing.

The implementation lives in the directory
``real-time-zero-shot-open-vocabulary-object-detection-using-a-lightweight_b200/`` (a name
Python cannot import because of the hyphens); this package only extends its ``__path__`` to
that directory so that ``ovdet.ops``, ``ovdet.heads``, ``ovdet.detector`` ... resolve to
the files there.  There is no second copy of any module.
"""
import os as _os

PKG_DIR = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "real-time-zero-shot-open-vocabulary-object-detection-using-a-lightweight_b200",
)
__path__.append(PKG_DIR)

__version__ = "0.1.0"
