"""Drop-in ``utils`` package: ``from utils.cv_evaluator import CVRetrievalEvaluator``
(analysis/run_cv_experiments.py:14) resolves to the B200 implementation.  Modules of the
reference's own ``utils`` package that are NOT on the retrieval hot path (e.g.
``utils.vlm_review``, imported at analysis/run_cv_experiments.py:15) stay importable:
any other ``utils/`` directory found later on sys.path is appended to this package's
search path."""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_sys.path.insert(0, _os.path.dirname(_here))
import _bootstrap  # noqa: E402,F401
_sys.path.pop(0)

for _entry in list(_sys.path):
    _cand = _os.path.join(_entry or ".", "utils")
    if _os.path.isdir(_cand) and _os.path.abspath(_cand) != _here and _os.path.abspath(_cand) not in map(_os.path.abspath, __path__):
        __path__.append(_cand)

from emr2a_b200.utils import *  # noqa: E402,F401,F403
from emr2a_b200.utils import __all__  # noqa: E402,F401
from emr2a_b200.utils import common, metrics, cv_evaluator  # noqa: E402,F401

_sys.modules[__name__ + ".common"] = common
_sys.modules[__name__ + ".metrics"] = metrics
_sys.modules[__name__ + ".cv_evaluator"] = cv_evaluator
