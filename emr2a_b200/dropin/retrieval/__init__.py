"""Drop-in ``retrieval`` package: put ``<repo>/emr2a_b200/dropin`` ahead of the reference
checkout on PYTHONPATH and ``from retrieval import RetrievalEvaluator``
(pipelines/step3_retrieval/evaluate_retrieval.py:12) resolves to the B200 implementation."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import _bootstrap  # noqa: E402,F401
_sys.path.pop(0)

from emr2a_b200.retrieval import *  # noqa: E402,F401,F403
from emr2a_b200.retrieval import __all__  # noqa: E402,F401
from emr2a_b200.retrieval import similarity, fusion, evaluator  # noqa: E402,F401

_sys.modules[__name__ + ".similarity"] = similarity
_sys.modules[__name__ + ".fusion"] = fusion
_sys.modules[__name__ + ".evaluator"] = evaluator
