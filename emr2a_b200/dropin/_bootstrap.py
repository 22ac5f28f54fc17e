"""Make ``emr2a_b200`` importable when only this drop-in directory is on PYTHONPATH."""
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO not in sys.path:
    sys.path.append(_REPO)
