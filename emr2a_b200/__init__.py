"""emr2a_b200 -- B200-native (sm_100a) retrieval hot path of EMR2A.

Sub-packages ``retrieval`` and ``utils`` mirror the reference's Python modules; the
compute is in ``libemr2a.so`` (hand-written CUDA, C-ABI in ``include/emr2a.h``).
"""
__version__ = "0.1.0"
