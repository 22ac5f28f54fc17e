"""Run a module of the reference project UNCHANGED on top of the B200 hot path.

    cd /path/to/emr2a-reference
    python /path/to/repo/emr2a_b200/run.py pipelines.step3_retrieval.run --manifest_path ... --embeddings_path ...
    python /path/to/repo/emr2a_b200/run.py analysis.run_cv_experiments --skip_encoding --embeddings_path e.npz ...

``python -m <module>`` from the reference root puts the current directory first on ``sys.path``, so the
reference's own ``retrieval`` / ``utils`` packages would win over PYTHONPATH.  This launcher inserts the
drop-in directory (same-named packages backed by libemr2a.so) ahead of everything, keeps the current
directory importable for the rest of the reference (``config``, ``data``, ``pipelines``, ``analysis``, and
``utils.vlm_review`` through the drop-in package's extended ``__path__``), and then executes the module as
``__main__`` with the remaining command line.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DROPIN = os.path.join(HERE, "dropin")


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        sys.stderr.write(__doc__)
        return 2
    module, rest = argv[0], argv[1:]
    cwd = os.getcwd()
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or cwd) not in (DROPIN, HERE)]
    if cwd not in sys.path and "" not in sys.path:
        sys.path.insert(0, cwd)
    sys.path.insert(0, DROPIN)
    for name in [m for m in sys.modules if m == "utils" or m.startswith("utils.") or m == "retrieval" or m.startswith("retrieval.")]:
        del sys.modules[name]
    sys.argv = [module] + rest
    runpy.run_module(module, run_name="__main__", alter_sys=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
