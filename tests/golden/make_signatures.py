"""Record the public call signatures of the reference's hot-path modules (names, parameter order, defaults) in
tests/golden/signatures.json.  Run in the build container, where /root/reference exists:

    python tests/golden/make_signatures.py

tests/test_host_cpu.py::test_public_signatures_match_the_reference compares emr2a_b200's same-named modules with it."""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402


def describe(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        default = None if p.default is inspect.Parameter.empty else repr(p.default)
        out.append([p.name, p.kind.name, default])
    return out


def public_api(module):
    api = {}
    for name, obj in sorted(vars(module).items()):
        if name.startswith("_") or getattr(obj, "__module__", None) != module.__name__:
            continue
        if inspect.isfunction(obj):
            api[name] = describe(obj)
        elif inspect.isclass(obj):
            methods = {m: describe(f) for m, f in sorted(vars(obj).items())
                       if inspect.isfunction(f) and (not m.startswith("__") or m == "__init__")}
            api[name] = {"methods": methods}
    return api


def main():
    retrieval, cvmod, metrics, common = load_reference()
    mods = {"retrieval.similarity": sys.modules["retrieval.similarity"], "retrieval.fusion": sys.modules["retrieval.fusion"],
            "retrieval.evaluator": sys.modules["retrieval.evaluator"], "utils.cv_evaluator": cvmod,
            "utils.metrics": metrics, "utils.common": common}
    out = {name: public_api(mod) for name, mod in mods.items()}
    out["retrieval.__all__"] = sorted(n for n in vars(retrieval) if not n.startswith("_") and callable(getattr(retrieval, n)))
    with open(os.path.join(HERE, "signatures.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print({k: (len(v) if isinstance(v, (dict, list)) else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
