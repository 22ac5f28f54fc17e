"""Golden outputs of the reference's own ENTRY POINTS (build container only -- needs /root/reference).

    python tests/golden/make_script_goldens.py

Runs, unchanged and in-process,
    python -m pipelines.step3_retrieval.run   --manifest_path .. --embeddings_path .. --output_dir ..
    python -m analysis.run_cv_experiments     --skip_encoding --embeddings_path .. --experiment_id .. --fusion concat|late ..
on a small synthetic cohort (manifest.jsonl + .npz files in the layouts those scripts read:
pipelines/step3_retrieval/evaluate_retrieval.py:30-33, analysis/run_cv_experiments.py:111-128), with RECORDING
wrappers around the evaluator classes the scripts construct (pipelines/step3_retrieval/evaluate_retrieval.py:64-78,
analysis/run_cv_experiments.py:383-397, 490-495): every constructor / call argument is captured, then the reference
implementation runs and writes its files.  Committed under tests/golden/scripts/:

    inputs.npz                      the synthetic cohort (to rebuild the input files anywhere)
    calls.json + calls.npz          the recorded constructor / call arguments (arrays in the npz)
    step3/retrieval_results.json    what the reference wrote
    cv_concat/, cv_late/            exp_<id>/config.json, summary.csv, fold_k/metrics.json as the reference wrote them

tests/test_gpu_scripts_replay.py replays exactly those calls on the drop-in classes (on the GPU box, where the
reference tree does not exist) and compares the files key for key and value for value under the gap rule.
The unseeded PCA inside run_cv (utils/cv_evaluator.py:89) is pinned by ``np.random.seed(SEED)`` immediately before the
recorded call; the replay does the same.
"""
import importlib.machinery as im
import json
import os
import runpy
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("EMR2A_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "scripts")
SEED = 4242
N, SLICES, D_IMG, D_TXT, N_CLS = 150, 3, 48, 40, 3


def _stub(name, **kw):
    m = types.ModuleType(name)
    m.__spec__ = im.ModuleSpec(name, None)
    m.__path__ = []
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


def stubs():
    """Modules the reference imports at module level but never touches on these code paths."""
    noop = lambda *a, **k: None            # noqa: E731

    class _Axes:                           # the PNG plot (utils/cv_evaluator.py:459-499) is the only matplotlib user
        def __getattr__(self, name):
            return noop
    plt = _stub("matplotlib.pyplot", subplots=lambda *a, **k: (None, [_Axes(), _Axes()]), tight_layout=noop, savefig=noop,
                close=noop)
    _stub("matplotlib", pyplot=plt)
    _stub("seaborn", heatmap=noop)
    _stub("qwen_vl_utils", process_vision_info=noop)
    _stub("timm.data", create_transform=noop, resolve_data_config=noop)
    _stub("timm", data=sys.modules["timm.data"], create_model=noop)
    _stub("open_clip", create_model_from_pretrained=noop, get_tokenizer=noop)


def cohort():
    sys.path.insert(0, REPO)
    from emr2a_b200 import synth
    sys.path.remove(REPO)
    rng = np.random.default_rng(2024)
    base = synth.two_modal(N, D_IMG, D_TXT, N_CLS, seed=61, sep=0.35)
    slices = base["image"][:, None, :] + 0.3 * rng.standard_normal((N, SLICES, D_IMG)).astype(np.float32)
    ids = [f"case_{j:04d}" for j in range(N)]
    return {"ids": np.array(ids), "image_slices": slices.astype(np.float32), "text": base["text"].astype(np.float32),
            "labels": base["labels"].astype(np.int32)}


def write_inputs(c, root):
    """The three files the scripts read (same function is used by the replay test to sanity-check the recorded arrays)."""
    manifest = os.path.join(root, "manifest.jsonl")
    with open(manifest, "w", encoding="utf-8") as fh:
        for pid, lab in zip(c["ids"], c["labels"]):
            fh.write(json.dumps({"patient_id": str(pid), "label": f"class_{int(lab)}", "slices": [], "meta": {}}) + "\n")
    np.savez(os.path.join(root, "step3_embeddings.npz"), **{str(pid): c["image_slices"][j] for j, pid in enumerate(c["ids"])})
    np.savez(os.path.join(root, "cv_embeddings.npz"), patient_ids=np.array([str(p) for p in c["ids"]], dtype=object),
             image_matrix=c["image_slices"], text_matrix=c["text"])
    return manifest


class Recorder:
    def __init__(self):
        self.calls = {}
        self.arrays = {}

    def put_array(self, name, arr):
        self.arrays[name] = np.asarray(arr)
        return {"npz": name}


def main():
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    stubs()
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    import retrieval
    import utils.cv_evaluator as cvmod
    rec = Recorder()
    c = cohort()
    np.savez_compressed(os.path.join(OUT, "inputs.npz"), **c)
    work = tempfile.mkdtemp(prefix="emr2a_script_goldens_")
    manifest = write_inputs(c, work)
    cwd = os.getcwd()
    os.chdir(work)                       # BaseConfig paths are relative; nothing is written outside the temp dir

    # ---- recording wrappers (call through to the reference implementation) ----
    state = {"tag": None}
    ho_eval = retrieval.RetrievalEvaluator.evaluate_retrieval
    ho_init = retrieval.RetrievalEvaluator.__init__

    def rec_ho_init(self, *a, **kw):
        rec.calls["step3"] = {"ctor_args": list(a), "ctor_kwargs": kw}
        return ho_init(self, *a, **kw)

    def rec_ho_eval(self, *a, **kw):
        assert not a, "the script calls evaluate_retrieval with keyword arguments only"
        call = {}
        for k, v in kw.items():
            if isinstance(v, np.ndarray):
                call[k] = rec.put_array(f"step3_{k}", v)
            else:
                call[k] = v
        rec.calls["step3"]["evaluate_retrieval"] = call
        return ho_eval(self, *a, **kw)

    retrieval.RetrievalEvaluator.__init__ = rec_ho_init
    retrieval.RetrievalEvaluator.evaluate_retrieval = rec_ho_eval
    cv_init, cv_run, cv_save = (cvmod.CVRetrievalEvaluator.__init__, cvmod.CVRetrievalEvaluator.run_cv,
                                cvmod.CVRetrievalEvaluator.save_results)

    def rec_cv_init(self, *a, **kw):
        rec.calls[state["tag"]] = {"ctor_args": list(a), "ctor_kwargs": kw}
        return cv_init(self, *a, **kw)

    def rec_cv_run(self, *a, **kw):
        assert not a, "the script calls run_cv with keyword arguments only"
        tag = state["tag"]
        ids = list(kw["patient_ids"])
        emb = kw["embeddings"]
        call = {"patient_ids": rec.put_array(f"{tag}_patient_ids", np.array(ids)),
                "labels": list(kw["labels"]),
                "embeddings": {"image": rec.put_array(f"{tag}_image", np.stack([emb[p]["image"] for p in ids])),
                               "text": rec.put_array(f"{tag}_text", np.stack([emb[p]["text"] for p in ids]))},
                "fusion": kw["fusion"], "top_k_list": list(kw["top_k_list"]), "w_text": kw["w_text"],
                "numpy_seed_before_call": SEED}
        rec.calls[tag]["run_cv"] = call
        np.random.seed(SEED)
        return cv_run(self, *a, **kw)

    def rec_cv_save(self, *a, **kw):
        assert not a
        rec.calls[state["tag"]]["save_results"] = {"experiment_id": kw["experiment_id"], "config": kw["config"]}
        return cv_save(self, *a, **kw)

    cvmod.CVRetrievalEvaluator.__init__ = rec_cv_init
    cvmod.CVRetrievalEvaluator.run_cv = rec_cv_run
    cvmod.CVRetrievalEvaluator.save_results = rec_cv_save

    def run_module(module, argv):
        old = sys.argv
        sys.argv = [module] + argv
        try:
            runpy.run_module(module, run_name="__main__", alter_sys=True)
        finally:
            sys.argv = old

    # ---- step 3 ----
    step3_args = ["--manifest_path", manifest, "--embeddings_path", os.path.join(work, "step3_embeddings.npz"),
                  "--output_dir", os.path.join(work, "step3_out"), "--top_k", "5"]
    run_module("pipelines.step3_retrieval.run", step3_args)
    os.makedirs(os.path.join(OUT, "step3"))
    shutil.copy(os.path.join(work, "step3_out", "retrieval_results.json"), os.path.join(OUT, "step3"))
    rec.calls["step3"]["argv"] = ["--top_k", "5"]

    # ---- run_cv_experiments: concat and late ----
    for tag, extra in (("cv_concat", ["--fusion", "concat", "--pca_dim", "16", "--top_k", "3"]),
                       ("cv_late", ["--fusion", "late", "--w_text", "0.25", "--pca_dim", "24", "--top_k", "5"])):
        state["tag"] = tag
        out_dir = os.path.join(work, tag)
        run_module("analysis.run_cv_experiments",
                   ["--manifest_path", manifest, "--skip_encoding", "--embeddings_path", os.path.join(work, "cv_embeddings.npz"),
                    "--output_dir", out_dir, "--experiment_id", tag, "--device", "cpu"] + extra)
        rec.calls[tag]["argv"] = extra
        src = os.path.join(out_dir, f"exp_{tag}")
        dst = os.path.join(OUT, tag)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("*.png"))
    os.chdir(cwd)
    with open(os.path.join(OUT, "calls.json"), "w", encoding="utf-8") as fh:
        json.dump(rec.calls, fh, indent=1, ensure_ascii=False, default=lambda o: o.item() if hasattr(o, "item") else str(o))
    np.savez_compressed(os.path.join(OUT, "calls.npz"), **rec.arrays)
    shutil.rmtree(work)
    total = 0
    for root, _, files in os.walk(OUT):
        for f in files:
            total += os.path.getsize(os.path.join(root, f))
    print(f"wrote {OUT}: {total / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
