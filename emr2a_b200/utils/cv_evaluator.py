"""``utils.cv_evaluator.CVRetrievalEvaluator`` on the B200.

Same constructor, methods, argument order, defaults, result keys, file formats and
errors as the reference class (utils/cv_evaluator.py:26-501).  Inside a fold, everything
after ``process_embeddings`` -- fusion, similarity, Top-K, votes, hit flags and confusion
counts (reference lines 186-310, a per-query python loop) -- is one K1 -> K2 -> K4 pass
on the GPU.  StratifiedKFold stays sklearn-on-host (identical fold membership).  The per-fold
StandardScaler + PCA run either on the host with sklearn, exactly as the reference does, or on the
GPU (``emr2a_b200.preprocess``: hand-written moment / standardise kernels + an exact float64
eigen-decomposition) -- see ``CVRetrievalEvaluator.preprocess``.
"""
import csv
import json
import logging
import os
import random
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np
from sklearn.decomposition import PCA
from sklearn.model_selection import StratifiedKFold
from sklearn.preprocessing import StandardScaler

from .. import native
from ..engine import get_engine
from ..labels import encode, gather_lists, score_lists
from .metrics import confusion_dict, prf_from_confusion

logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s")
logger = logging.getLogger(__name__)

_SUMMARY_METRICS = ("top1", "top3", "top5", "vote_acc", "weighted_vote_acc",
                    "macro_precision", "macro_recall", "macro_f1")


class CVRetrievalEvaluator:
    #: where StandardScaler + PCA of ``process_embeddings`` run (env ``EMR2A_PREPROCESS``, default "auto"):
    #:   "host"  sklearn on the CPU, as the reference (bit-for-bit its preprocessing, given the same numpy seed);
    #:   "gpu"   emr2a_b200.preprocess: scaler identical to sklearn's, PCA = the deterministic exact basis
    #:           (float64 covariance + eigen-decomposition), all on the device.  sklearn's own fp32 solvers sit
    #:           ~2e-6 ("covariance_eigh") to ~2e-4 ("full") from that basis on the golden folds, its unseeded
    #:           "randomized" solver further and differently on every run;
    #:   "auto"  the device wherever that keeps the parity contract or the contract is void, sklearn otherwise:
    #:             no PCA (scaler only) ............ device (bit-identical to sklearn)
    #:             sklearn would use covariance_eigh  device (rows within ~2e-6 of sklearn's, scores within 1e-5)
    #:             sklearn would use randomized ...... device (the reference does not seed it, utils/cv_evaluator.py:89:
    #:                                                its own output changes from run to run; the exact basis is what
    #:                                                the random one approximates -- set "host" to reproduce a run
    #:                                                whose numpy RNG you seeded yourself)
    #:             sklearn would use full ............ host  (deterministic in sklearn, and its fp32 result is further
    #:                                                from the exact basis than the 1e-5 score tolerance)
    preprocess: str = os.environ.get("EMR2A_PREPROCESS", "auto")

    def __init__(self, cv_folds: int = 5, pca_dim: int = 128, top_k: int = 5, seed: int = 42):
        self.cv_folds = cv_folds
        self.pca_dim = pca_dim
        self.top_k = top_k
        self.seed = seed
        self.rng = np.random.RandomState(seed)
        self.random = random.Random(seed)

    def _preprocess_on_gpu(self, n_train: int, n_features: int) -> bool:
        mode = self.preprocess
        if mode not in ("host", "gpu", "auto"):
            raise ValueError(f"unknown preprocess mode {mode!r} (host, gpu, auto)")
        if mode != "auto":
            return mode == "gpu"
        from ..preprocess import sklearn_solver
        n_comp = min(self.pca_dim, n_train - 1, n_features)
        return n_comp <= 0 or sklearn_solver(n_train, n_features, n_comp) != "full"

    # ------------------------------------------------------------------ host
    def stratified_split(self, patient_ids: List[str], labels: List[str]) -> List[Tuple[List[str], List[str]]]:
        """StratifiedKFold(cv_folds, shuffle=True, random_state=seed) -> (train_ids, test_ids)
        per fold (utils/cv_evaluator.py:41-54)."""
        folds = StratifiedKFold(n_splits=self.cv_folds, shuffle=True, random_state=self.seed)
        ids = np.asarray(patient_ids, dtype=object)
        return [(ids[tr].tolist(), ids[te].tolist()) for tr, te in folds.split(patient_ids, labels)]

    def _make_serializable(self, obj):
        """numpy -> python natives for json (utils/cv_evaluator.py:56-71)."""
        if isinstance(obj, dict):
            return {k: self._make_serializable(v) for k, v in obj.items()}
        if isinstance(obj, list):
            return [self._make_serializable(v) for v in obj]
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.bool_):
            return bool(obj)
        return obj

    def process_embeddings(self, train_embeddings: np.ndarray, test_embeddings: np.ndarray
                           ) -> Tuple[np.ndarray, np.ndarray]:
        """StandardScaler -> PCA(min(pca_dim, n-1, D)) fitted on the train fold
        (utils/cv_evaluator.py:73-93), then row normalisation (K1).  Device or host, see ``preprocess``."""
        train_embeddings = np.asarray(train_embeddings)
        test_embeddings = np.asarray(test_embeddings)
        if train_embeddings.ndim == 2 and self._preprocess_on_gpu(*train_embeddings.shape):
            tr, te = self.process_embeddings_device(train_embeddings, test_embeddings)
            tr, te = tr.cpu().numpy(), te.cpu().numpy()
            if train_embeddings.dtype == np.float64:
                tr, te = tr.astype(np.float64), te.astype(np.float64)
            return tr, te
        scaler = StandardScaler()
        tr = scaler.fit_transform(train_embeddings)
        te = scaler.transform(test_embeddings)
        n_comp = min(self.pca_dim, tr.shape[0] - 1, tr.shape[1])
        if n_comp > 0:
            pca = PCA(n_components=n_comp)
            tr = pca.fit_transform(tr)
            te = pca.transform(te)
        return self._normalize_rows(tr), self._normalize_rows(te)

    def process_embeddings_device(self, train_embeddings, test_embeddings):
        """``process_embeddings`` entirely on the device (host or device arrays in, device tensors out)."""
        from .. import preprocess as pp
        eng = get_engine()
        tf = pp.fit(train_embeddings, self.pca_dim, eng)
        return pp.transform(tf, train_embeddings, eng), pp.transform(tf, test_embeddings, eng)

    # ------------------------------------------------------- GPU primitives
    def _normalize_rows(self, arr: np.ndarray) -> np.ndarray:
        """``arr / (||row|| + 1e-8)`` (utils/cv_evaluator.py:95-97) -- K1."""
        arr = np.asarray(arr)
        out = get_engine().normalize_fuse(arr, flags=native.NF_ROWNORM).f32.cpu().numpy()
        return out.astype(arr.dtype, copy=False) if arr.dtype == np.float64 else out

    def concat_fusion(self, img_vec: np.ndarray, txt_vec: np.ndarray) -> np.ndarray:
        """Image-first concat + row normalisation (utils/cv_evaluator.py:99-105) -- one K1 pass."""
        return get_engine().normalize_fuse(np.asarray(img_vec), np.asarray(txt_vec),
                                           flags=native.NF_ROWNORM).f32.cpu().numpy()

    def compute_cosine_similarity(self, query_vec: np.ndarray, db_vecs: np.ndarray) -> np.ndarray:
        """Plain ``db @ q`` on pre-normalised rows (utils/cv_evaluator.py:107-112)."""
        eng = get_engine()
        q = np.asarray(query_vec, dtype=np.float32)
        db = np.asarray(db_vecs, dtype=np.float32)
        return eng.scores(q.reshape(1, -1), db)[0].cpu().numpy()

    def retrieve_topk(self, query_vec: np.ndarray, db_vecs: np.ndarray, db_labels: List[str], top_k: int,
                      db_ids: Optional[List[str]] = None) -> Tuple[List[str], List[float], List[str]]:
        """Top-k labels / scores / ids of one query (utils/cv_evaluator.py:114-130)."""
        eng = get_engine()
        q = eng.normalize_fuse(np.asarray(query_vec).reshape(1, -1), flags=0)
        db = eng.normalize_fuse(np.asarray(db_vecs), flags=0)
        k = max(1, min(int(top_k), db.n))
        keys = eng.topk_search(q, db, k, "fp32")
        from ..engine import unpack_keys
        scores, idx = unpack_keys(keys)
        order = [int(j) for j in idx[0] if j >= 0]
        top_labels = [db_labels[j] for j in order]
        top_scores = [float(s) for s in scores[0][:len(order)]]
        top_ids = [db_ids[j] for j in order] if db_ids else [f"neighbor_{j}" for j in order]
        return top_labels, top_scores, top_ids

    def compute_vote_accuracy(self, top_labels: List[List[str]], top_scores: List[List[float]],
                              true_labels: List[str], weighted: bool = False) -> float:
        """Majority (Counter.most_common) or weighted (score-sum) vote accuracy over given
        Top-K lists (utils/cv_evaluator.py:132-155) -- K4 on synthetic keys: slot (i, j)
        becomes database row i*K + j carrying label top_labels[i][j]."""
        eng = get_engine()
        n = len(true_labels)
        if n == 0:
            raise ZeroDivisionError("division by zero")
        k = max((len(r) for r in top_labels), default=0)
        classes = sorted(set(true_labels) | {l for row in top_labels for l in row})
        lut = {c: i for i, c in enumerate(classes)}
        truth = np.fromiter((lut[l] for l in true_labels), dtype=np.int32, count=n)
        slot_labels = np.full((n, max(k, 1)), -1, dtype=np.int32)
        keys = np.zeros((n, max(k, 1)), dtype=np.uint64)
        for i, (labs, scs) in enumerate(zip(top_labels, top_scores)):
            m = min(len(labs), len(scs))
            slot_labels[i, :m] = [lut[l] for l in labs[:m]]
            s = np.asarray(scs[:m], dtype=np.float32).view(np.uint32).astype(np.uint64)
            o = np.where(s & np.uint64(0x80000000), (~s) & np.uint64(0xFFFFFFFF), s ^ np.uint64(0x80000000))
            keys[i, :m] = (o << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - (np.uint64(i * max(k, 1)) + np.arange(m, dtype=np.uint64)))
        import torch
        res = eng.vote_metrics(torch.from_numpy(keys.view(np.int64)).to(eng.device), slot_labels.reshape(-1), truth,
                               len(classes), k_list=[], wacc_f32=False, per_query=False, want_lists=False)
        good = int(res["vote_counts"][0, 2 if weighted else 1].item())
        return good / n

    # ------------------------------------------------------------------ fold
    def evaluate_fold(
        self,
        train_img: Optional[np.ndarray],
        train_txt: Optional[np.ndarray],
        test_img: Optional[np.ndarray],
        test_txt: Optional[np.ndarray],
        train_labels: List[str],
        test_labels: List[str],
        test_ids: List[str],
        fusion: str = "concat",
        top_k_list: Optional[List[int]] = None,
        w_text: float = 0.5,
        train_ids: Optional[List[str]] = None,
    ) -> Dict:
        """One CV fold; result keys as in utils/cv_evaluator.py:157-334."""
        if top_k_list is None:
            top_k_list = [1, 3, 5, self.top_k]
        tr_img = te_img = tr_txt = te_txt = None

        def process(train, test):
            shape = tuple(train.shape)
            if len(shape) == 2 and self._preprocess_on_gpu(*shape):
                return self.process_embeddings_device(train, test)      # stays in HBM for the search
            return self.process_embeddings(np.asarray(train), np.asarray(test))

        if train_img is not None and test_img is not None:
            tr_img, te_img = process(train_img, test_img)
        if train_txt is not None and test_txt is not None:
            tr_txt, te_txt = process(train_txt, test_txt)
        return self.evaluate_processed_fold(tr_img, tr_txt, te_img, te_txt, train_labels, test_labels, test_ids,
                                            fusion, top_k_list, w_text, train_ids)

    def evaluate_processed_fold(self, tr_img, tr_txt, te_img, te_txt, train_labels, test_labels, test_ids,
                                fusion="concat", top_k_list=None, w_text=0.5, train_ids=None, lists: bool = True,
                                codes=None) -> Dict:
        """The GPU part of ``evaluate_fold``: from processed (unit-row) arrays onward
        (utils/cv_evaluator.py:186-334).  ``lists=False`` leaves the per-sample python lists out;
        ``codes = (classes, train_codes, test_codes)`` passes already encoded labels (array fast path)."""
        if top_k_list is None:
            top_k_list = [1, 3, 5, self.top_k]
        eng = get_engine()
        flags_fused = native.NF_ROWNORM
        if fusion == "image_only":
            if tr_img is None or te_img is None:
                raise ValueError("image_only fusion requires image embeddings")
            db_segs, q_segs, qw, flags = (tr_img,), (te_img,), (1.0, 1.0), 0
        elif fusion == "text_only":
            if tr_txt is None or te_txt is None:
                raise ValueError("text_only fusion requires text embeddings")
            db_segs, q_segs, qw, flags = (tr_txt,), (te_txt,), (1.0, 1.0), 0
        elif fusion == "concat":
            if tr_img is None or te_img is None or tr_txt is None or te_txt is None:
                raise ValueError("concat fusion requires both image and text embeddings")
            db_segs, q_segs, qw, flags = (tr_img, tr_txt), (te_img, te_txt), (1.0, 1.0), flags_fused
        elif fusion == "late":
            if tr_img is None or te_img is None or tr_txt is None or te_txt is None:
                raise ValueError("late fusion requires both image and text embeddings")
            # w*<Tq,Td> + (1-w)*<Iq,Id> == <[(1-w) Iq ; w Tq], [Id ; Td]>: weights folded into the query rows
            db_segs, q_segs = (tr_img, tr_txt), (te_img, te_txt)
            qw, flags = (np.float32(1 - w_text), np.float32(w_text)), 0
        else:
            raise ValueError(f"Unknown fusion type: {fusion}")

        if codes is not None:
            classes, db_codes, q_codes = codes
        else:
            classes, (db_codes, q_codes) = encode(train_labels, test_labels)
        n_cls = len(classes)
        n_db, n_q = len(db_codes), len(q_codes)
        k_eff = max(1, min(int(self.top_k), n_db))
        out = eng.search_and_vote(db_segs, q_segs, db_codes, q_codes, n_cls, k_eff,
                                  db_weights=(1.0, 1.0), q_weights=qw, db_flags=flags, q_flags=flags,
                                  k_list=[int(k) for k in top_k_list], wacc_f32=False, want_lists=lists)

        results: Dict = {}
        hits = out["hit_counts"][0].cpu().numpy()
        for j, k in enumerate(top_k_list):
            results[f"top{k}"] = np.float64(hits[j]) / np.float64(n_q)
        votes = out["vote_counts"][0].cpu().numpy()
        results["vote_acc"] = int(votes[1]) / n_q
        results["weighted_vote_acc"] = int(votes[2]) / n_q

        conf = out["confusion"][0].cpu().numpy()
        prf = prf_from_confusion(conf[1], classes)           # macro metrics are over the MAJORITY vote
        results["macro_precision"] = np.mean([v["precision"] for v in prf.values()])
        results["macro_recall"] = np.mean([v["recall"] for v in prf.values()])
        results["macro_f1"] = np.mean([v["f1"] for v in prf.values()])
        results["confusion_matrix_top1"] = confusion_dict(conf[0], classes)
        results["confusion_matrix_vote"] = confusion_dict(conf[1], classes)

        if not lists:
            results["test_patient_ids"] = test_ids
            return results
        idx = out["top_idx"].cpu().numpy()
        valid = (idx >= 0).sum(axis=1)
        results["all_top_labels"] = gather_lists(train_labels, idx, valid)
        results["all_top_scores"] = score_lists(out["top_scores"].cpu().numpy(), valid)
        if train_ids:
            results["all_top_patient_ids"] = gather_lists(train_ids, idx, valid)
        else:
            results["all_top_patient_ids"] = [[f"neighbor_{j}" for j in row[:v]] for row, v in zip(idx, valid)]
        results["test_patient_ids"] = test_ids
        return results

    # ------------------------------------------------------------------- CV
    def run_cv(self, patient_ids: List[str], labels: List[str], embeddings: Dict[str, Dict[str, np.ndarray]],
               fusion: str = "concat", top_k_list: Optional[List[int]] = None, w_text: float = 0.5) -> Dict:
        """Fold loop + summary (utils/cv_evaluator.py:336-389)."""
        splits = self.stratified_split(patient_ids, labels)
        label_of = dict(zip(patient_ids, labels))
        # the embeddings dict is stacked ONCE (the reference re-stacks N dict lookups per fold, :366-371);
        # fold matrices are row gathers of these arrays -- identical values
        row_of = {p: i for i, p in enumerate(patient_ids)}
        all_img = all_txt = None
        if fusion in {"concat", "image_only", "late"}:
            all_img = np.stack([embeddings[p]["image"] for p in patient_ids])
        if fusion in {"concat", "text_only", "late"}:
            all_txt = np.stack([embeddings[p]["text"] for p in patient_ids])
        # folds preprocessed on the device: upload each modality once, fold matrices are device row gathers
        n_train0 = len(splits[0][0]) if splits else 0
        eng = None
        if all_img is not None and self._preprocess_on_gpu(n_train0, all_img.shape[1]):
            eng = get_engine()
            all_img = eng.to_device(all_img.astype(np.float32, copy=False))
        if all_txt is not None and self._preprocess_on_gpu(n_train0, all_txt.shape[1]):
            eng = get_engine()
            all_txt = eng.to_device(all_txt.astype(np.float32, copy=False))

        def rows_of(mat, rows):
            if isinstance(mat, np.ndarray):
                return mat[rows]
            import torch
            return mat.index_select(0, torch.as_tensor(rows, dtype=torch.int64, device=mat.device))

        fold_results = []
        for fold, (train_ids, test_ids) in enumerate(splits):
            logger.info(f"Processing fold {fold + 1}/{self.cv_folds}")
            logger.info(f"Train: {len(train_ids)}, Test: {len(test_ids)}")
            train_labels = [label_of[p] for p in train_ids]
            test_labels = [label_of[p] for p in test_ids]
            counts: Dict[str, int] = {}
            for lab in train_labels:
                counts[lab] = counts.get(lab, 0) + 1
            logger.info(f"Train label distribution: {counts}")

            tr_rows = [row_of[p] for p in train_ids]
            te_rows = [row_of[p] for p in test_ids]
            tr_img = te_img = tr_txt = te_txt = None
            if fusion in {"concat", "image_only", "late"}:
                tr_img, te_img = rows_of(all_img, tr_rows), rows_of(all_img, te_rows)
            if fusion in {"concat", "text_only", "late"}:
                tr_txt, te_txt = rows_of(all_txt, tr_rows), rows_of(all_txt, te_rows)
            res = self.evaluate_fold(tr_img, tr_txt, te_img, te_txt, train_labels, test_labels, test_ids,
                                     fusion, top_k_list, w_text, train_ids)
            res["fold"] = fold + 1
            res["train_ids"] = train_ids
            fold_results.append(res)
            logger.info(f"Fold {fold + 1} results: Top1={res['top1']:.4f}, "
                        f"Vote Acc={res['vote_acc']:.4f}, "
                        f"Weighted Acc={res['weighted_vote_acc']:.4f}")
        return {"fold_results": fold_results, "summary": self._compute_summary(fold_results)}

    def run_cv_arrays(self, labels, image=None, text=None, fusion: str = "concat",
                      top_k_list: Optional[List[int]] = None, w_text: float = 0.5,
                      patient_ids: Optional[List[str]] = None, lists: bool = False) -> Dict:
        """Array form of ``run_cv`` for large cohorts, the WHOLE reference pipeline on the device: the same
        StratifiedKFold split (host, sklearn), then per fold StandardScaler + PCA fitted on the train rows
        (``emr2a_b200.preprocess``, i.e. ``preprocess = "gpu"`` semantics whatever the attribute says), fusion,
        Top-K search, votes and metrics.  ``image`` / ``text``: [N, D] matrices (numpy or device tensors) in
        ``labels`` order -- no per-patient dict, no python loop over cases.  Same result structure as ``run_cv``;
        with ``lists=False`` (default) the per-sample python lists are left out."""
        import torch
        from .. import preprocess as pp
        if top_k_list is None:
            top_k_list = [1, 3, 5, self.top_k]
        need_img = fusion in {"concat", "image_only", "late"}
        need_txt = fusion in {"concat", "text_only", "late"}
        if fusion not in {"concat", "image_only", "text_only", "late"}:
            raise ValueError(f"Unknown fusion type: {fusion}")
        if need_img and image is None:
            raise ValueError(f"{fusion} fusion requires image embeddings" if fusion == "image_only"
                             else f"{fusion} fusion requires both image and text embeddings")
        if need_txt and text is None:
            raise ValueError(f"{fusion} fusion requires text embeddings" if fusion == "text_only"
                             else f"{fusion} fusion requires both image and text embeddings")
        eng = get_engine()
        n = len(labels)
        ids = list(patient_ids) if patient_ids is not None else None
        if isinstance(labels, np.ndarray) and labels.dtype.kind in "iub":
            # integer class labels: classes = sorted(set(labels)) without a python pass over the cases
            uniq, codes = np.unique(labels, return_inverse=True)
            classes, codes, label_list = uniq.tolist(), codes.astype(np.int32), labels
        else:
            label_list = labels.tolist() if isinstance(labels, np.ndarray) else list(labels)
            classes, (codes,) = encode(label_list)
        codes_t = eng.to_device(codes, torch.int32)
        mats = {}
        if need_img:
            mats["image"] = eng._embedding(image)[0].float()
        if need_txt:
            mats["text"] = eng._embedding(text)[0].float()
        skf = StratifiedKFold(n_splits=self.cv_folds, shuffle=True, random_state=self.seed)
        fold_results = []
        for fold, (tr, te) in enumerate(skf.split(np.zeros(n, dtype=np.int8), label_list)):
            logger.info(f"Processing fold {fold + 1}/{self.cv_folds}")
            logger.info(f"Train: {len(tr)}, Test: {len(te)}")
            tr_t = torch.from_numpy(tr).to(eng.device)
            te_t = torch.from_numpy(te).to(eng.device)
            proc = {}
            for name, mat in mats.items():
                x_tr = mat.index_select(0, tr_t)
                tf = pp.fit(x_tr, self.pca_dim, eng)
                proc[name] = (pp.transform(tf, x_tr, eng), pp.transform(tf, mat.index_select(0, te_t), eng))
                del x_tr
            tr_img, te_img = proc.get("image", (None, None))
            tr_txt, te_txt = proc.get("text", (None, None))
            train_labels = [label_list[j] for j in tr] if lists else None
            train_ids = [ids[j] for j in tr] if (lists and ids is not None) else None
            test_ids = [ids[j] for j in te] if ids is not None else te.tolist()
            res = self.evaluate_processed_fold(tr_img, tr_txt, te_img, te_txt, train_labels, None, test_ids, fusion,
                                               top_k_list, w_text, train_ids, lists=lists,
                                               codes=(classes, codes_t.index_select(0, tr_t), codes_t.index_select(0, te_t)))
            res["fold"] = fold + 1
            res["train_ids"] = ([ids[j] for j in tr] if ids is not None else tr.tolist()) if lists else []
            fold_results.append(res)
            logger.info(f"Fold {fold + 1} results: Top1={res['top1']:.4f}, "
                        f"Vote Acc={res['vote_acc']:.4f}, "
                        f"Weighted Acc={res['weighted_vote_acc']:.4f}")
        return {"fold_results": fold_results, "summary": self._compute_summary(fold_results)}

    def run_cv_processed(self, patient_ids: List[str], labels: List[str], image: Optional[np.ndarray],
                         text: Optional[np.ndarray], fusion: str = "concat", top_k_list: Optional[List[int]] = None,
                         w_text: float = 0.5, lists: bool = True) -> Dict:
        """Array fast path of ``run_cv`` for embeddings that are ALREADY processed (unit rows; no per-fold
        scaler/PCA): the same StratifiedKFold split, then all folds in ONE fold-masked GPU pass
        (``Engine.cv_search_and_vote``) instead of one database upload + search per fold.  Returns the
        same ``{"fold_results": [...], "summary": {...}}`` structure; with ``lists=False`` the per-sample
        python lists (``all_top_*``) are left out (at 10M cases they dwarf the GPU time)."""
        if top_k_list is None:
            top_k_list = [1, 3, 5, self.top_k]
        if fusion == "image_only":
            if image is None:
                raise ValueError("image_only fusion requires image embeddings")
            segs, qw, flags = (image,), (1.0, 1.0), 0
        elif fusion == "text_only":
            if text is None:
                raise ValueError("text_only fusion requires text embeddings")
            segs, qw, flags = (text,), (1.0, 1.0), 0
        elif fusion == "concat":
            if image is None or text is None:
                raise ValueError("concat fusion requires both image and text embeddings")
            segs, qw, flags = (image, text), (1.0, 1.0), native.NF_ROWNORM
        elif fusion == "late":
            if image is None or text is None:
                raise ValueError("late fusion requires both image and text embeddings")
            segs, qw, flags = (image, text), (np.float32(1 - w_text), np.float32(w_text)), 0
        else:
            raise ValueError(f"Unknown fusion type: {fusion}")
        n = len(labels)
        folds = np.zeros(n, dtype=np.uint8)
        skf = StratifiedKFold(n_splits=self.cv_folds, shuffle=True, random_state=self.seed)
        members = []
        for f, (_, te) in enumerate(skf.split(patient_ids, labels)):
            folds[te] = f
            members.append(te)
        classes, (codes,) = encode(labels)
        n_cls = len(classes)
        eng = get_engine()
        k_eff = max(1, min(int(self.top_k), n - max(len(m) for m in members)))
        out = eng.cv_search_and_vote(segs, codes, folds, n_cls, k_eff, flags=flags, q_weights=qw,
                                     k_list=[int(k) for k in top_k_list], n_folds=self.cv_folds, want_lists=lists)
        hits = out["hit_counts"].cpu().numpy()
        votes = out["vote_counts"].cpu().numpy()
        conf = out["confusion"].cpu().numpy()
        if lists:
            idx_all = out["top_idx"].cpu().numpy()
            sc_all = out["top_scores"].cpu().numpy()
        ids_arr = np.asarray(patient_ids, dtype=object)
        fold_results = []
        for f, te in enumerate(members):
            n_q = len(te)
            r: Dict = {}
            for j, k in enumerate(top_k_list):
                r[f"top{k}"] = np.float64(hits[f, j]) / np.float64(n_q)
            r["vote_acc"] = int(votes[f, 1]) / n_q
            r["weighted_vote_acc"] = int(votes[f, 2]) / n_q
            prf = prf_from_confusion(conf[f, 1], classes)
            r["macro_precision"] = np.mean([v["precision"] for v in prf.values()])
            r["macro_recall"] = np.mean([v["recall"] for v in prf.values()])
            r["macro_f1"] = np.mean([v["f1"] for v in prf.values()])
            r["confusion_matrix_top1"] = confusion_dict(conf[f, 0], classes)
            r["confusion_matrix_vote"] = confusion_dict(conf[f, 1], classes)
            if lists:
                idx = idx_all[te]
                valid = (idx >= 0).sum(axis=1)
                r["all_top_labels"] = gather_lists(labels, idx, valid)
                r["all_top_scores"] = score_lists(sc_all[te], valid)
                r["all_top_patient_ids"] = gather_lists(patient_ids, idx, valid)
            r["test_patient_ids"] = ids_arr[te].tolist()
            r["fold"] = f + 1
            r["train_ids"] = ids_arr[np.setdiff1d(np.arange(n), te)].tolist() if lists else []
            fold_results.append(r)
        return {"fold_results": fold_results, "summary": self._compute_summary(fold_results)}

    def _compute_summary(self, all_results: List[Dict]) -> Dict:
        """mean / std / min / max over folds of the 8 fixed metrics (utils/cv_evaluator.py:391-405)."""
        summary = {}
        for name in _SUMMARY_METRICS:
            vals = [r[name] for r in all_results]
            summary[name] = {"mean": float(np.mean(vals)), "std": float(np.std(vals)),
                             "min": float(np.min(vals)), "max": float(np.max(vals))}
        return summary

    # ---------------------------------------------------------------- output
    def save_results(self, results: Dict, output_dir: Path, experiment_id: str, config: Dict):
        """config.json, fold_k/metrics.json, summary.csv, confusion_matrices.png
        (utils/cv_evaluator.py:407-441)."""
        exp_dir = Path(output_dir) / f"exp_{experiment_id}"
        exp_dir.mkdir(parents=True, exist_ok=True)
        with (exp_dir / "config.json").open("w", encoding="utf-8") as fh:
            json.dump(config, fh, ensure_ascii=False, indent=2)
        for fold_result in results["fold_results"]:
            fold_dir = exp_dir / f"fold_{fold_result['fold']}"
            fold_dir.mkdir(exist_ok=True)
            with (fold_dir / "metrics.json").open("w", encoding="utf-8") as fh:
                json.dump(self._make_serializable(fold_result), fh, ensure_ascii=False, indent=2)
        self._save_summary_csv(results["summary"], exp_dir / "summary.csv")
        if "vlm_review" in results:
            with (exp_dir / "vlm_review_summary.json").open("w", encoding="utf-8") as fh:
                json.dump(results["vlm_review"], fh, ensure_ascii=False, indent=2)
        self._plot_confusion_matrices(results, exp_dir)
        logger.info(f"Results saved to {exp_dir}")

    def _save_summary_csv(self, summary: Dict, output_path: Path):
        with Path(output_path).open("w", newline="", encoding="utf-8") as fh:
            w = csv.writer(fh)
            w.writerow(["Metric", "Mean", "Std", "Min", "Max"])
            for name, st in summary.items():
                w.writerow([name] + [f"{st[key]:.4f}" for key in ("mean", "std", "min", "max")])

    def _plot_confusion_matrices(self, results: Dict, output_dir: Path):
        """Fold-averaged confusion heat maps (utils/cv_evaluator.py:459-499).  matplotlib and
        seaborn are optional here: without them the PNG is skipped with a warning."""
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
            import seaborn as sns
        except Exception as exc:                                     # pragma: no cover - plotting only
            logger.warning(f"confusion_matrices.png skipped (plotting libraries unavailable: {exc})")
            return
        names = sorted({k for r in results["fold_results"] for k in r["confusion_matrix_top1"].keys()})
        folds = results["fold_results"]

        def mean_matrix(key):
            return sum(np.array([[r[key][t][p] for p in names] for t in names], dtype=float) for r in folds) / len(folds)

        fig, axes = plt.subplots(1, 2, figsize=(12, 5))
        for ax, key, title in ((axes[0], "confusion_matrix_top1", "Confusion Matrix (Top1)"),
                               (axes[1], "confusion_matrix_vote", "Confusion Matrix (Vote)")):
            sns.heatmap(mean_matrix(key), annot=True, fmt=".1f", cmap="Blues", xticklabels=names, yticklabels=names, ax=ax)
            ax.set_title(title)
            ax.set_xlabel("Predicted")
            ax.set_ylabel("True")
        plt.tight_layout()
        plt.savefig(Path(output_dir) / "confusion_matrices.png", dpi=150, bbox_inches="tight")
        plt.close()
        logger.info(f"Confusion matrices saved to {Path(output_dir) / 'confusion_matrices.png'}")
