"""``retrieval.evaluator.RetrievalEvaluator`` on the B200.

Same constructor, methods, argument order, defaults, result keys and errors as
the reference class (retrieval/evaluator.py:12-275).  The per-query python loops
of the reference (cosine -> argsort -> label compare, once per k) are replaced by
one K1 -> K2 -> K4 pass per metric family on the GPU; StandardScaler / PCA and the
python-``random`` split stay on the host exactly as in the reference.
"""
import os
import random
from typing import Dict, List, Optional, Tuple

import numpy as np
from sklearn.decomposition import PCA
from sklearn.preprocessing import StandardScaler

from .. import native
from ..engine import get_engine
from ..labels import encode
from .fusion import _MODES

_WEIGHTED_K = 5      # the reference hard-codes Top-5 for the weighted vote (retrieval/evaluator.py:220,243)


class RetrievalEvaluator:
    def __init__(self, test_ratio: float = 0.2, seed: int = 42, use_pca: bool = False, pca_dim: int = 128):
        self.test_ratio = test_ratio
        self.seed = seed
        self.use_pca = use_pca
        self.pca_dim = pca_dim
        self.rng = random.Random(seed)

    # ------------------------------------------------------------------ host
    def stratified_split(self, labels: List[str]) -> Tuple[List[int], List[int]]:
        """Per-class shuffle with the evaluator's own ``random.Random`` then a
        ``round(n * test_ratio)`` cut clamped to [1, n-1]; singleton classes go to
        train (retrieval/evaluator.py:26-48).  Host logic: the RNG stream must match."""
        groups: Dict[str, List[int]] = {}
        for pos, lab in enumerate(labels):
            groups.setdefault(str(lab), []).append(pos)
        train: List[int] = []
        test: List[int] = []
        for members in groups.values():
            self.rng.shuffle(members)
            if len(members) <= 1:
                train += members
                continue
            cut = int(round(len(members) * self.test_ratio))
            cut = max(1, min(cut, len(members) - 1))
            test += members[:cut]
            train += members[cut:]
        return train, test

    preprocess: str = os.environ.get("EMR2A_PREPROCESS", "auto")

    def _preprocess_on_gpu(self, n_train: int, n_features: int) -> bool:
        mode = self.preprocess
        if mode not in ("host", "gpu", "auto"):
            raise ValueError(f"unknown preprocess mode {mode!r} (host, gpu, auto)")
        if mode != "auto":
            return mode == "gpu"
        from ..preprocess import sklearn_solver
        n_comp = min(self.pca_dim, n_train - 1, n_features) if self.use_pca else 0
        return n_comp <= 0 or sklearn_solver(n_train, n_features, n_comp) != "full"      # see CVRetrievalEvaluator.preprocess

    def process_embeddings(self, train_embeddings: np.ndarray, test_embeddings: np.ndarray
                           ) -> Tuple[np.ndarray, np.ndarray]:
        """StandardScaler (+ PCA when ``use_pca``) fitted on train (retrieval/evaluator.py:50-73), row
        normalisation by K1.  Runs on the device under the same rule as
        ``CVRetrievalEvaluator.preprocess`` (env ``EMR2A_PREPROCESS``: host / gpu / auto), else sklearn on the host."""
        train_embeddings = np.asarray(train_embeddings)
        test_embeddings = np.asarray(test_embeddings)
        if train_embeddings.ndim == 2 and self._preprocess_on_gpu(*train_embeddings.shape):
            from .. import preprocess as pp
            eng = get_engine()
            tf = pp.fit(train_embeddings, self.pca_dim if self.use_pca else None, eng)
            tr = pp.transform(tf, train_embeddings, eng).cpu().numpy()
            te = pp.transform(tf, test_embeddings, eng).cpu().numpy()
            if train_embeddings.dtype == np.float64:
                tr, te = tr.astype(np.float64), te.astype(np.float64)
            return tr, te
        scaler = StandardScaler()
        tr = scaler.fit_transform(train_embeddings)
        te = scaler.transform(test_embeddings)
        if self.use_pca:
            n_comp = min(self.pca_dim, tr.shape[0] - 1, tr.shape[1])
            if n_comp > 0:
                pca = PCA(n_components=n_comp)
                tr = pca.fit_transform(tr)
                te = pca.transform(te)
        return self._normalize_rows(tr), self._normalize_rows(te)

    def _normalize_rows(self, arr: np.ndarray) -> np.ndarray:
        """``arr / (||row|| + 1e-8)`` by the K1 kernel (retrieval/evaluator.py:75-77)."""
        arr = np.asarray(arr)
        out = get_engine().normalize_fuse(arr, flags=native.NF_ROWNORM).f32.cpu().numpy()
        return out.astype(arr.dtype, copy=False) if arr.dtype == np.float64 else out

    def align_dims(self, train_text, test_text, train_image, test_image):
        """retrieval/evaluator.py:79-92."""
        if train_text is not None and test_text is not None:
            train_text, test_text = self.process_embeddings(train_text, test_text)
        if train_image is not None and test_image is not None:
            train_image, test_image = self.process_embeddings(train_image, test_image)
        return train_text, test_text, train_image, test_image

    # ------------------------------------------------------------------- GPU
    @staticmethod
    def _family(eng, operands, db_codes, q_codes, n_classes, top_k_list, prec):
        """One search at K = max(k list, 5) serves every top-k accuracy and the
        Top-5 weighted vote of a metric family.  ``operands(prec) -> (db_op, q_op)`` runs K1 for a precision arm:
        when the rescore arm cannot verify more queries than its exact re-scan list holds (dense score
        neighbourhoods, duplicated cases) the search is repeated with the 3-pass tensor-core arm, as every other
        caller of ``topk_search`` does -- an unverified filter result is never reported."""
        db_op, q_op = operands(prec)
        n_db = db_op.n
        k_max = max(max(top_k_list), _WEIGHTED_K)
        k_eff = min(k_max, n_db) if n_db > 0 else 1
        keys = eng.topk_search(q_op, db_op, k_eff, prec)
        if prec == "rescore":
            _, overflow = eng.consume_status()
            if overflow:
                db_op, q_op = operands("bf16x3")
                keys = eng.topk_search(q_op, db_op, k_eff, "bf16x3")
        return RetrievalEvaluator._family_from_keys(eng, keys, db_codes, q_codes, n_classes, top_k_list)

    @staticmethod
    def _family_from_keys(eng, keys, db_codes, q_codes, n_classes, top_k_list):
        k_eff = int(keys.shape[1])
        hits = eng.vote_metrics(keys, db_codes, q_codes, n_classes, k_list=list(top_k_list), wacc_f32=True,
                                per_query=False, want_lists=False)
        k5 = min(_WEIGHTED_K, k_eff)
        w = eng.vote_metrics(keys[:, :k5].contiguous(), db_codes, q_codes, n_classes, k_list=[], wacc_f32=True,
                             per_query=False, want_lists=True)
        n_q = max(len(q_codes), 1)
        accs = (hits["hit_counts"][0].cpu().numpy() / len(q_codes)) if len(q_codes) else np.zeros(len(top_k_list))
        weighted = int(w["vote_counts"][0, 2].item()) / n_q
        return [float(a) for a in accs], weighted, w["top_labels"]

    def _cosine_family(self, db, qs, db_codes, q_codes, n_classes, top_k_list):
        """cosine (both sides re-normalised with epsilon, retrieval/similarity.py:4-7) -> metrics."""
        eng = get_engine()
        prec = eng.pick_precision(len(qs), len(db), db.shape[1], max(max(top_k_list), _WEIGHTED_K))

        def operands(p):
            return eng.prepare(db, flags=native.NF_ROWNORM, precision=p), eng.prepare(qs, flags=native.NF_ROWNORM, precision=p)
        return self._family(eng, operands, db_codes, q_codes, n_classes, top_k_list, prec)

    def evaluate_retrieval(
        self,
        train_text: Optional[np.ndarray],
        test_text: Optional[np.ndarray],
        train_image: Optional[np.ndarray],
        test_image: Optional[np.ndarray],
        train_labels: List[str],
        test_labels: List[str],
        text_weight: float = 0.4,
        fusion_type: str = "late",
        score_mode: str = "none",
        top_k_list: List[int] = [1, 3, 5],
    ) -> Dict:
        """Result keys as in retrieval/evaluator.py:94-176."""
        eng = get_engine()
        classes, (db_codes, q_codes) = encode(train_labels, test_labels)
        n_cls = len(classes)
        results: Dict = {}
        if fusion_type == "early":
            if train_text is None or test_text is None or train_image is None or test_image is None:
                raise ValueError("Early fusion requires both text and image embeddings")
            # early_fusion (text first, weights w / 1-w, row-normalised) then cosine
            tw, iw = np.float32(text_weight), np.float32(1 - text_weight)
            fused_db = eng.normalize_fuse(train_text, train_image, tw, iw, native.NF_ROWNORM).f32
            fused_q = eng.normalize_fuse(test_text, test_image, tw, iw, native.NF_ROWNORM).f32
            accs, weighted, _ = self._cosine_family(fused_db, fused_q, db_codes, q_codes, n_cls, top_k_list)
            for k, a in zip(top_k_list, accs):
                results[f"top{k}"] = a
            results["weighted"] = weighted
            return results

        if test_text is not None and train_text is not None:
            accs, weighted, _ = self._cosine_family(np.asarray(train_text), np.asarray(test_text), db_codes, q_codes,
                                                    n_cls, top_k_list)
            for k, a in zip(top_k_list, accs):
                results[f"text_top{k}"] = a
            results["text_weighted"] = weighted
        if test_image is not None and train_image is not None:
            accs, weighted, _ = self._cosine_family(np.asarray(train_image), np.asarray(test_image), db_codes, q_codes,
                                                    n_cls, top_k_list)
            for k, a in zip(top_k_list, accs):
                results[f"image_top{k}"] = a
            results["image_weighted"] = weighted
        if test_text is not None and test_image is not None:
            mode = _MODES.get(score_mode, native.SCORE_NONE)
            if mode == native.SCORE_NONE:
                # w*cos_T + (1-w)*cos_I as ONE contraction: unit segments, weights folded into the query side
                n_db, dim = len(train_labels), train_text.shape[1] + train_image.shape[1]
                prec = eng.pick_precision(len(test_labels), n_db, dim, max(max(top_k_list), _WEIGHTED_K))

                def operands(p):
                    return (eng.prepare(train_text, train_image, 1.0, 1.0, native.NF_SEGNORM, p),
                            eng.prepare(test_text, test_image, np.float32(text_weight), np.float32(1 - text_weight),
                                        native.NF_SEGNORM, p))
                accs, weighted, top5 = self._family(eng, operands, db_codes, q_codes, n_cls, top_k_list, prec)
            elif self._late_fused(len(test_labels), len(train_labels)):
                # z-score / min-max without the [Q, N] score matrix: per-query statistics from database moments /
                # K = 1 searches, then the ordinary fused search with scaled query segments (emr2a_b200/late.py)
                from ..late import late_fusion_search
                k_eff = min(max(max(top_k_list), _WEIGHTED_K), len(train_labels))
                keys = late_fusion_search(np.asarray(train_text), np.asarray(train_image), np.asarray(test_text),
                                          np.asarray(test_image), text_weight, mode, k_eff, engine=eng)
                accs, weighted, top5 = self._family_from_keys(eng, keys, db_codes, q_codes, n_cls, top_k_list)
            else:
                # small problems: materialise [Q, N] and apply the reference's elementwise arithmetic op for op
                fused = self._late_score_matrix(train_text, test_text, train_image, test_image, text_weight, mode)
                accs = [self._topk_acc_from_device_scores(fused, db_codes, q_codes, n_cls, k) for k in top_k_list]
                weighted, top5 = self._weighted_from_device_scores(fused, db_codes, q_codes, n_cls)
            for k, a in zip(top_k_list, accs):
                results[f"top{k}"] = a
            results["weighted"] = weighted
            lab = np.asarray(classes, dtype=object)
            t5 = top5.cpu().numpy()
            results["all_top_labels_top5"] = [[lab[c] for c in row if c >= 0] for row in t5]
        return results

    #: late fusion with z-score / min-max: score matrices beyond this many elements are never materialised
    #: (env ``EMR2A_LATE_FUSED=1`` / ``0`` forces the fused / materialised path)
    late_materialise_max: int = 1 << 26

    def _late_fused(self, n_q: int, n_db: int) -> bool:
        forced = os.environ.get("EMR2A_LATE_FUSED")
        if forced in ("0", "1"):
            return forced == "1"
        return n_q * n_db > self.late_materialise_max

    def _late_score_matrix(self, train_text, test_text, train_image, test_image, text_weight, mode):
        eng = get_engine()
        st = eng.scores(eng.normalize_fuse(test_text, flags=native.NF_ROWNORM).f32,
                        eng.normalize_fuse(train_text, flags=native.NF_ROWNORM).f32)
        si = eng.scores(eng.normalize_fuse(test_image, flags=native.NF_ROWNORM).f32,
                        eng.normalize_fuse(train_image, flags=native.NF_ROWNORM).f32)
        return eng.late_fuse_scores(st, si, text_weight, mode)

    @staticmethod
    def _topk_acc_from_device_scores(scores, db_codes, q_codes, n_cls, k) -> float:
        eng = get_engine()
        keys = eng.topk_from_scores(scores, min(k, scores.shape[1]))
        r = eng.vote_metrics(keys, db_codes, q_codes, n_cls, k_list=[k], wacc_f32=True, per_query=False, want_lists=False)
        return int(r["hit_counts"][0, 0].item()) / len(q_codes)

    @staticmethod
    def _weighted_from_device_scores(scores, db_codes, q_codes, n_cls):
        eng = get_engine()
        keys = eng.topk_from_scores(scores, min(_WEIGHTED_K, scores.shape[1]))
        r = eng.vote_metrics(keys, db_codes, q_codes, n_cls, k_list=[], wacc_f32=True, per_query=False, want_lists=True)
        return int(r["vote_counts"][0, 2].item()) / len(q_codes), r["top_labels"]

    # --- the reference's helper methods, same signatures (retrieval/evaluator.py:178-275) ---
    def _compute_top_k_accuracy(self, train_embeddings, test_embeddings, train_labels, test_labels, top_k) -> float:
        classes, (db_codes, q_codes) = encode(train_labels, test_labels)
        accs, _, _ = self._cosine_family(np.asarray(train_embeddings), np.asarray(test_embeddings), db_codes, q_codes,
                                         len(classes), [top_k])
        return accs[0]

    def _compute_top_k_accuracy_from_scores(self, scores, train_labels, test_labels, top_k) -> float:
        classes, (db_codes, q_codes) = encode(train_labels, test_labels)
        dev = get_engine().to_device(np.asarray(scores, dtype=np.float32))
        return self._topk_acc_from_device_scores(dev, db_codes, q_codes, len(classes), top_k)

    def _compute_weighted_accuracy(self, train_embeddings, test_embeddings, train_labels, test_labels) -> float:
        classes, (db_codes, q_codes) = encode(train_labels, test_labels)
        _, weighted, _ = self._cosine_family(np.asarray(train_embeddings), np.asarray(test_embeddings), db_codes,
                                             q_codes, len(classes), [_WEIGHTED_K])
        return weighted

    def _compute_weighted_accuracy_from_scores(self, scores, train_labels, test_labels) -> float:
        classes, (db_codes, q_codes) = encode(train_labels, test_labels)
        dev = get_engine().to_device(np.asarray(scores, dtype=np.float32))
        return self._weighted_from_device_scores(dev, db_codes, q_codes, len(classes))[0]

    def get_all_top_labels(self, scores, train_labels, test_labels, top_k: int = 5) -> List[List[str]]:
        """Top-k labels of every row of a score matrix (retrieval/evaluator.py:258-275)."""
        eng = get_engine()
        dev = eng.to_device(np.asarray(scores, dtype=np.float32))
        keys = eng.topk_from_scores(dev, min(top_k, dev.shape[1]))
        from ..engine import unpack_keys
        _, idx = unpack_keys(keys)
        return [[train_labels[j] for j in row if j >= 0] for row in idx]
