"""Run under torchrun: sharded search at world_size W must give bit-identical keys to a
single-GPU search of the whole database (global-index tie rule)."""
import os, sys
import torch, torch.distributed as dist
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from emr2a_b200 import native, synth
from emr2a_b200.dist import fold_balanced_ranges, ranges_to_rows, shard_range, sharded_cv_search_and_vote, sharded_search_and_vote
from emr2a_b200.engine import get_engine

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
eng = get_engine(dev)
n, d, q, k, c = 300_000, 256, 3000, 10, 3
flags = native.NF_SEGNORM | native.NF_ROWNORM
labels = synth.device_labels(0, n, c, 11, dev)
qi, ql = synth.device_block(50_003_968, q, d, c, 11, dev, label_seed=11)
qt, _ = synth.device_block(50_003_968, q, d, c, 12, dev, label_seed=11)
ok = True
for prec in ("rescore", "rescore-selfcontained", "bf16x3", "fp32"):
    coop = prec == "rescore"                      # cooperative shards (staged K2, two collectives) vs self-contained ones
    name, prec = prec, prec.split("-")[0]
    lo, hi = shard_range(n, rank, world)
    di, _ = synth.device_block(lo, hi - lo, d, c, 11, dev, label_seed=11)
    dt, _ = synth.device_block(lo, hi - lo, d, c, 12, dev, label_seed=11)
    r = sharded_search_and_vote(eng, (di, dt), (qi, qt), labels, ql, c, k, lo, flags, flags, precision=prec, cooperative=coop)
    fi, _ = synth.device_block(0, n, d, c, 11, dev, label_seed=11)
    ft, _ = synth.device_block(0, n, d, c, 12, dev, label_seed=11)
    full = eng.search_and_vote((fi, ft), (qi, qt), labels, ql, c, k, db_flags=flags, q_flags=flags, precision=prec)
    same = torch.equal(r["keys"], full["keys"]) and torch.equal(r["pred_vote"], full["pred_vote"]) and torch.equal(r["confusion"], full["confusion"])
    print(f"rank {rank}/{world} {name}: sharded == single-GPU: {same} (unverified {r.get('unverified')})", flush=True)
    ok = ok and same
# all-queries CV over the sharded cohort == the single-GPU one-pass CV
n_cv, n_folds = 200_000, 5
lo, hi = shard_range(n_cv, rank, world)
xi, _ = synth.device_block(lo, hi - lo, d, c, 19, dev, label_seed=19)
xt, _ = synth.device_block(lo, hi - lo, d, c, 20, dev, label_seed=19)
lab = synth.device_labels(0, n_cv, c, 19, dev)
fold = (torch.arange(n_cv, device=dev, dtype=torch.int64) * n_folds // n_cv).to(torch.uint8)
fi, _ = synth.device_block(0, n_cv, d, c, 19, dev, label_seed=19)
ft, _ = synth.device_block(0, n_cv, d, c, 20, dev, label_seed=19)
for prec in ("rescore", "bf16x3"):
    r = sharded_cv_search_and_vote(eng, (xi, xt), lab, fold, c, 5, lo, flags, k_list=[1, 3, 5], precision=prec,
                                   n_folds=n_folds, q_block=65536, want_lists=True)
    full = eng.cv_search_and_vote((fi, ft), lab, fold, c, 5, flags=flags, k_list=[1, 3, 5], precision=prec, n_folds=n_folds,
                                  distributed=False)
    names = ("hit_counts", "vote_counts", "confusion", "group_sizes", "top_idx", "top_scores", "pred_vote", "pred_weighted")
    same = all(torch.equal(r[key], full[key]) for key in names)
    print(f"rank {rank}/{world} cv {prec}: sharded == single-GPU: {same} (unverified {r['unverified']})", flush=True)
    ok = ok and same
    # fold-balanced shards (a slice of every fold per rank; the C5 layout of bench.py), rows generated per range
    ranges = fold_balanced_ranges([int((fold == f).sum()) for f in range(n_folds)], rank, world)
    bi = torch.cat([synth.device_block(g0, cnt, d, c, 19, dev, label_seed=19)[0] for g0, cnt in ranges])
    bt = torch.cat([synth.device_block(g0, cnt, d, c, 20, dev, label_seed=19)[0] for g0, cnt in ranges])
    r = sharded_cv_search_and_vote(eng, (bi, bt), lab, fold, c, 5, 0, flags, k_list=[1, 3, 5], precision=prec,
                                   n_folds=n_folds, q_block=16384, want_lists=True, row_ids=ranges_to_rows(ranges, dev))
    same = all(torch.equal(r[key], full[key]) for key in names)
    print(f"rank {rank}/{world} cv {prec}: fold-balanced shards == single-GPU: {same}", flush=True)
    ok = ok and same
    # the engine call itself goes multi-GPU under torchrun (every rank passes the same arrays), here with UNSORTED folds
    mixed = ((torch.arange(n_cv, device=dev, dtype=torch.int64) * 7919) % n_folds).to(torch.uint8)
    auto = eng.cv_search_and_vote((fi, ft), lab, mixed, c, 5, flags=flags, k_list=[1, 3, 5], precision=prec, n_folds=n_folds,
                                  distributed=True)
    single = eng.cv_search_and_vote((fi, ft), lab, mixed, c, 5, flags=flags, k_list=[1, 3, 5], precision=prec, n_folds=n_folds,
                                    distributed=False)
    same = all(torch.equal(auto[key], single[key]) for key in names)
    print(f"rank {rank}/{world} cv {prec}: engine auto-distributed == single-GPU: {same}", flush=True)
    ok = ok and same
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
