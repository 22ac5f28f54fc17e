"""Per-stage CUDA-event breakdown of one C2 step of the row-sharded search under torchrun (rank 0 prints).
Stages are bracketed with events on the launching stream by wrapping the engine methods and the two collectives of
emr2a_b200/dist.py; the step itself is unchanged.  EMR2A_COOP_SHARDS=0 shows the self-contained shards."""
import os, sys, collections
import torch, torch.distributed as dist
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from emr2a_b200 import native, synth
import emr2a_b200.dist as ed
from emr2a_b200.engine import get_engine

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
eng = get_engine(dev)
n_db, d, n_q, k, c, seed = 1_000_000, 512, 10_000, 10, 3, 1234
lo, hi = ed.shard_range(n_db, rank, world)
flags = native.NF_SEGNORM | native.NF_ROWNORM
di, _ = synth.device_block(lo, hi - lo, d, c, seed, dev, label_seed=seed)
dt, _ = synth.device_block(lo, hi - lo, d, c, seed + 1, dev, label_seed=seed)
labels = synth.device_labels(0, n_db, c, seed, dev)
qi, ql = synth.device_block(50_003_968, n_q, d, c, seed, dev, label_seed=seed)
qt, _ = synth.device_block(50_003_968, n_q, d, c, seed + 1, dev, label_seed=seed)

spans = collections.defaultdict(list)
def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    def inner(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(*a, **kw); e1.record()
        spans[label or name].append((e0, e1))
        return out
    setattr(obj, name, inner)
for m in ("prepare", "topk_search", "topk_filter", "rescore_candidates", "merge_payload", "verify_merged", "topk_merge", "vote_metrics"):
    wrap(eng, m)
wrap(ed, "gather_keys")
if world > 1:
    wrap(ed.dist, "all_reduce")

def step():
    return ed.sharded_search_and_vote(eng, (di, dt), (qi, qt), labels, ql, c, k, lo, flags, flags, k_list=[1, 3, 5, k], defer_status=True)
for _ in range(5):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
spans.clear()
steps = 50
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(steps):
    step()
t1.record()
torch.cuda.synchronize()
total = t0.elapsed_time(t1) / steps
if rank == 0:
    print(f"world {world} coop={os.environ.get('EMR2A_COOP_SHARDS', '1')}: {total:.3f} ms/step")
    acc = 0.0
    for name, evs in spans.items():
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        acc += ms
        print(f"  {name:20s} {ms:8.3f} ms  ({len(evs) // steps} calls/step)")
    print(f"  {'(between stages)':20s} {total - acc:8.3f} ms")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
