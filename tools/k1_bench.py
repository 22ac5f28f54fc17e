"""K1 (normalize+fuse) alone: achieved HBM GB/s per output mode."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from emr2a_b200 import native, synth
from emr2a_b200.engine import get_engine
eng = get_engine(); dev = eng.device
n, d = int(os.environ.get("N", 1_000_000)), int(os.environ.get("D", 512))
d2 = int(os.environ.get("D2", d))
a, _ = synth.device_block(0, n, d, 3, 11, dev); b, _ = synth.device_block(0, n, d2, 3, 12, dev)
isz = 4
if os.environ.get("DTYPE") == "bf16":
    a, b, isz = a.to(torch.bfloat16), b.to(torch.bfloat16), 2
flags = native.NF_SEGNORM | native.NF_ROWNORM
D = d + d2
modes = {"f32 out (API)": dict(want_f32=True, want_planes=False), "hi+lo planes (bf16x3)": dict(want_f32=False, want_planes=True, want_lo=True),
         "f32+hi+stats (rescore)": dict(want_f32=True, want_planes=True, want_lo=False, want_stats=True),
         "hi only (bf16x1)": dict(want_f32=False, want_planes=True, want_lo=False)}
bytes_per_row = {"f32 out (API)": D * isz + D * 4, "hi+lo planes (bf16x3)": D * isz + D * 4, "f32+hi+stats (rescore)": D * isz + D * 4 + D * 2,
                 "hi only (bf16x1)": D * isz + D * 2}
for name, kw in modes.items():
    for _ in range(3):
        o = eng.normalize_fuse(a, b, 1.0, 1.0, flags, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        o = eng.normalize_fuse(a, b, 1.0, 1.0, flags, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = n * bytes_per_row[name] / 1e9
    print(f"K1 {name}: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s algorithmic ({gb:.2f} GB)  = {gb/ms*1e3/6549.4:.2f} of measured HBM peak", flush=True)
