"""Kernel timeline of one graph-replayed training step via torch.profiler (CUPTI): per-stream busy time, how much of
the side-stream (transformer branch) work overlaps main-stream kernels, and the gaps of the main stream."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from oracle import hdf_oracle as O
from hdenseformer_b200 import trainer as T
from hdenseformer_b200.loss import CEPlusDice, DeepSuperloss
from hdenseformer_b200.models import HDenseFormer_32

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
size = (144, 144, 144)
net = HDenseFormer_32(2, 2, size, 12).to(dev).train()
crit = DeepSuperloss(CEPlusDice(ignore_index=0))
opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True, capturable=True)
x, t = O.synth_petct(2, size, seed=0).to(dev), O.synth_label(2, 2, size, seed=0).to(dev)
g = T.GraphedTrainStep(net, crit, opt, x, t, use_bf16=True)
for _ in range(3): g.step(x, t)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    g.step(x, t)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
ks = sorted([(e.time_range.start, e.time_range.end, e.name, getattr(e, "cuda_stream", None) if hasattr(e, "cuda_stream") else None) for e in ev])
print("kernels", len(ks))
if not ks: sys.exit(0)
t0 = ks[0][0]; t1 = max(k[1] for k in ks)
print("span ms", (t1 - t0) / 1e3)
# union busy time
busy = 0; cur_s, cur_e = ks[0][0], ks[0][1]
for s, e, _, _ in ks[1:]:
    if s > cur_e: busy += cur_e - cur_s; cur_s, cur_e = s, e
    else: cur_e = max(cur_e, e)
busy += cur_e - cur_s
print("union busy ms", busy / 1e3, "sum of durations ms", sum(e - s for s, e, _, _ in ks) / 1e3)
big = [(s, e, n) for s, e, n, _ in ks if ("tc_conv" in n)]
small = [(s, e, n) for s, e, n, _ in ks if ("tc_conv" not in n)]
print("tc_conv kernels", len(big), "sum ms", sum(e - s for s, e, _ in big) / 1e3)
# how much small-kernel time falls inside some tc_conv kernel interval
ov = 0
for s, e, n in small:
    for bs, be, _ in big:
        lo, hi = max(s, bs), min(e, be)
        if hi > lo: ov += hi - lo
print("non-conv kernel time ms", sum(e - s for s, e, _ in small) / 1e3, "of which overlapping a tc_conv kernel ms", ov / 1e3)
# conv kernel durations: in-step vs list
import collections
d = collections.defaultdict(list)
for s, e, n in big: d[n[:40]].append((e - s) / 1e3)
for k, v in d.items(): print(k, "n", len(v), "sum", round(sum(v), 2), "max", round(max(v), 3))
# biggest gaps where nothing runs
gaps = []; cur_e = ks[0][1]
for s, e, n, _ in ks[1:]:
    if s > cur_e: gaps.append((s - cur_e, (cur_e - t0) / 1e3, n[:50]))
    cur_e = max(cur_e, e)
gaps.sort(reverse=True)
print("idle total ms", sum(g_[0] for g_ in gaps) / 1e3, "largest gaps (us, at ms, next kernel):", [(round(a, 1), round(b, 2), c) for a, b, c in gaps[:8]])

# ---- forward-phase markers (ms from step start)
def first(name, nth=0):
    c = [k for k in ks if name in k[2]]
    return ((c[nth][0] - t0) / 1e3, (c[nth][1] - t0) / 1e3) if len(c) > nth else None
def last(name):
    c = [k for k in ks if name in k[2]]
    return ((c[-1][0] - t0) / 1e3, (c[-1][1] - t0) / 1e3) if c else None
print("patch-embed gemm (PatchA) first/last:", first("PatchA"), last("PatchA,"))
af = [k for k in ks if "attn_fwd" in k[2] or "tok_c_fwd" in k[2]]
if af:
    print("attn_fwd / tok_c_fwd count", len(af), "first start", (af[0][0] - t0) / 1e3, "last end", (af[-1][1] - t0) / 1e3)
print("dct_c_fwd last end", last("dct_c_fwd"), "tok_a_fwd first/last", first("tok_a_fwd"), last("tok_a_fwd"))
print("tc_conv_ws (start,end)", [(round((k[0] - t0) / 1e3, 3), round((k[1] - t0) / 1e3, 3)) for k in ks if "tc_conv_ws" in k[2]])
print("upsample2_fwd ends", [round((k[1] - t0) / 1e3, 3) for k in ks if "upsample2_fwd" in k[2]])
print("maxpool_fwd starts", [round((k[0] - t0) / 1e3, 3) for k in ks if "maxpool_fwd" in k[2]])
print("first 3 tc_conv_fwd (start,end)", [(round((k[0] - t0) / 1e3, 3), round((k[1] - t0) / 1e3, 3)) for k in ks if "tc_conv_fwd" in k[2]][:8])
print("loss_reduce first start", first("loss_reduce"))
print("attn_bwd first/last", first("attn_bwd"), last("attn_bwd"))
print("last tc_conv_wgrad end", last("tc_conv_wgrad"))
print("adam start", first("FusedOptimizer"))
# per-kernel average in-step duration for the token kernels
import re as _re
def _short(n):
    n = n.replace("(anonymous namespace)::", "").replace("void ", "")
    m = _re.match(r"([\w:]+)", n)
    return (m.group(1) if m else n)[:48]
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n, _ in ks:
    key = _short(n)
    agg[key][0] += 1; agg[key][1] += (e - s) / 1e3
for k, (n, tt) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{tt:8.2f} ms {n:5d} {1e3*tt/n:8.1f} us  {k}")

# ---- the token-kernel chain of the forward pass: (start ms, duration us, gap to the previous kernel of the same stream us, name)
import re
def short(n):
    m = re.search(r"(\w+)(<|\()", n.replace("(anonymous namespace)::", "").replace("void ", ""))
    return m.group(1) if m else n[:30]
evs = [(e.time_range.start, e.time_range.end, short(e.name), getattr(e, "stream", None)) for e in ev]
streams = collections.defaultdict(list)
for s, e, n, st in sorted(evs):
    streams[st].append((s, e, n))
print("streams:", {k: len(v) for k, v in streams.items()})
for st, lst in streams.items():
    toks = [x for x in lst if any(t in x[2] for t in ("attn", "dct_", "tok_", "gemm_tile", "layernorm", "reduce_partials", "patch"))]
    if len(toks) < 50: continue
    print(f"--- stream {st}: first 40 kernels of the token chain")
    prev = None
    for s, e, n in lst[:int(os.environ.get("HDF_TL_FIRST", "40"))]:
        print(f"  {(s - t0) / 1e3:8.3f} ms  {(e - s):7.1f} us  gap {(s - prev) if prev else 0:7.1f} us  {n}")
        prev = e
    break

# ---- a window of the backward pass (all streams): what runs next to the transformer-backward chain
for win in os.environ.get("HDF_TL_WINDOWS", "24.0:24.9").split(","):
    lo, hi = (float(v) for v in win.split(":"))
    print(f"--- kernels starting in [{lo}, {hi}] ms")
    for s, e, n, _ in sorted(evs):
        if lo <= (s - t0) / 1e3 <= hi:
            print(f"  {(s - t0) / 1e3:8.3f} ms  {(e - s):7.1f} us  {n}")

# ---- token-branch kernels that "took" more than 100 us (waiting for SM resources next to a persistent convolution kernel)
tok = ("attn", "dct_", "gemm_tile", "layernorm", "reduce_partials", "patch", "ln_param", "act_dropout", "colsum", "posemb")
slow = [(s, e, n) for s, e, n, _ in sorted(evs) if any(t in n for t in tok) and e - s > 100]
print(f"--- token kernels longer than 100 us: {len(slow)}, total {sum(e - s for s, e, _ in slow) / 1e3:.2f} ms")
for s, e, n in slow:
    co = [(bs, be, bn) for bs, be, bn, _ in evs if "tc_conv" in bn and bs < e and be > s]
    print(f"  {(s - t0) / 1e3:8.3f} ms  {(e - s):7.1f} us  {n:24s} overlapping: " +
          ", ".join(f"{bn}[{(bs - t0) / 1e3:.2f}-{(be - t0) / 1e3:.2f}]" for bs, be, bn in co))
