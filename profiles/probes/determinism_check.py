"""Run the same forward + loss + backward N times and compare every output and gradient BITWISE with the first run:
any stream race or uninitialised read shows up here long before it trips a tolerance in the parity tests."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import hdf_oracle as O
from hdenseformer_b200.loss import CEPlusDice, DeepSuperloss
from hdenseformer_b200.models import HDenseFormer

def run(in_ch, n_cls, nf, size, td, batch, bf16, reps):
    shapes = O.param_shapes(in_ch, n_cls, nf, size, td)
    sd = O.synth_state_dict(shapes, seed=7)
    m = HDenseFormer(in_ch, n_cls, nf, image_size=size, transformer_depth=td)
    m.load_state_dict(sd); m = m.cuda().eval()
    x = (O.synth_petct(batch, size, seed=3) if in_ch == 2 else O.synth_mr(batch, in_ch, size, seed=3)).cuda()
    t = O.synth_label(batch, n_cls, size, seed=3).cuda()
    crit = DeepSuperloss(CEPlusDice(ignore_index=0))
    ref = None; bad = {}
    for r in range(reps):
        m.zero_grad(set_to_none=True)
        if bf16:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                outs = m(x)
        else:
            outs = m(x)
        loss = crit(outs, t); loss.backward()
        torch.cuda.synchronize()
        cur = {f"out{i}": o.detach().clone() for i, o in enumerate(outs)}
        cur["loss"] = loss.detach().clone()
        cur.update({"grad:" + k: p.grad.detach().clone() for k, p in m.named_parameters()})
        if ref is None: ref = cur; continue
        for k, v in cur.items():
            if not torch.equal(v, ref[k]):
                d = (v.float() - ref[k].float()).abs().max().item()
                bad.setdefault(k, []).append((r, d))
    tag = f"in{in_ch} nf{nf} {size} td{td} B{batch} {'bf16' if bf16 else 'fp32'}"
    if not bad: print(f"[deterministic] {tag}: {reps} identical runs")
    else:
        print(f"[NONDETERMINISTIC] {tag}: {len(bad)} tensors differ")
        for k, v in sorted(bad.items(), key=lambda kv: -max(d for _, d in kv[1]))[:12]:
            print(f"   {k}: runs {[r for r, _ in v][:8]} max|diff| {max(d for _, d in v):.3e} (|ref|max {ref[k].float().abs().max().item():.3e})")

reps = int(os.environ.get("REPS", 25))
run(2, 2, 16, (32, 32, 32), 4, 2, False, reps)
run(3, 2, 8, (16, 32, 48), 8, 2, False, reps)
run(2, 2, 16, (64, 64, 64), 4, 1, True, reps)
run(2, 2, 32, (96, 96, 96), 12, 2, True, max(6, reps // 4))
