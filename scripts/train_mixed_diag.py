"""Per-parameter gradient agreement of the mixed-precision training path with the fp32 oracle (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth, train_ref
from tests.helpers import build_product_model
DEV = "cuda:0"
seed, B, T, H = 44, 4, 6, int(os.environ.get("H", 256))
x = synth.make_input(seed, B, T); masks = train_ref.make_masks(seed, 2 * B); tgt = train_ref.make_targets(seed, 2 * B)
sd = synth.make_state_dict(seed, 1, H)
_, _, ref = train_ref.TrainOracle(sd, seed, 1, H).loss_and_grads(x, masks, tgt)
caps = {}
for prec in ("fp32", "bf16"):
    model, _ = build_product_model(seed, T, 1, H, prec, DEV)
    model.train()
    model.train_fp32_bptt = bool(os.environ.get("FP32_BPTT"))
    model._debug_capture = caps.setdefault(prec, {})
    out = model(torch.from_numpy(x).to(DEV), is_train=True, dropout_masks=torch.from_numpy(masks).to(DEV))[-1]
    train_ref.synthetic_loss(out, tgt).backward()
    print("==", prec)
    for name, p in model.named_parameters():
        if name not in ref: continue
        g = p.grad.detach().cpu().double().reshape(-1); r = ref[name].double().reshape(-1)
        if float(r.abs().max()) == 0: continue
        cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-30))
        rel = float((g - r).abs().max() / r.abs().max())
        rn = float((g - r).norm() / r.norm())
        if "gru" not in name: continue
        print(f"  {name:45s} cos {cos:.5f}  max-rel {rel:.3f}  norm-rel {rn:.3f}  |ref| {float(r.norm()):.3e}")

a, b = caps["fp32"], caps["bf16"]
Hh = H
def rel(x, y):
    return float((x - y).norm() / (x.norm() + 1e-30))
for k in a:
    if a[k].dim() == 2 and a[k].shape[1] == 3 * Hh and k in ("g_hcat", "h_cat"):
        for nm, lo, hi in (("F", 0, Hh), ("S", Hh, 2 * Hh), ("B", 2 * Hh, 3 * Hh)):
            print(f"{k}[{nm}] rel diff bf16 vs fp32: {rel(a[k][:, lo:hi], b[k][:, lo:hi]):.4f}")
    else:
        print(f"{k} rel diff: {rel(a[k], b[k]):.4f}")
for t in range(T):
    print(f"step {t}: seq_f {rel(a['seq_f'][t], b['seq_f'][t]):.4f} seq_b {rel(a['seq_b'][t], b['seq_b'][t]):.4f} "
          f"dgi_f {rel(a['dgi_f'][t*B:(t+1)*B], b['dgi_f'][t*B:(t+1)*B]):.4f} dgi_b {rel(a['dgi_b'][t*B:(t+1)*B], b['dgi_b'][t*B:(t+1)*B]):.4f}")
