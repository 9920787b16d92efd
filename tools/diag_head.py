"""GPU diagnostic (not a test): where do GPU and oracle head gradients differ?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_classifier_b200 import _cabi as cabi
from oracle import head_oracle as ho

B, D, C = 32, 768, 20
g = torch.Generator().manual_seed(9)
p = ho.init_head(D, C)
pg = {k: v.clone().cuda().contiguous() for k, v in p.items()}
X = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=1)
y = torch.randint(0, C, (B,), generator=g)
masks = tuple(((torch.rand(B, n, generator=g) >= 0.1).float() / 0.9) for n in (D, D // 2))
for use_masks in (False, True):
    mk = masks if use_masks else None
    loss_ref, grads, _ = ho.head_grads(X, y, p, mk, "ce")
    # GPU gradient through a train step with lr = 0 is not observable; use a tiny SGD-free trick: run ac_head_grad (no masks)
    if not use_masks:
        gout = {k: torch.zeros_like(v) for k, v in pg.items()}
        cabi.head_grad(X.cuda(), y.cuda(), pg, grad_out=gout)
        for k in ho.PARAM_ORDER:
            d = (gout[k].cpu() - grads[k]).abs()
            rel = d.max() / grads[k].abs().max()
            bad = (d > 1e-3 * grads[k].abs().max()).sum().item()
            print(f"no-mask grad {k}: max abs diff {d.max():.3e} (rel to max {rel:.2e}), elements off by >1e-3*max: {bad}/{d.numel()}")
    else:
        # with masks: one Adam step; recover sign(g) from the update
        m = {k: torch.zeros_like(v) for k, v in pg.items()}
        v = {k: torch.zeros_like(t) for k, t in pg.items()}
        p0 = {k: t.clone() for k, t in pg.items()}
        cabi.head_train_step(X.cuda(), y.cuda(), pg, m, v, step=1, masks=(masks[0].cuda(), masks[1].cuda()), weight_decay=0.0)
        for k in ho.PARAM_ORDER:
            g_gpu = m[k].cpu() / 0.1                      # m = 0.1 * clipped grad at step 1
            coef = min(1.0, 1.0 / (torch.sqrt(sum((grads[n].double() ** 2).sum() for n in ho.PARAM_ORDER)).item() + 1e-6))
            g_ref = grads[k] * coef
            d = (g_gpu - g_ref).abs()
            bad = (d > 1e-3 * g_ref.abs().max())
            print(f"masked grad {k}: max abs diff {d.max():.3e} vs max|g| {g_ref.abs().max():.3e}; off elements {bad.sum().item()}/{d.numel()}")
            if bad.any() and g_ref.dim() == 2:
                rows = bad.any(1).nonzero().flatten().tolist()
                cols = bad.any(0).nonzero().flatten().tolist()
                print(f"   rows with mismatches: {len(rows)} (first {rows[:10]}), cols: {len(cols)} (first {cols[:10]})")
