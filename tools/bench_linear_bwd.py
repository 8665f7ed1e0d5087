"""Times formulations of the point-branch Linear backward (dW, db) on the bench shapes (TF32 allowed)."""
import torch
torch.backends.cuda.matmul.allow_tf32 = True
dev = "cuda"
N = 55312
def t(fn, it=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / it
for out, inp in [(256, 32), (128, 256), (96, 128), (256, 96), (20, 96)]:
    x = torch.randn(N, inp, device=dev); g = torch.randn(N, out, device=dev); w = torch.randn(out, inp, device=dev)
    ones = torch.ones(N, device=dev)
    xa = torch.cat([x, torch.ones(N, 1, device=dev)], 1)
    r = {}
    r["g.t@x"] = t(lambda: g.t() @ x)
    r["(x.t@g).t"] = t(lambda: (x.t() @ g))
    r["g.t@x bf16"] = t(lambda: g.t().bfloat16() @ x.bfloat16())
    r["g.sum0"] = t(lambda: g.sum(0))
    r["g.t@ones"] = t(lambda: torch.mv(g.t(), ones))
    r["ones@g"] = t(lambda: ones.unsqueeze(0) @ g)
    r["dx=g@w"] = t(lambda: g @ w)
    r["fwd x@w.t"] = t(lambda: x @ w.t())
    torch.backends.cuda.matmul.allow_tf32 = False
    r["g.t@x fp32"] = t(lambda: g.t() @ x)
    r["(x.t@g) fp32"] = t(lambda: (x.t() @ g))
    torch.backends.cuda.matmul.allow_tf32 = True
    print("out %3d in %3d : " % (out, inp) + "  ".join("%s %.1f" % kv for kv in r.items()))
