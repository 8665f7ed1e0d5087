"""Per-parameter gradient parity report (GPU vs CPU oracle) for one training step; prints the worst tensors."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.conftest import collate_scans  # noqa: E402


def main():
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200.spvcnn import Net3DSeg
    from oracle import ft_glue as og, ts_ops as ts
    batch = collate_scans("nuscenes", 2)
    coords, feats = batch["coords"], batch["feats"]
    n = coords.shape[0]
    g = torch.Generator().manual_seed(3)
    img = torch.randn(n, 96, generator=g)
    labels = torch.randint(0, 20, (n,), generator=g)
    for mode, omode in (("f32", None), ("tc", None), ("tc", "bf16")):
        os.environ["FT3D_CONV"] = mode
        ts.OPERAND_DTYPE = omode
        torch.manual_seed(1)
        o = og.Net3DSeg(fusion="middle").train()
        o.dropout.p = 0.0
        lo = torch.nn.functional.cross_entropy(o(ts.SparseTensor(feats, coords), img)["lidar_seg_logit"], labels)
        lo.backward()
        po = dict(o.named_parameters())
        gmax = max(p.grad.norm().item() for p in o.parameters() if p.grad is not None)
        m = Net3DSeg(fusion="middle")
        m.load_state_dict(o.state_dict())
        m = m.cuda().train()
        m.dropout.p = 0.0
        taps = {}
        out = m(ft.SparseTensor(feats.cuda(), coords.cuda()), img.cuda(), taps=taps)
        lg = torch.nn.functional.cross_entropy(out["lidar_seg_logit"], labels.cuda())
        lg.backward()
        rows = []
        for name, p in m.named_parameters():
            go = po[name].grad.double()
            d = (p.grad.double().cpu() - go).norm().item()
            rows.append((d / max(go.norm().item(), 1e-4 * gmax), d / max(go.norm().item(), 1e-30), go.norm().item(), name))
        rows.sort(reverse=True)
        print("gpu mode %s vs oracle arithmetic %s:  loss gpu %.6f oracle %.6f  gmax %.3e" % (mode, omode or "fp32", lg.item(), lo.item(), gmax))
        for r in rows[:12]:
            print("   err %.3e  rel %.3e  |g| %.3e  %s" % r)
        med = sorted(r[0] for r in rows)[len(rows) // 2]
        print("   median err %.3e over %d tensors" % (med, len(rows)))


if __name__ == "__main__":
    main()
