"""Diagnostic (not a pytest test): fused-upsample row kernel on a multi-tile-per-CTA problem."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'vocal-remover_b200'))
from lib import _native  # noqa: E402

N, Cl, h, w, Cs, Cout, act = [int(v) for v in (sys.argv[1:8] if len(sys.argv) > 7 else (2, 32, 256, 128, 16, 16, 1))]
g = torch.Generator().manual_seed(7)
low = torch.randn(N, Cl, h, w, generator=g)
skip = torch.randn(N, Cs, 2 * h, 2 * w, generator=g)
wgt = torch.randn(Cout, Cl + Cs, 3, 3, generator=g) / ((Cl + Cs) * 9) ** 0.5
b = torch.randn(Cout, generator=g) * 0.1
ctx = _native.Context(0, 2048, 1024, 32, 128, 256, 1, 0)
outs = []
for fused in (0, 1):
    y = torch.empty((N, Cout, 2 * h, 2 * w), dtype=torch.float32, device='cuda')
    dl, ds, dw, db = low.cuda(), skip.cuda(), wgt.cuda(), b.cuda()
    ctx.check(ctx.lib.vr_debug_decoder(ctx.handle, _native.ptr(dl), N, Cl, h, w, _native.ptr(ds), Cs, _native.ptr(dw),
                                       _native.ptr(db), Cout, act, fused, _native.ptr(y), _native.stream_ptr()),
              'vr_debug_decoder')
    torch.cuda.synchronize()
    outs.append(y.cpu())
    print('fused', fused, 'done', flush=True)
print('max diff fused vs staged', (outs[0] - outs[1]).abs().max().item())
