"""Diagnostic (not a pytest test): row-streaming tcgen05 kernel with (mode 1) and without (mode 0, default) the UMMA
matrix-base-offset field in its pixel-shifted descriptors.  Result on B200: mode 0 exact, mode 1 garbage."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'vocal-remover_b200'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_gpu_parity import _ref_conv, _run_debug_conv  # noqa: E402
from lib import _native  # noqa: E402

CASES = [
    (1, 64, 8, 128, 32, 3, 1, (1, 1), 1),
    (1, 16, 8, 128, 16, 3, 1, (1, 1), 1),
    (1, 32, 16, 256, 32, 3, 1, (1, 1), 2),
    (2, 97, 8, 128, 32, 3, 1, (1, 1), 1),
    (1, 192, 16, 128, 64, 3, 1, (1, 1), 1),
    (1, 2, 8, 256, 8, 3, 1, (1, 1), 0),
]
lib = _native.load_library()
for mode in (0, 1):
    lib.vr_debug_set(0, mode)
    for case in CASES:
        N, Cin, H, W, Cout, k, stride, dil, act = case
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(N, Cin, H, W, generator=g)
        w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
        b = torch.randn(Cout, generator=g) * 0.1
        ctx = _native.Context(0, 2048, 1024, 32, 128, 256, 1, 0)
        try:
            y = _run_debug_conv(ctx, x, w, b, k, stride, dil, act, 1)
            ref = _ref_conv(x, w, b, k, stride, dil, act)
            err = (y - ref).abs().max().item()
            # where is the error: interior vs. border columns
            e = (y - ref).abs()
            print('mode', mode, case, 'maxerr %.3e' % err, 'ref max %.2f' % ref.abs().max().item(),
                  'col-err(first 4 cols) %s' % [round(v, 4) for v in e.amax(dim=(0, 1, 2))[:4].tolist()],
                  'row-err(first 3 rows) %s' % [round(v, 4) for v in e.amax(dim=(0, 1, 3))[:3].tolist()], flush=True)
        except Exception as ex:
            print('mode', mode, case, 'EXC', str(ex)[:200], flush=True)
lib.vr_debug_set(0, 0)
