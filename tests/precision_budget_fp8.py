"""Would fp8 correction products pass the gate?  (CPU emulation; not a pytest module; follow-up of precision_budget.py)

The 0.33 bound of the split-bf16 convolution comes from issuing three bf16 MMAs per product.  The two correction products
(lo*hi and hi*lo) are ~2^-9 of the main one, so they need only a few significant bits: with both operands of the
corrections rounded to fp8 (e4m3, per-tensor power-of-two scale) they could run as kind::f8f6f4 MMAs at twice the bf16
rate - 2 units of tensor time per product instead of 3.  This script measures the mask error of that scheme with the
same emulation as precision_budget.py:

    3pass     hi*hi + lo*hi + hi*lo                       all bf16 (the product path)
    fp8corr   hi*hi + q8(lo)*q8(w_hi) + q8(hi)*q8(w_lo)   corrections in e4m3

for every layer at once and for one layer at a time.  Usage: python tests/precision_budget_fp8.py [out.tsv]
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import precision_budget as pb  # noqa: E402
from precision_budget import net_oracle, separator_oracle, stft_oracle, synth  # noqa: E402


def q8(x):
    """e4m3 rounding with a per-tensor power-of-two scale that puts max|x| just under 256 (of 448)."""
    m = x.abs().max().item()
    if m == 0.0:
        return x
    s = 2.0 ** np.floor(np.log2(256.0 / m))
    return (x * s).to(torch.float8_e4m3fn).to(torch.float32) / s


def conv_fp8corr(sd, p, x, stride=1, pad=1, dil=1, act='relu'):
    if pb.SCHEME.get(p, '3pass') != 'fp8corr':
        return pb.conv_bn_act_emulated(sd, p, x, stride, pad, dil, act)
    if p not in pb.LAYERS:
        pb.LAYERS.append(p)
    w = net_oracle._t(sd, p + '.conv.0.weight').double()
    g, b = net_oracle._t(sd, p + '.conv.1.weight').double(), net_oracle._t(sd, p + '.conv.1.bias').double()
    m, v = net_oracle._t(sd, p + '.conv.1.running_mean').double(), net_oracle._t(sd, p + '.conv.1.running_var').double()
    scale = g / torch.sqrt(v + net_oracle.BN_EPS)
    wf = (w * scale[:, None, None, None]).float()
    bias = (b - m * scale).float()
    xh, xl = pb.split(x)
    wh, wl = pb.split(wf)
    kw = dict(stride=stride, padding=pad, dilation=dil)
    y = F.conv2d(xh, wh, None, **kw) + F.conv2d(q8(xl), q8(wh), None, **kw) + F.conv2d(q8(xh), q8(wl), None, **kw)
    y = y + bias[None, :, None, None]
    y = F.relu(y) if act == 'relu' else F.leaky_relu(y, 0.01)
    h, l = pb.split(y)
    return h + l


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    sd = synth.to_torch_state_dict(synth.make_state_dict())
    wave = synth.sine_mix(10.0)
    X = stft_oracle.wave_to_spectrogram(wave, 1024, 2048)
    pad_l, pad_r, roi = separator_oracle.make_padding(X.shape[2], 256, 64)
    Xp = np.pad(X, ((0, 0), (0, 0), (pad_l, pad_r)))
    Xp /= np.abs(X).max()
    x = torch.from_numpy(np.abs(Xp[None, :, :, 128:384]).astype(np.float32))
    ref = net_oracle.forward(sd, x)
    exact = net_oracle.conv_bn_act
    net_oracle.conv_bn_act = conv_fp8corr
    lines = []

    def run(tag):
        t0 = time.time()
        err = (net_oracle.forward(sd, x) - ref).abs().max().item()
        lines.append('%s\t%.3e' % (tag, err))
        print(lines[-1], '(%.1f s)' % (time.time() - t0), flush=True)
        return err

    base = run('all layers 3pass')
    pb.SCHEME.clear()
    for p in pb.LAYERS:
        pb.SCHEME[p] = 'fp8corr'
    run('all layers fp8corr')
    for p in list(pb.LAYERS):
        pb.SCHEME.clear()
        pb.SCHEME[p] = 'fp8corr'
        run('fp8corr\t%s' % p)
    net_oracle.conv_bn_act = exact
    if out_path:
        with open(out_path, 'w') as f:
            f.write('# mask max-abs error vs the fp32 oracle, first window of the 10 s input; baseline (all 3pass) %.3e\n' % base)
            f.write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    main()
