"""Per-layer precision budget of the split-bf16 tensor-core convolution (CPU emulation; not a pytest module).

The product computes every convolution as hi*hi + lo*hi + hi*lo of bf16 pairs (x = hi + lo, 16-bit significand) with fp32
accumulation, and stores every activation as such a pair.  This script emulates that arithmetic inside the oracle's
functional CascadedNet (oracle/net_oracle.py) and measures, for ONE layer at a time, how much mask error is added when
that layer alone drops one of the two correction products:

    3pass   hi*hi + lo*hi + hi*lo          (the product path)
    no_wlo  (hi + lo) * w_hi               two MMAs per k-step: weights rounded to bf16
    no_xlo  x_hi * (w_hi + w_lo)           two MMAs per k-step: activations rounded to bf16
    1pass   x_hi * w_hi

Output: one line per layer with the mask max-abs error against the fp32 oracle on the first window of the 10 s input
(seeded synthetic checkpoint, lib/synth.py).  Usage: python tests/precision_budget.py [out.tsv]
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'vocal-remover_b200'))
from lib import synth  # noqa: E402
from oracle import net_oracle, separator_oracle, stft_oracle  # noqa: E402


def bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def split(x):
    hi = bf16(x)
    return hi, bf16(x - hi)


SCHEME = {}     # layer prefix -> scheme (default 3pass)
LAYERS = []     # filled on the first pass, in execution order


def conv_bn_act_emulated(sd, p, x, stride=1, pad=1, dil=1, act='relu'):
    if p not in LAYERS:
        LAYERS.append(p)
    w = net_oracle._t(sd, p + '.conv.0.weight').double()
    g, b = net_oracle._t(sd, p + '.conv.1.weight').double(), net_oracle._t(sd, p + '.conv.1.bias').double()
    m, v = net_oracle._t(sd, p + '.conv.1.running_mean').double(), net_oracle._t(sd, p + '.conv.1.running_var').double()
    scale = g / torch.sqrt(v + net_oracle.BN_EPS)
    wf = (w * scale[:, None, None, None]).float()            # BN folded at load time (engine.cu make_conv)
    bias = (b - m * scale).float()
    scheme = SCHEME.get(p, '3pass')
    xh, xl = split(x)                                        # activations are stored as hi + lo
    wh, wl = split(wf)
    kw = dict(stride=stride, padding=pad, dilation=dil)
    if scheme == '3pass':
        y = F.conv2d(xh, wh, None, **kw) + F.conv2d(xl, wh, None, **kw) + F.conv2d(xh, wl, None, **kw)
    elif scheme == 'no_wlo':
        y = F.conv2d(xh, wh, None, **kw) + F.conv2d(xl, wh, None, **kw)
    elif scheme == 'no_xlo':
        y = F.conv2d(xh, wh, None, **kw) + F.conv2d(xh, wl, None, **kw)
    else:
        y = F.conv2d(xh, wh, None, **kw)
    y = y + bias[None, :, None, None]
    y = F.relu(y) if act == 'relu' else F.leaky_relu(y, 0.01)
    h, l = split(y)
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
    net_oracle.conv_bn_act = conv_bn_act_emulated
    lines = []

    def run(tag):
        t0 = time.time()
        err = (net_oracle.forward(sd, x) - ref).abs().max().item()
        lines.append('%s\t%.3e' % (tag, err))
        print(lines[-1], '(%.1f s)' % (time.time() - t0), flush=True)
        return err

    base = run('all layers 3pass')
    for scheme in ('no_wlo', 'no_xlo'):
        for p in list(LAYERS):
            SCHEME.clear()
            SCHEME[p] = scheme
            run('%s\t%s' % (scheme, p))
    for scheme in ('no_wlo', 'no_xlo', '1pass'):
        SCHEME.clear()
        for p in LAYERS:
            SCHEME[p] = scheme
        run('all layers %s' % scheme)
    net_oracle.conv_bn_act = exact
    if out_path:
        with open(out_path, 'w') as f:
            f.write('# mask max-abs error vs the fp32 oracle, first window of the 10 s input; baseline (all 3pass) %.3e\n' % base)
            f.write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    main()
