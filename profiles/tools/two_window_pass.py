"""One 2-window pass of the default net through the product path (used under compute-sanitizer / ncu)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'vocal-remover_b200'))
import inference  # noqa: E402
from lib import nets, synth  # noqa: E402

dev = torch.device('cuda:0')
model = nets.CascadedNet(2048, 1024, 32, 128)
model.load_state_dict(synth.to_torch_state_dict(synth.make_state_dict()))
model.to(dev)
sp = inference.Separator(model, dev, int(os.environ.get('VR_BATCH', '2')), 256, False)
wave = synth.sine_mix(float(os.environ.get('VR_SECONDS', '4.0')))
inst, voc = sp.separate_wave(wave, tta=bool(int(os.environ.get('VR_TTA', '0'))))
torch.cuda.synchronize()
print('ok', inst.shape, float(np.abs(inst).max()), float(np.abs(inst + voc - wave[:, :inst.shape[1]]).max()))
