"""CPU oracle for ``inference.Separator`` (inference.py:16-102).  TEST INFRASTRUCTURE.

numpy restatement of the sliding-window orchestration: make_padding (lib/dataset.py:198-205),
constant padding + global normaliser (inference.py:73-74, 86-87, 93-94), 256-frame windows at
stride roi (inference.py:42-68), mask truncation, TTA half-roi shifted pass (inference.py:89-98)
and the mask application of ``_postprocess`` (inference.py:32-36; ``--postprocess`` /
merge_artifacts is out of scope).  The per-window model is ``oracle.net_oracle``.
"""
import numpy as np
import torch

from . import net_oracle


def make_padding(width, cropsize, offset):
    """lib/dataset.py:198-205."""
    left = offset
    roi_size = cropsize - offset * 2
    if roi_size == 0:
        roi_size = cropsize
    right = roi_size - (width % roi_size) + left
    return left, right, roi_size


def _separate(sd, X_spec_pad, roi_size, n_fft, cropsize, offset, batchsize):
    """inference.py:42-68."""
    patches = (X_spec_pad.shape[2] - 2 * offset) // roi_size
    masks = []
    for i in range(0, patches, batchsize):
        batch = np.asarray([X_spec_pad[:, :, j * roi_size:j * roi_size + cropsize]
                            for j in range(i, min(patches, i + batchsize))])
        m = net_oracle.predict_mask(sd, torch.from_numpy(np.abs(batch)), n_fft, offset).numpy()
        masks.append(np.concatenate(list(m), axis=2))
    return np.concatenate(masks, axis=2)


def apply_mask(X_spec, mask):
    """inference.py:32-36 (postprocess=False)."""
    X_mag = np.abs(X_spec)
    X_phase = np.angle(X_spec)
    y_spec = mask * X_mag * np.exp(1.j * X_phase)
    v_spec = (1 - mask) * X_mag * np.exp(1.j * X_phase)
    return y_spec, v_spec


def separate_mask(sd, X_spec, n_fft=2048, cropsize=256, offset=64, batchsize=4):
    """Mask (2, bins, T) of inference.py:70-77."""
    n_frame = X_spec.shape[2]
    pad_l, pad_r, roi = make_padding(n_frame, cropsize, offset)
    X_pad = np.pad(X_spec, ((0, 0), (0, 0), (pad_l, pad_r)), mode='constant')
    X_pad /= np.abs(X_spec).max()
    mask = _separate(sd, X_pad, roi, n_fft, cropsize, offset, batchsize)
    return mask[:, :, :n_frame]


def separate_tta_mask(sd, X_spec, n_fft=2048, cropsize=256, offset=64, batchsize=4):
    """Mask of inference.py:83-98.  Note the normaliser is ``X_spec_pad.max()`` on a COMPLEX array
    (numpy lexicographic max: largest real part, ties by imaginary part), not max|X|."""
    n_frame = X_spec.shape[2]
    pad_l, pad_r, roi = make_padding(n_frame, cropsize, offset)
    X_pad = np.pad(X_spec, ((0, 0), (0, 0), (pad_l, pad_r)), mode='constant')
    X_pad /= X_pad.max()
    mask = _separate(sd, X_pad, roi, n_fft, cropsize, offset, batchsize)
    pad_l += roi // 2
    pad_r += roi // 2
    X_pad = np.pad(X_spec, ((0, 0), (0, 0), (pad_l, pad_r)), mode='constant')
    X_pad /= X_pad.max()
    mask_tta = _separate(sd, X_pad, roi, n_fft, cropsize, offset, batchsize)
    mask_tta = mask_tta[:, :, roi // 2:]
    return (mask[:, :, :n_frame] + mask_tta[:, :, :n_frame]) * 0.5


def separate(sd, X_spec, tta=False, **kw):
    mask = (separate_tta_mask if tta else separate_mask)(sd, X_spec, **kw)
    return apply_mask(X_spec, mask)
