"""Audio file I/O either side of the hot path (out of scope for the B200 kernels).

The reference uses ``librosa.load(..., res_type='kaiser_fast')`` and ``soundfile.write``
(inference.py:136-138,173,178).  They are used when installed; otherwise a stdlib ``wave`` reader /
writer handles PCM WAV at the requested sample rate (no resampling).
"""
import wave as _wave

import numpy as np


def load(path, sr, mono=False, dtype=np.float32):
    try:
        import librosa
        return librosa.load(path, sr=sr, mono=mono, dtype=dtype, res_type='kaiser_fast')
    except ImportError:
        pass
    with _wave.open(path, 'rb') as f:
        nch, width, rate, nframes = f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()
        raw = f.readframes(nframes)
    if rate != sr:
        raise RuntimeError('input is %d Hz but --sr is %d and librosa/resampy are not installed' % (rate, sr))
    if width == 2:
        x = np.frombuffer(raw, dtype='<i2').astype(np.float32) / 32768.0
    elif width == 4:
        x = np.frombuffer(raw, dtype='<i4').astype(np.float32) / 2147483648.0
    elif width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise RuntimeError('unsupported WAV sample width %d' % width)
    x = x.reshape(-1, nch).T
    if mono or nch == 1:
        x = x.mean(axis=0)
    return np.ascontiguousarray(x.astype(dtype)), rate


def write(path, data, sr):
    """data: (L, channels) float array, like soundfile.write."""
    try:
        import soundfile as sf
        sf.write(path, data, sr)
        return
    except ImportError:
        pass
    data = np.asarray(data)
    if data.ndim == 1:
        data = data[:, None]
    pcm = np.clip(np.round(data * 32767.0), -32768, 32767).astype('<i2')
    with _wave.open(path, 'wb') as f:
        f.setnchannels(pcm.shape[1])
        f.setsampwidth(2)
        f.setframerate(sr)
        f.writeframes(pcm.tobytes())
