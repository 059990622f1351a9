"""Window-sharded multi-GPU execution of the hot path (one process per GPU).

Windows of a track are independent given the global normaliser (inference.py:74; BN is eval-mode, the
LSTM state is per window), so rank r runs a contiguous block of window indices and the masks meet on
rank 0 before the overlap-add.  Every rank holds the wave, computes the (cheap) STFT and
the normaliser itself, so the only exchange is ONE gather of fp32 mask blocks over NCCL / NVLink.
``world == 1`` degenerates to the fused single-GPU call.
"""
import numpy as np
import torch

from . import _native


def shard_windows(n_windows, world, rank):
    """Contiguous block [first, first+count) of window indices for ``rank`` (ceil split, last ranks may be empty)."""
    per = -(-n_windows // world)
    first = min(n_windows, rank * per)
    count = max(0, min(n_windows, first + per) - first)
    return first, count, per


def window_count(n_frames, cropsize, offset):
    """Number of windows Separator._separate cuts (inference.py:44) after make_padding (lib/dataset.py:198-205)."""
    roi = cropsize - 2 * offset
    if roi == 0:
        roi = cropsize
    pad_r = roi - (n_frames % roi) + offset
    return (offset + n_frames + pad_r - 2 * offset) // roi, roi


def mask_block_frames(first, count, per, roi):
    """Frame range [lo, hi) of the track covered by a rank's dense (2, bins, per*roi) block."""
    return first * roi, (first + count) * roi


def assemble_mask(blocks, n_frames):
    """Rank-ordered list of (2, bins, per*roi) blocks -> (2, bins, n_frames) mask (inference.py:66,77)."""
    return torch.cat(list(blocks), dim=2)[:, :, :n_frames].contiguous()


def gather_blocks(block, world, rank, group=None):
    """The single exchange step of the sharded path: every rank's mask block -> rank 0."""
    import torch.distributed as dist
    gathered = [torch.empty_like(block) for _ in range(world)] if rank == 0 else None
    dist.gather(block, gathered, dst=0, group=group)
    return gathered


def separate_wave(sp, d_wave, tta=False, world=1, rank=0, group=None):
    """CUDA wave (2, L) on every rank -> (inst, voc) CUDA waves on rank 0 (None elsewhere)."""
    if tta and world > 1:
        raise NotImplementedError('multi-GPU --tta: shard files instead (SURVEY 8(f) rank 3)')
    if world == 1:
        return sp.separate_wave(d_wave, tta=tta)
    ctx = sp._ctx()
    dev = d_wave.device
    model = sp.model
    hop, n_fft = model.hop_length, model.n_fft
    bins = n_fft // 2 + 1
    L = d_wave.shape[1]
    T = 1 + L // hop
    n_windows, roi = window_count(T, sp.cropsize, sp.offset)
    first, count, per = shard_windows(n_windows, world, rank)
    st = _native.stream_ptr()
    with torch.cuda.device(dev):
        spec = torch.empty((2, bins, T), dtype=torch.complex64, device=dev)
        norm = torch.empty(1, dtype=torch.float32, device=dev)
        ctx.check(ctx.lib.vr_stft(ctx.handle, _native.ptr(d_wave), L, _native.ptr(spec), T, _native.ptr(norm), st),
                  'vr_stft')
        # this rank's masks as a dense (2, bins, per*roi) block: frame j of the track lands at j - first*roi
        block = torch.zeros((2, bins, per * roi), dtype=torch.float32, device=dev)
        if count > 0:
            ctx.check(ctx.lib.vr_separate_windows(ctx.handle, _native.ptr(spec), T, _native.ptr(norm), sp.offset,
                                                  first, count, _native.ptr(block), per * roi, first * roi, 0, st),
                      'vr_separate_windows')
        gathered = gather_blocks(block, world, rank, group)
        if rank != 0:
            return None, None
        mask = assemble_mask(gathered, T)
        Lo = hop * (T - 1)
        inst = torch.empty((2, Lo), dtype=torch.float32, device=dev)
        voc = torch.empty((2, Lo), dtype=torch.float32, device=dev)
        ctx.check(ctx.lib.vr_apply_mask_istft(ctx.handle, _native.ptr(spec), _native.ptr(mask), T, _native.ptr(inst),
                                              _native.ptr(voc), st), 'vr_apply_mask_istft')
        return inst, voc


def separate_wave_host(sp, h_wave, h_inst, h_voc, tta=False, world=1, rank=0, group=None):
    """Host (pinned) wave -> host (pinned) stems on rank 0; H2D and D2H copies are part of the call."""
    dev = torch.device('cuda', sp._ctx().device_index)
    if world == 1:
        ctx = sp._ctx()
        with torch.cuda.device(dev):
            ctx.check(ctx.lib.vr_separate_wave_host(ctx.handle, h_wave.data_ptr(), h_wave.shape[1], 1 if tta else 0,
                                                    h_inst.data_ptr(), h_voc.data_ptr(), _native.stream_ptr()),
                      'vr_separate_wave_host')
        return
    d_wave = h_wave.to(dev, non_blocking=True)
    inst, voc = separate_wave(sp, d_wave, tta=tta, world=world, rank=rank, group=group)
    if rank == 0:
        h_inst.copy_(inst, non_blocking=True)
        h_voc.copy_(voc, non_blocking=True)
    torch.cuda.synchronize(dev)
