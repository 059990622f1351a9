"""Window-sharded multi-GPU execution of the hot path (one process per GPU).

Windows of a track are independent given the global normaliser (inference.py:74; BN is eval-mode, the
LSTM state is per window), so rank r runs a contiguous block of window indices and the masks meet on
rank 0 before the overlap-add.  Every rank holds the wave, computes the (cheap) STFT and
the normaliser itself, so the only exchange is the masks:

* default (``VR_GATHER=p2p``): rank 0's whole-track mask buffer is mapped into every rank through CUDA IPC
  and the mask epilogue kernel (sigmoid + crop of the last 1x1 conv) stores its shard straight into rank 0's
  HBM over NVLink - compute and "gather" are one kernel; two tiny NCCL barriers order producers / consumer.
* ``VR_GATHER=nccl``: one ``torch.distributed.gather`` of dense per-rank blocks (the plain-library baseline).

``world == 1`` degenerates to the fused single-GPU call.
"""
import ctypes
import os
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


_shared = {}   # (ctx id, bytes) -> (device pointer, owner flag)


def _shared_mask(ctx, nbytes, world, rank, dev, group):
    """Rank 0's mask buffer, mapped on every rank (cached per context and size)."""
    import torch.distributed as dist
    key = (id(ctx), int(nbytes))
    if key in _shared:
        return _shared[key][0]
    handle = ctypes.create_string_buffer(64)
    ptr = _native.c_vp()
    if rank == 0:
        ctx.check(ctx.lib.vr_shared_alloc(ctx.handle, nbytes, ctypes.byref(ptr), handle), 'vr_shared_alloc')
    t = torch.tensor(list(handle.raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0, group=group)
    if rank != 0:
        raw = bytes(t.cpu().tolist())
        ctx.check(ctx.lib.vr_shared_open(ctx.handle, raw, ctypes.byref(ptr)), 'vr_shared_open')
    _shared[key] = (ptr, rank == 0)
    return ptr


def separate_wave(sp, d_wave, tta=False, world=1, rank=0, group=None):
    """CUDA wave (2, L) on every rank -> (inst, voc) CUDA waves on rank 0 (None elsewhere)."""
    if tta and world > 1:
        raise NotImplementedError('multi-GPU --tta: shard files instead (SURVEY 8(f) rank 3)')
    if world == 1:
        return sp.separate_wave(d_wave, tta=tta)
    if os.environ.get('VR_GATHER', 'p2p') == 'p2p':
        return _separate_wave_p2p(sp, d_wave, world, rank, group)
    return _separate_wave_nccl(sp, d_wave, world, rank, group)


def _separate_wave_p2p(sp, d_wave, world, rank, group):
    import torch.distributed as dist
    ctx = sp._ctx()
    dev = d_wave.device
    hop, n_fft = sp.model.hop_length, sp.model.n_fft
    bins = n_fft // 2 + 1
    L = d_wave.shape[1]
    T = 1 + L // hop
    n_windows, roi = window_count(T, sp.cropsize, sp.offset)
    first, count, per = shard_windows(n_windows, world, rank)
    st = _native.stream_ptr()
    with torch.cuda.device(dev):
        mask_ptr = _shared_mask(ctx, 2 * bins * T * 4, world, rank, dev, group)
        spec = torch.empty((2, bins, T), dtype=torch.complex64, device=dev)
        norm = torch.empty(1, dtype=torch.float32, device=dev)
        ctx.check(ctx.lib.vr_stft(ctx.handle, _native.ptr(d_wave), L, _native.ptr(spec), T, _native.ptr(norm), st),
                  'vr_stft')
        if count > 0:
            # the epilogue kernel writes frames [first*roi, (first+count)*roi) of rank 0's mask directly
            ctx.check(ctx.lib.vr_separate_windows(ctx.handle, _native.ptr(spec), T, _native.ptr(norm), sp.offset,
                                                  first, count, mask_ptr, T, 0, 0, st), 'vr_separate_windows')
        torch.cuda.current_stream().synchronize()   # remote stores are complete when the kernels are
        dist.barrier(group=group)                    # every shard has landed in rank 0's HBM
        inst = voc = None
        if rank == 0:
            Lo = hop * (T - 1)
            inst = torch.empty((2, Lo), dtype=torch.float32, device=dev)
            voc = torch.empty((2, Lo), dtype=torch.float32, device=dev)
            ctx.check(ctx.lib.vr_apply_mask_istft(ctx.handle, _native.ptr(spec), mask_ptr, T, _native.ptr(inst),
                                                  _native.ptr(voc), st), 'vr_apply_mask_istft')
            torch.cuda.current_stream().synchronize()
        dist.barrier(group=group)                    # the mask may be overwritten by the next call from here on
        return inst, voc


def _separate_wave_nccl(sp, d_wave, world, rank, group):
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
