"""Window-sharded multi-GPU execution of the hot path (one process per GPU).

Windows of a track are independent given the global normaliser (inference.py:74; BN is eval-mode, the LSTM
state is per window), so rank r owns a contiguous block of window indices and, with it, the
corresponding frame span of everything else on the path:

* default (``VR_GATHER=sharded``, needs hop == n_fft/2): every stage is sharded.  A rank computes the
  STFT of just the frames its windows read, max|X| over them (one 4-byte all-reduce gives the global normaliser),
  its masks, receives ONE halo mask frame (8 KB) from its right neighbour, and runs the masked inverse STFT +
  overlap-add of its own output span.  With ``tta`` (inference.py:83-98) the second, half-window-shifted pass is
  sharded by the same frame spans: a rank runs the count+1 shifted windows that overlap its span and averages them
  into its own mask frames, so the combine needs no exchange either; the STFT is then computed whole on every rank
  (0.3 ms per 4 minutes) because the TTA normaliser is numpy's lexicographic complex max of the whole track.  In the device-resident form the overlap-add kernel stores that span
  straight into rank 0's stem buffers, which are mapped into every rank through CUDA IPC (compute + gather in one
  kernel over NVLink); in the host form every rank moves only its own slice over PCIe.
* ``VR_GATHER=p2p``: only the net is sharded; the mask epilogue kernel stores into rank 0's mask buffer over
  NVLink, rank 0 runs the inverse STFT of the whole track.
* ``VR_GATHER=nccl``: as p2p but with one ``torch.distributed.gather`` of dense mask blocks (library baseline).

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


# Per-context caches, grow-only (a new track length re-uses the buffers when it fits, otherwise replaces them):
#   ('ipc', ctx, tag) -> (device pointer, owner flag, capacity in bytes)   buffers in rank 0's HBM mapped on every rank
#   ('ws', ctx) / ('hostws', ctx) -> (capacity, tensors...)                per-rank workspaces
# Keys hold the context object itself (not id(ctx), which can be recycled after a context is closed); release(ctx)
# drops everything a context owns (IPC mappings are closed with vr_shared_close).
_shared = {}
_calls = {}    # ctx -> number of sharded device-resident calls (selects one of two stem buffer sets, see below)


def release(ctx):
    """Free the cached workspaces and close the IPC mappings of ``ctx`` (call before closing the context)."""
    for key in [k for k in _shared if k[1] is ctx]:
        val = _shared.pop(key)
        if key[0] == 'ipc':
            ctx.lib.vr_shared_close(ctx.handle, val[0], 1 if val[1] else 0)
    _calls.pop(ctx, None)


class SharedHostBuffer(object):
    """One page-locked host buffer shared by all ranks of the node: a POSIX shared-memory segment that every rank maps and
    registers with CUDA (``cudaHostRegister``), so that the device-to-host copy of each rank's span lands in the SAME
    buffer and the host-buffer form of the path (``separate_wave_host``) returns assembled stems on every rank - the host
    counterpart of the peer-mapped stem buffers of the device-resident form.

    Collective: build it with ``SharedHostBuffer.create`` on every rank.  ``create`` returns None on EVERY rank when any
    rank fails a step (no /dev/shm, registration refused, ...); callers then fall back to per-rank pinned buffers."""

    def __init__(self):
        self.shm = None
        self.tensor = None
        self.registered = False
        self.owner = False

    @staticmethod
    def _agree(ok, world, group):
        import torch.distributed as dist
        if world == 1:
            return bool(ok)
        on_gpu = dist.get_backend(group) == 'nccl'
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device='cuda' if on_gpu else 'cpu')
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        return bool(t.item())

    @classmethod
    def create(cls, shape, world, rank, group=None, register=None):
        """float32 buffer of ``shape``; register=None registers with CUDA when a device is available."""
        import numpy as np
        from multiprocessing import shared_memory
        self = cls()
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = max(4, 4 * n)
        names = [None]
        ok = True
        try:
            if rank == 0:
                # tmpfs pages are allocated on first touch and a full /dev/shm then raises SIGBUS: make sure the whole
                # segment fits and reserve it up front (ENOSPC here instead of a bus error later)
                st = os.statvfs('/dev/shm')
                if st.f_bavail * st.f_frsize < nbytes + (64 << 20):
                    raise OSError('not enough room in /dev/shm for %d bytes' % nbytes)
                self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
                self.owner = True
                os.posix_fallocate(self.shm._fd, 0, nbytes)
                names[0] = self.shm.name
        except Exception:
            ok = False
            names[0] = None
            if self.shm is not None:   # created but not reservable
                try:
                    self.shm.unlink()
                    self.shm.close()
                except Exception:
                    pass
                self.shm = None
                self.owner = False
        if world > 1:
            import torch.distributed as dist
            dist.broadcast_object_list(names, src=0, group=group)
        ok = ok and names[0] is not None
        try:
            if ok and rank != 0:
                self.shm = shared_memory.SharedMemory(name=names[0])
                try:   # the creator unlinks the segment; keep this process's resource tracker from doing it as well
                    from multiprocessing import resource_tracker
                    resource_tracker.unregister(self.shm._name, 'shared_memory')
                except Exception:
                    pass
            if ok:
                arr = np.ndarray((n,), dtype=np.float32, buffer=self.shm.buf)
                self.tensor = torch.from_numpy(arr).view(*[int(d) for d in shape])
                if register is None:
                    register = torch.cuda.is_available()
                if register:
                    rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), nbytes, 0)
                    self.registered = int(rc) == 0
                    ok = self.registered and self.tensor.is_pinned()
        except Exception:
            ok = False
        if not cls._agree(ok, world, group):
            self.close(world, group, collective=False)
            return None
        return self

    def close(self, world=1, group=None, collective=True):
        """Unregister and unmap; rank 0 removes the segment after every rank has let go of it."""
        if self.registered:
            try:
                torch.cuda.synchronize()
                torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            except Exception:
                pass
            self.registered = False
        self.tensor = None
        if collective and world > 1:
            import torch.distributed as dist
            dist.barrier(group=group)
        if self.shm is not None:
            if self.owner:
                try:
                    self.shm.unlink()
                except Exception:
                    pass
            try:
                self.shm.close()   # raises BufferError while a caller still holds a view; the mapping then goes with it
            except Exception:
                pass
            self.shm = None


class _RawCudaArray(object):
    """Zero-copy torch view of a raw device pointer (``torch.as_tensor`` reads __cuda_array_interface__)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': '<f4', 'data': (int(ptr), False),
                                         'version': 2, 'strides': None}


def _shared_buffer(ctx, tag, nbytes, world, rank, dev, group):
    """A buffer in rank 0's HBM, mapped on every rank (cached per context, tag and size)."""
    import torch.distributed as dist
    key = ('ipc', ctx, tag)
    if key in _shared:
        if _shared[key][2] >= int(nbytes):
            return _shared[key][0]
        old = _shared.pop(key)   # same decision on every rank: sizes derive from the track length only
        ctx.check(ctx.lib.vr_shared_close(ctx.handle, old[0], 1 if old[1] else 0), 'vr_shared_close')
    handle = ctypes.create_string_buffer(64)
    ptr = _native.c_vp()
    if rank == 0:
        ctx.check(ctx.lib.vr_shared_alloc(ctx.handle, nbytes, ctypes.byref(ptr), handle), 'vr_shared_alloc')
    t = torch.tensor(list(handle.raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0, group=group)
    if rank != 0:
        raw = bytes(t.cpu().tolist())
        ctx.check(ctx.lib.vr_shared_open(ctx.handle, raw, ctypes.byref(ptr)), 'vr_shared_open')
    _shared[key] = (ptr, rank == 0, int(nbytes))
    return ptr


def separate_wave(sp, d_wave, tta=False, world=1, rank=0, group=None):
    """CUDA wave (2, L) on every rank -> (inst, voc) CUDA waves on rank 0 (None elsewhere)."""
    if world == 1:
        return sp.separate_wave(d_wave, tta=tta)
    mode = os.environ.get('VR_GATHER', 'sharded')
    n_windows, _ = window_count(1 + d_wave.shape[1] // sp.model.hop_length, sp.cropsize, sp.offset)
    if mode == 'sharded' and (sp.model.hop_length * 2 != sp.model.n_fft or n_windows < world):
        mode = 'p2p'
    if tta and mode != 'sharded':
        # the TTA combine is a read-modify-write of the mask: it stays local to the rank that owns the frames
        raise NotImplementedError('multi-GPU --tta needs the sharded mode (VR_GATHER=sharded, hop == n_fft/2, '
                                  'at least one window per rank)')
    if mode == 'sharded':
        return _separate_wave_sharded(sp, d_wave, world, rank, group, to_rank0=True, tta=tta)
    if mode == 'p2p':
        return _separate_wave_p2p(sp, d_wave, world, rank, group)
    return _separate_wave_nccl(sp, d_wave, world, rank, group)


def shard_plan(n_frames, cropsize, offset, world, rank):
    """Frame spans of one rank: (first, count, roi, f0, f1, a, b, k0, k1).

    [f0, f1) mask frames it produces, [a, b) spectrogram frames its windows read (f1 < b whenever f1 < n_frames,
    so the halo frame is included), [k0, k1) output hops it reconstructs (samples [hop*k0, hop*k1))."""
    n_windows, roi = window_count(n_frames, cropsize, offset)
    first, count, _ = shard_windows(n_windows, world, rank)
    f0 = min(n_frames, first * roi)
    f1 = min(n_frames, (first + count) * roi)
    a = max(0, min(n_frames, first * roi - offset))
    b = min(n_frames, (first + count) * roi + offset) if count > 0 else a
    k0 = min(f0, n_frames - 1)
    k1 = min(f1, n_frames - 1)
    return first, count, roi, f0, f1, a, b, k0, k1


def tta_window_range(first, count, n_windows):
    """Windows [g0, g0+c) of the shifted second TTA pass (inference.py:91-96) that overlap the mask frames of first-pass
    windows [first, first+count): shifted window g covers track frames [roi*g - roi/2, roi*g + roi/2), and the pass
    has n_windows + 1 windows."""
    if count <= 0:
        return first, 0
    return first, min(count + 1, n_windows + 1 - first)


def _separate_wave_sharded(sp, d_wave, world, rank, group, to_rank0, local_out=None, tta=False):
    """Everything sharded (see module docstring).  ``to_rank0``: stems are assembled in rank 0's HBM by the
    overlap-add kernels (returns them on rank 0); otherwise each rank writes its span into ``local_out``."""
    import torch.distributed as dist
    ctx = sp._ctx()
    dev = d_wave.device
    hop, n_fft = sp.model.hop_length, sp.model.n_fft
    bins = n_fft // 2 + 1
    L = d_wave.shape[1]
    T = 1 + L // hop
    Lo = hop * (T - 1)
    first, count, roi, f0, f1, a, b, k0, k1 = shard_plan(T, sp.cropsize, sp.offset, world, rank)
    with torch.cuda.device(dev):
        st = _native.stream_ptr()
        key = ('ws', ctx)
        if key not in _shared or _shared[key][0] < T:
            _shared.pop(key, None)
            _shared[key] = (T, torch.empty((2 * bins * T,), dtype=torch.complex64, device=dev),
                            torch.empty((2 * bins * T,), dtype=torch.float32, device=dev))
        spec = _shared[key][1][:2 * bins * T].view(2, bins, T)
        mask = _shared[key][2][:2 * bins * T].view(2, bins, T)
        norm = torch.zeros(1, dtype=torch.float32, device=dev)
        if tta:
            # inference.py:87,94: the normaliser is |lexicographic complex max| of the whole (padded) track
            ctx.check(ctx.lib.vr_stft(ctx.handle, _native.ptr(d_wave), L, _native.ptr(spec), T, None, st), 'vr_stft')
            ctx.check(ctx.lib.vr_normaliser(ctx.handle, _native.ptr(spec), T, 1, _native.ptr(norm), st),
                      'vr_normaliser')
        else:
            if b > a:
                ctx.check(ctx.lib.vr_stft_range(ctx.handle, _native.ptr(d_wave), L, _native.ptr(spec), T, a, b, st),
                          'vr_stft_range')
                ctx.check(ctx.lib.vr_normaliser_range(ctx.handle, _native.ptr(spec), T, a, b, _native.ptr(norm), st),
                          'vr_normaliser_range')
            dist.all_reduce(norm, op=dist.ReduceOp.MAX, group=group)        # inference.py:74, 4 bytes
        if count > 0:
            ctx.check(ctx.lib.vr_separate_windows(ctx.handle, _native.ptr(spec), T, _native.ptr(norm), sp.offset,
                                                  first, count, _native.ptr(mask), T, 0, 0, st),
                      'vr_separate_windows')
            if tta:
                # second pass, padded by a further roi/2 on the left: mask frame j of its concatenation is track
                # frame j - roi/2, averaged into what the first pass left there ((old + new) / 2, inference.py:98).
                # Its edge windows also touch roi/2 frames of the neighbours' spans in the LOCAL mask; those frames
                # are never read here (the halo frame below is overwritten by the neighbour's value).
                n_windows, _ = window_count(T, sp.cropsize, sp.offset)
                g0, c2 = tta_window_range(first, count, n_windows)
                ctx.check(ctx.lib.vr_separate_windows(ctx.handle, _native.ptr(spec), T, _native.ptr(norm),
                                                      sp.offset + roi // 2, g0, c2, _native.ptr(mask), T, roi // 2, 1,
                                                      st), 'vr_separate_windows (tta)')
        # halo: output hop k needs frames k and k+1, so the last hop of this span needs the first mask frame of
        # the right neighbour
        ops = []
        send_col = recv_col = None
        if rank > 0 and count > 0 and f0 < T:
            send_col = mask[:, :, f0].contiguous()
            ops.append(dist.P2POp(dist.isend, send_col, rank - 1, group=group))
        if count > 0 and f1 < T:
            recv_col = torch.empty((2, bins), dtype=torch.float32, device=dev)
            ops.append(dist.P2POp(dist.irecv, recv_col, rank + 1, group=group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if recv_col is not None:
            mask[:, :, f1].copy_(recv_col)
        if to_rank0:
            # Two stem buffer sets, used alternately: the stems a call returns on rank 0 stay valid until the call after
            # next, whose remote stores are ordered (by the completion all-reduce of the call in between, which rank 0
            # enqueues after whatever it launched on the returned tensors) behind rank 0's reads - no entry barrier.
            which = _calls.get(ctx, 0) & 1
            _calls[ctx] = _calls.get(ctx, 0) + 1
            inst_ptr = _shared_buffer(ctx, 'inst%d' % which, 2 * Lo * 4, world, rank, dev, group)
            voc_ptr = _shared_buffer(ctx, 'voc%d' % which, 2 * Lo * 4, world, rank, dev, group)
        else:
            inst_ptr, voc_ptr = _native.ptr(local_out[0]), _native.ptr(local_out[1])
        if k1 > k0:
            ctx.check(ctx.lib.vr_apply_mask_istft_range(ctx.handle, _native.ptr(spec), _native.ptr(mask), T, k0, k1,
                                                        inst_ptr, voc_ptr, st), 'vr_apply_mask_istft_range')
        if not to_rank0:
            return hop * k0, hop * k1
        done = torch.zeros(1, dtype=torch.float32, device=dev)
        dist.all_reduce(done, group=group)     # stream-ordered: every rank's remote stores precede it
        if rank != 0:
            return None, None
        inst = torch.as_tensor(_RawCudaArray(inst_ptr.value, (2, Lo)), device=dev)
        voc = torch.as_tensor(_RawCudaArray(voc_ptr.value, (2, Lo)), device=dev)
        return inst, voc


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
        mask_ptr = _shared_buffer(ctx, 'mask', 2 * bins * T * 4, world, rank, dev, group)
        spec = torch.empty((2, bins, T), dtype=torch.complex64, device=dev)
        norm = torch.empty(1, dtype=torch.float32, device=dev)
        ctx.check(ctx.lib.vr_stft(ctx.handle, _native.ptr(d_wave), L, _native.ptr(spec), T, _native.ptr(norm), st),
                  'vr_stft')
        if count > 0:
            # the epilogue kernel writes frames [first*roi, (first+count)*roi) of rank 0's mask directly
            ctx.check(ctx.lib.vr_separate_windows(ctx.handle, _native.ptr(spec), T, _native.ptr(norm), sp.offset,
                                                  first, count, mask_ptr, T, 0, 0, st), 'vr_separate_windows')
        flag = torch.zeros(1, dtype=torch.float32, device=dev)
        dist.all_reduce(flag, group=group)           # stream-ordered: every shard has landed in rank 0's HBM
        inst = voc = None
        if rank == 0:
            Lo = hop * (T - 1)
            inst = torch.empty((2, Lo), dtype=torch.float32, device=dev)
            voc = torch.empty((2, Lo), dtype=torch.float32, device=dev)
            ctx.check(ctx.lib.vr_apply_mask_istft(ctx.handle, _native.ptr(spec), mask_ptr, T, _native.ptr(inst),
                                                  _native.ptr(voc), st), 'vr_apply_mask_istft')
        dist.all_reduce(flag, group=group)           # the mask may be overwritten by the next call from here on
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
    """Host (pinned) wave -> host (pinned) stems; H2D and D2H copies are part of the call.

    world == 1: the whole track.  world > 1 (sharded mode): every rank copies in only the samples its frames read
    and copies out only its own slice ``[:, s0:s1]`` of the stems into ITS ``h_inst`` / ``h_voc`` (returned as
    (s0, s1)); the slices of all ranks tile the track.  Other modes leave the full stems on rank 0.
    """
    dev = torch.device('cuda', sp._ctx().device_index)
    if world == 1:
        ctx = sp._ctx()
        with torch.cuda.device(dev):
            ctx.check(ctx.lib.vr_separate_wave_host(ctx.handle, h_wave.data_ptr(), h_wave.shape[1], 1 if tta else 0,
                                                    h_inst.data_ptr(), h_voc.data_ptr(), _native.stream_ptr()),
                      'vr_separate_wave_host')
        return 0, h_inst.shape[1]
    hop, n_fft = sp.model.hop_length, sp.model.n_fft
    L = h_wave.shape[1]
    T = 1 + L // hop
    n_windows, _ = window_count(T, sp.cropsize, sp.offset)
    mode = os.environ.get('VR_GATHER', 'sharded')
    if mode != 'sharded' or tta or hop * 2 != n_fft or n_windows < world:
        d_wave = h_wave.to(dev, non_blocking=True)
        inst, voc = separate_wave(sp, d_wave, tta=tta, world=world, rank=rank, group=group)
        if rank == 0:
            h_inst.copy_(inst, non_blocking=True)
            h_voc.copy_(voc, non_blocking=True)
        torch.cuda.synchronize(dev)
        return (0, h_inst.shape[1]) if rank == 0 else (0, 0)
    _, _, _, _, _, a, b, _, _ = shard_plan(T, sp.cropsize, sp.offset, world, rank)
    with torch.cuda.device(dev):
        key = ('hostws', sp._ctx())
        Lo = hop * (T - 1)
        if key not in _shared or _shared[key][0] < L:
            _shared.pop(key, None)
            _shared[key] = (L, torch.empty((2 * L,), dtype=torch.float32, device=dev),
                            torch.empty((2 * L,), dtype=torch.float32, device=dev),
                            torch.empty((2 * L,), dtype=torch.float32, device=dev))
        d_wave = _shared[key][1][:2 * L].view(2, L)
        d_inst = _shared[key][2][:2 * Lo].view(2, Lo)
        d_voc = _shared[key][3][:2 * Lo].view(2, Lo)
        if b > a:   # samples read by frames [a, b)
            w0 = max(0, a * hop - n_fft // 2)
            w1 = min(L, (b - 1) * hop + n_fft // 2)
            for c in range(2):   # row slices are contiguous: direct DMA from pinned memory, no host staging copy
                d_wave[c, w0:w1].copy_(h_wave[c, w0:w1], non_blocking=True)
        s0, s1 = _separate_wave_sharded(sp, d_wave, world, rank, group, to_rank0=False, local_out=(d_inst, d_voc))
        if s1 > s0:
            for c in range(2):
                h_inst[c, s0:s1].copy_(d_inst[c, s0:s1], non_blocking=True)
                h_voc[c, s0:s1].copy_(d_voc[c, s0:s1], non_blocking=True)
        torch.cuda.synchronize(dev)
        return s0, s1
