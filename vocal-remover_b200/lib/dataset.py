"""Inference-side pieces of the reference's lib/dataset.py: the window geometry used by Separator and the file pairing
used by pseudo.py (training-set construction is out of scope)."""
import os


def make_padding(width, cropsize, offset):
    """Same contract as the reference's lib/dataset.py:198-205: returns (left, right, roi_size).

    The padded width left + width + right is a multiple of roi_size plus 2*offset, so the padded
    spectrogram splits into whole cropsize-wide windows at stride roi_size.
    """
    roi_size = cropsize - 2 * offset
    if roi_size == 0:
        roi_size = cropsize
    return offset, roi_size - (width % roi_size) + offset, roi_size


INPUT_EXTS = ['.wav', '.m4a', '.mp3', '.mp4', '.flac']


def make_pair(mix_dir, inst_dir):
    """lib/dataset.py:144-160: sorted mixture / instrument files of two directories, zipped pairwise."""
    X_list = sorted([os.path.join(mix_dir, fname) for fname in os.listdir(mix_dir)
                     if os.path.splitext(fname)[1] in INPUT_EXTS])
    y_list = sorted([os.path.join(inst_dir, fname) for fname in os.listdir(inst_dir)
                     if os.path.splitext(fname)[1] in INPUT_EXTS])
    return list(zip(X_list, y_list))


def shard_files(filelist, world, rank):
    """File-level sharding of a many-file job (pseudo.py:41-74) over ``world`` processes (one per GPU): round-robin, so
    that every rank gets a similar mix of track lengths; whole files, no exchange between ranks."""
    return [f for i, f in enumerate(filelist) if i % world == rank]
