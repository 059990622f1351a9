"""Inference-side piece of the reference's lib/dataset.py: the window geometry used by Separator."""


def make_padding(width, cropsize, offset):
    """Same contract as the reference's lib/dataset.py:198-205: returns (left, right, roi_size).

    The padded width left + width + right is a multiple of roi_size plus 2*offset, so the padded
    spectrogram splits into whole cropsize-wide windows at stride roi_size.
    """
    roi_size = cropsize - 2 * offset
    if roi_size == 0:
        roi_size = cropsize
    return offset, roi_size - (width % roi_size) + offset, roi_size
