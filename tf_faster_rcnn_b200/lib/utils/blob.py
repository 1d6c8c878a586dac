"""Image list -> zero-padded NHWC float32 blob (lib/utils/blob.py:17-30)."""
import numpy as np


def im_list_to_blob(ims):
    """Stack HxWx3 images top-left aligned into [N, Hmax, Wmax, 3] float32, zero padded."""
    hmax = max(im.shape[0] for im in ims)
    wmax = max(im.shape[1] for im in ims)
    blob = np.zeros((len(ims), hmax, wmax, 3), dtype=np.float32)
    for i, im in enumerate(ims):
        blob[i, :im.shape[0], :im.shape[1], :] = im
    return blob
