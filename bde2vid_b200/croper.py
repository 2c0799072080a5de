"""Pad/crop geometry of the reference driver (utils_func/inference_utils.py:26-32, 69-114).

Only index arithmetic lives here; the zero padding itself is produced by the voxeliser kernel
(it writes the padded grid directly) and the crop is a view."""
from math import ceil, floor


def optimal_crop_size(max_size, max_subsample_factor):
    """Smallest multiple of 2^max_subsample_factor that is >= max_size."""
    s = 2 ** max_subsample_factor
    return s * ceil(max_size / s)


class Croper:
    def __init__(self, num_encoders):
        self.num_encoders = num_encoders
        self.width = self.height = None
        self.width_crop_size = self.height_crop_size = None

    def update_params(self, width, height):
        self.width, self.height = width, height
        self.width_crop_size = optimal_crop_size(width, self.num_encoders)
        self.height_crop_size = optimal_crop_size(height, self.num_encoders)
        self.padding_top = ceil(0.5 * (self.height_crop_size - height))
        self.padding_bottom = floor(0.5 * (self.height_crop_size - height))
        self.padding_left = ceil(0.5 * (self.width_crop_size - width))
        self.padding_right = floor(0.5 * (self.width_crop_size - width))
        self.cx = floor(self.width_crop_size / 2)
        self.cy = floor(self.height_crop_size / 2)
        self.ix0 = self.cx - floor(width / 2)
        self.ix1 = self.cx + ceil(width / 2)
        self.iy0 = self.cy - floor(height / 2)
        self.iy1 = self.cy + ceil(height / 2)

    def pad(self, x):
        """Zero-pad a [..., H, W] tensor (view-level plumbing for callers that hold dense voxel grids)."""
        import torch.nn.functional as F
        h, w = x.shape[-2:]
        if h != self.height_crop_size or w != self.width_crop_size:
            if h != self.height or w != self.width:
                self.update_params(w, h)
            x = F.pad(x, (self.padding_left, self.padding_right, self.padding_top, self.padding_bottom))
        return x

    def crop(self, img):
        return img[..., self.iy0:self.iy1, self.ix0:self.ix1] if self.num_encoders != -1 else img
