"""Data format at the host boundary of the path.

The reference's loaders convert uint8 frames to float tensors on the host
(utils/custom_transforms.py:74: ``torch.from_numpy(im).float() / 255``) and ship fp32 to the GPU
(data/kitti_loader.py:60-98).  ``images_from_uint8`` does the same conversion on the device, bit for
bit (the same IEEE division), so the frames can cross the host link as bytes -- a quarter of the
traffic that bounds the end-to-end rate of the loss step."""
from . import _raw, ops


def images_from_uint8(u8, out=None):
    """[..., H, W] uint8 CUDA tensor -> fp32 in [0, 1], identical to ``u8.float() / 255`` evaluated on the host."""
    ops._require_cuda(u8, out)
    with ops._guard(u8):
        return _raw.u8_to_float(ops.lib(), u8, out)
