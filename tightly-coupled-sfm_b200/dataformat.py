"""Data format at the host boundary of the path.

The reference's loaders convert uint8 frames to float tensors on the host
(utils/custom_transforms.py:74: ``torch.from_numpy(im).float() / 255``) and ship fp32 to the GPU
(data/kitti_loader.py:60-98).  ``images_from_uint8`` does the same conversion on the device, bit for
bit (the same IEEE division), so the frames can cross the host link as bytes -- a quarter of the
traffic that bounds the end-to-end rate of the loss step."""
import torch

from . import _raw, ops


def images_from_uint8(u8, out=None):
    """[..., H, W] uint8 CUDA tensor -> fp32 in [0, 1], identical to ``u8.float() / 255`` evaluated on the host."""
    ops._require_cuda(u8, out)
    with ops._guard(u8):
        return _raw.u8_to_float(ops.lib(), u8, out)


def pack_slab(tensors, device=None, pin=False, align=64):
    """Collates a dict of equally typed tensors into ONE contiguous buffer and returns (slab, views): `views[name]`
    has the shape and values of `tensors[name]` and aliases the slab (every view starts on an `align`-element
    boundary: 256 bytes for fp32).  A minibatch packed this way on the host (pin=True: page-locked) and on the device
    with the same dict order crosses the host link as a single copy, `dev_slab.copy_(host_slab, non_blocking=True)`,
    instead of one copy per tensor -- the loader's eleven tensors per KITTI triplet minibatch otherwise pay the copy
    engine's set-up latency eleven times."""
    items = list(tensors.items())
    if not items:
        raise ValueError("pack_slab: no tensors")
    dtype = items[0][1].dtype
    if any(t.dtype != dtype for _, t in items):
        raise TypeError("pack_slab: the tensors of one slab must share a dtype")
    offsets, total = [], 0
    for _, t in items:
        offsets.append(total)
        total += -(-t.numel() // align) * align
    if pin:
        slab = torch.zeros(total, dtype=dtype).pin_memory()
    else:
        slab = torch.zeros(total, dtype=dtype, device=device if device is not None else items[0][1].device)
    views = {}
    for (name, t), off in zip(items, offsets):
        v = slab[off:off + t.numel()].view(t.shape)
        v.copy_(t)
        views[name] = v
    return slab, views
