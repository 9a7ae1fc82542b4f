"""Drop-in for the reference's ``utils/geometry_helpers.py`` (its second copy of
``euler2mat``, imported at losses.py:6 and never called)."""
import torch

from .stn import _axis_rotation


def euler2mat(angle):
    """utils/geometry_helpers.py:5-40: R = Rx.bmm(Ry).bmm(Rz) from [B,3] euler angles."""
    zeros = angle[:, 2].detach() * 0
    ones = zeros.detach() + 1
    rx, ry, rz = (_axis_rotation(i, torch.cos(angle[:, i]), torch.sin(angle[:, i]), zeros, ones) for i in range(3))
    return rx.bmm(ry).bmm(rz)
