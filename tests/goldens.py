"""Loader for the committed golden fixtures (tests/golden/*.npz, generated from
the unmodified reference by tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))

TRAIN_CFG = {"l1_weight": 0.15, "l_ssim_weight": 0.85, "l_smooth_weight": 0.05, "num_scales": 1,
             "l_depth_consist_weight": 0.14, "min_depth": 0.06, "max_depth": 2.67, "l_smooth": False,
             "l_reconstruction": True, "l_inverse": True, "l_depth_consist": False,
             "with_auto_mask": True, "l_ssim": True, "with_depth_mask": False}
FULL_CFG = dict(TRAIN_CFG, l_depth_consist=True, with_depth_mask=True)
NOAUTO_CFG = dict(FULL_CFG, with_auto_mask=False)
SMOOTH_CFG = dict(TRAIN_CFG, l_smooth=True)
PAIR_CFGS = {"train": TRAIN_CFG, "full": FULL_CFG, "noauto": NOAUTO_CFG}
LOSS_CFGS = {"train": TRAIN_CFG, "full": FULL_CFG, "smooth": SMOOTH_CFG}
PFT_OPTIONS = {"num_source_imgs": 2, "diff_img_argmin": True, "automasking": True,
               "l_inverse_reconstruction": True, "l_depth_consist": True, "l_depth_consist_weight": 0.15,
               "l_depth_init": True, "l_depth_init_weight": 0.1, "l_smooth": False, "l_smooth_weight": 0.05,
               "l_pose_consist": False, "plotting": False, "epochs": 20}


class Golden:
    def __init__(self, name, device="cpu"):
        self.name = name
        self.device = device
        self._z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))

    def __contains__(self, key):
        return key in self._z.files

    def np(self, key):
        return self._z[key]

    def t(self, key):
        a = self._z[key]
        if a.dtype == np.uint8:
            a = a.astype(np.float32)
        return torch.from_numpy(a).to(self.device)

    def frames(self):
        """Inputs in the layout of tcsfm_b200.synth.make_frames."""
        disps = [self.t("in/disp%d" % j) for j in range(3)]
        lo, hi = 0.06, 2.67
        depths = [1 / (1 / hi + (1 / lo - 1 / hi) * d) for d in disps]
        return {"target": self.t("in/target"), "sources": [self.t("in/source%d" % j) for j in range(2)],
                "disps": disps, "depths": depths,
                "poses": [self.t("in/pose%d" % j) for j in range(2)],
                "poses_inv": [self.t("in/pose_inv%d" % j) for j in range(2)], "K": self.t("in/K")}


def rel_l2(a, b):
    a = a.detach().double().flatten().cpu()
    b = b.detach().double().flatten().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / den if den > 0 else (a - b).norm().item()
