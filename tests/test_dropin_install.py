"""dropin.install() makes the reference's own import statements resolve to the fused modules."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r'''
import sys
sys.path.insert(0, %r)
import tcsfm_b200.dropin as dropin
dropin.install()
from models.stn import *            # the statement at reference losses.py:7 / train_mono.py:4
from losses import SSIM_Loss, get_smooth_loss, Compute_Loss
from utils.geometry_helpers import euler2mat
import tcsfm_b200.stn as s, tcsfm_b200.losses as l
assert inverse_warp2 is s.inverse_warp2 and pose_vec2mat is s.pose_vec2mat and check_sizes is s.check_sizes
assert Compute_Loss is l.Compute_Loss and SSIM_Loss is l.SSIM_Loss
for name in ("pixel2cam", "cam2pixel", "cam2pixel2", "euler2mat", "quat2mat", "inverse_warp", "set_id_grid"):
    assert name in globals(), name
print("ok")
''' % ROOT


def test_install_in_fresh_interpreter():
    out = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr
