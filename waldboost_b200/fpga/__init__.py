"""Integer channel features of the reference's FPGA variant (reference waldboost/fpga/__init__.py, channels.py).
Only the channel functions are in scope here; the bank-restricted tree training of waldboost.fpga is not."""
from .channels import grad_hist_4_u1, grad_mag_u1

__all__ = ["grad_hist_4_u1", "grad_mag_u1"]
