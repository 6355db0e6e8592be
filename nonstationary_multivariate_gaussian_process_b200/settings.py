"""Constants of the reference's Utility/settings.py:3-6 (must match bit for bit)."""
import torch

jitter = 1e-6
torchType = torch.DoubleTensor
precision = 1e-6
