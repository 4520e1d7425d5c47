"""speak_hack_b200 — B200-native (sm_100a) implementation of the IRFD hot path of johndpope/SPEAK-hack.

Public surface mirrors the reference modules (model.py / styleganv1.py); the math runs in libirfd_b200.so.
"""
from .model import IRFD, IRFDLoss, StyleGANLoss, mse_loss  # noqa: F401
from .generator import FC, ApplyNoise, ApplyStyle, StyleGenerator, SynthesisBlock, SynthesisNetwork  # noqa: F401
from .encoder import ResNet50Encoder  # noqa: F401
from .discriminator import StyleDiscriminator  # noqa: F401
