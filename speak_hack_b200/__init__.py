"""speak_hack_b200 — B200-native (sm_100a) implementation of the IRFD hot path of johndpope/SPEAK-hack.

Public surface mirrors the reference modules (model.py / styleganv1.py); the math runs in libirfd_b200.so.
"""
__all__ = ["_lib", "ops"]
