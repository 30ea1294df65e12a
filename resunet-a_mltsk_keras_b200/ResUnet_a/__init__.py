"""Drop-in for the reference's ``ResUnet_a`` package: ``from ResUnet_a.model2 import Resunet_a``."""
