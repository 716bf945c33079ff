"""B200-native CFM Euler solve + DAC-VAE decode hot path."""
__version__ = "0.1.0"
