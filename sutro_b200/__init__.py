"""sutro_b200 — B200-native local backend for the sutro.infer() hot path."""
__version__ = "0.1.0"
