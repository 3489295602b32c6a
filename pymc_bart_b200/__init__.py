"""pymc_bart_b200 — B200-native PGBART."""
__version__ = "0.1.0"
