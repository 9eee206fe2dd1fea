"""Sequence-bias model, expected-cleavage prediction and dispersion model (B200 path)."""
__all__ = ["bias", "dispersion", "predict"]
