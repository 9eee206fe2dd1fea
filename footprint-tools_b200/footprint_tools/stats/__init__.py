"""Statistics of the scoring path: window combination, posterior, FDR helpers (B200 path)."""
__all__ = ["distributions", "windowing", "posterior", "utils", "fdr"]
