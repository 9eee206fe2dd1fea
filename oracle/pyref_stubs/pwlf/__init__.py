"""Empty stand-in for the un-vendored `pwlf` package (TEST INFRASTRUCTURE ONLY).
Only needed so `footprint_tools.modeling.dispersion` imports; the dispersion-model
*fit* (dispersion.pyx:357-469) is therefore not runnable in the oracle."""


class PiecewiseLinFit(object):
    def __init__(self, *a, **k):
        raise NotImplementedError("pwlf is not available in this container")
