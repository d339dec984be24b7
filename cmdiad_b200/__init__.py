"""cmdiad_b200 -- B200-native (sm_100a) memory-bank anomaly scoring path of evenrose/CMDIAD.

Layout: csrc/ (CUDA kernels + the C ABI of include/cmdiad_b200.h), _lib.py (ctypes binding), bank.py (device bank
handle), methods.py (host-side mirror of the reference's method classes), synth.py (synthetic patch features).
The CUDA library is mandatory: nothing here computes on the CPU.
"""
from . import synth  # noqa: F401  (pure numpy, importable without the library)


def __getattr__(name):
    # lazy so that `import cmdiad_b200` works before the library is built; any use of the product path loads it
    if name in ("Bank", "Comm", "ScoreResult", "BatchResult", "upsample_blur", "coreset_rownorms"):
        from . import bank
        return getattr(bank, name)
    if name in ("Features", "RGBFeatures", "DepthFeatures", "PointFeatures", "DoubleRGBPointFeatures",
                "RGBorXYZWithOneHallucination", "RGBorXYZWithOneHallucinationFromFeature", "METHODS", "default_args"):
        from . import methods
        return getattr(methods, name)
    raise AttributeError(name)
