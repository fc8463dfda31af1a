"""Stub so that the reference's modules/barcode_trimmer.py imports (edlib is absent; primer
trimming is out of scope). TEST INFRASTRUCTURE ONLY."""


def align(*a, **k):
    raise NotImplementedError("edlib is not available in this image")
