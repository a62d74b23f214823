"""Test-harness stand-in for the ``ltp`` package (Chinese word segmentation), which the reference's deployment modules import
at module level (deployment/module_load.py:7, inference_preporcessing.py:8) and use only in the ``chn_ltp`` OCR parse mode."""


class LTP:
    def __init__(self, *a, **k):
        raise RuntimeError("ltp stand-in: the chn_ltp parse mode is not exercised by the harness")
