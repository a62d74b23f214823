"""Harness-only stand-in for the `seqeval` package (absent from this image and its wheelhouse; imported at module level by
the reference's pipeline/criteria.py:2-7).  Micro-averaged token-level scores over the non-"O" tags -- enough for the
reference's train/validate scripts to run end to end; NOT a product component."""
