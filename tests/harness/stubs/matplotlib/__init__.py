"""Harness-only stand-in for matplotlib (absent from this image; imported at module level by the reference's
utils/ViBERTgrid_visualize.py:5, which pipeline/train_val_utils.py:21 imports).  Nothing on the tested path plots."""
