"""Drop-in shadow of the reference's ``model/ViBERTgrid_net.py``.

Put this directory AHEAD of the reference checkout on PYTHONPATH (and set PYTHONSAFEPATH=1 so the
script directory does not win):

    PYTHONSAFEPATH=1 PYTHONPATH=/root/repo/dropin:/root/repo:/path/to/ViBERTgrid-PyTorch \
        torchrun --nnodes 1 --nproc_per_node 8 /path/to/ViBERTgrid-PyTorch/eval_SROIE.py -c cfg.yaml

``model`` is a namespace package in the reference (no __init__.py), so only this one module is shadowed:
``model.crf``, ``pipeline.*`` and ``data.*`` still resolve to the reference.  Do NOT add model/__init__.py here.
"""
from vibertgrid_pytorch_b200.net import ViBERTgridNet  # noqa: F401

print(f"[vibertgrid_b200] drop-in ViBERTgridNet active ({__file__})")
