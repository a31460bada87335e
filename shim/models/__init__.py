"""`models` package shim: put this directory in front of the reference tree on PYTHONPATH and the reference's own
main.py / engine.py / infer.py (`from models import build_model`, A2/main.py:13, A1/main.py:21) build the B200 path
instead of the ATen one, unchanged:

    PYTHONPATH=/root/repo/shim:/root/repo python main.py --spatial_prior learned --no_aux_loss --num_query_pattern 1 ...

Same contract as the reference's models/__init__.py:13-14: build_model(args) -> (model, criterion, postprocessors).
"""
from counting_detr_b200.models import build, build_model  # noqa: F401
