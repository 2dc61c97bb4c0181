"""`torchinfo.summary` as the reference calls it (pipeline_utils.py:384): a parameter count line instead of the table."""
import logging


def summary(model, input_size=None, **kwargs):
    n = sum(p.numel() for p in model.parameters())
    logging.info(f"{type(model).__name__}: {n:,} parameters (compat torchinfo stand-in; input {input_size})")
    return None
