"""Process-tagged logger (reference: gllm/utils/__init__.py:35-41, worker.py:65-73)."""
import logging
import os
import sys

logger = logging.getLogger("gllm_b200")
if not logger.handlers:
    _h = logging.StreamHandler(sys.stderr)
    _h.setFormatter(logging.Formatter("%(asctime)s %(levelname)s %(message)s", "%H:%M:%S"))
    logger.addHandler(_h)
    logger.setLevel(os.environ.get("GLLM_B200_LOG", "INFO"))
    logger.propagate = False


def set_prefix(prefix: str):
    for h in logger.handlers:
        h.setFormatter(logging.Formatter(f"%(asctime)s %(levelname)s [{prefix}] %(message)s", "%H:%M:%S"))
