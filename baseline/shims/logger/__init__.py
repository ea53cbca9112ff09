"""Stand-in for the PyPI `logger` package the reference imports (`from logger import logger`,
gllm/utils/__init__.py:27): a stdlib logger with one stream handler — the reference only re-formats the
handlers (`init_logger`) and calls info / warning / error on it.

Environment glue for this image (not part of the reference): the reference's `gllm/_C` is the vLLM `_C` extension;
in the vLLM build installed here (0.22) the norm / rope / activation ops it calls (`torch.ops._C.rms_norm`, ...)
live in a second library, `_C_stable_libtorch.abi3.so`, which the reference never loads. When
GLLM_REF_PRELOAD_LIBS lists such libraries they are loaded here, i.e. in every process that imports the reference.
"""
import logging
import os
import sys

logger = logging.getLogger("gllm")
if not logger.handlers:
    _h = logging.StreamHandler(sys.stderr)
    logger.addHandler(_h)
    logger.setLevel(os.environ.get("GLLM_REF_LOG", "INFO"))
    logger.propagate = False

if not hasattr(logger, "warning_once"):          # vLLM-style helper the reference's utils call on some platforms
    _seen = set()

    def _warning_once(msg, *args):
        if msg not in _seen:
            _seen.add(msg)
            logger.warning(msg, *args)

    logger.warning_once = _warning_once

if os.environ.get("GLLM_REF_ALIAS_VLLM") == "1":
    # The reference's native modules (gllm._C, gllm._moe_C, gllm.vllm_flash_attn) ARE vLLM's: its setup.py copies
    # them out of a vLLM wheel. Here they are taken from the vLLM installed in the image; aliasing the module
    # names makes sure every shared object is loaded exactly once per process (a second copy of the same library
    # would register its torch ops twice), whether or not the symlinks under baseline/_ref survived a file copy.
    import importlib
    for _ours, _theirs in (("gllm._C", "vllm._C"), ("gllm._moe_C", "vllm._moe_C"),
                           ("gllm.vllm_flash_attn", "vllm.vllm_flash_attn")):
        try:
            sys.modules.setdefault(_ours, importlib.import_module(_theirs))
        except Exception as _e:  # noqa: BLE001
            logger.warning("could not alias %s -> %s: %r", _ours, _theirs, _e)

_libs = [p for p in os.environ.get("GLLM_REF_PRELOAD_LIBS", "").split(":") if p]
if _libs:
    import torch
    for _p in _libs:
        try:
            torch.ops.load_library(_p)
        except Exception as _e:  # noqa: BLE001
            logger.warning("could not preload %s: %r", _p, _e)


def _transformers5_config_glue():
    """transformers 5 folds `rope_theta` / `rope_scaling` of a checkpoint's config.json into one `rope_parameters`
    dict (`rope_scaling` becomes an alias of it, `rope_theta` disappears). The reference is written against
    transformers 4 (`config.rope_theta`, `config.rope_scaling is None` for plain RoPE, models/qwen3.py:44-56):
    give the loaded config objects back the transformers-4 attribute view. The reference itself is untouched."""
    try:
        import transformers
    except Exception:  # noqa: BLE001
        return
    if getattr(transformers.AutoConfig, "_gllm_ref_glue", False):
        return
    orig = transformers.AutoConfig.from_pretrained.__func__

    def _fix(cfg):
        rp = getattr(cfg, "rope_parameters", None)
        if isinstance(rp, dict) and "rope_type" in rp:
            theta = rp.get("rope_theta")
            rest = {k: v for k, v in rp.items() if k != "rope_theta"}
            plain = rest.get("rope_type", "default") == "default" and "mrope_section" not in rest
            try:
                cfg.rope_scaling = None if plain else rest
            except Exception:  # noqa: BLE001
                pass
            if plain:
                cfg.__dict__["rope_parameters"] = None
            if theta is not None:
                cfg.__dict__["rope_theta"] = theta
        for sub in ("text_config", "vision_config"):
            if getattr(cfg, sub, None) is not None and sub in getattr(cfg, "__dict__", {}):
                _fix(getattr(cfg, sub))
        return cfg

    def from_pretrained(cls, *a, **kw):
        return _fix(orig(cls, *a, **kw))

    transformers.AutoConfig.from_pretrained = classmethod(from_pretrained)
    transformers.AutoConfig._gllm_ref_glue = True


_transformers5_config_glue()
