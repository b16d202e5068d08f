"""Per-module cache of weights in tensor-core operand form (bf16 copy / fp32 / tf32 hi-lo split).

Entries are keyed by (name, precision) and invalidated when the source parameters change
(torch bumps ``_version`` on every in-place update, e.g. an optimizer step or load_state_dict)."""
import torch

from .. import ops

VALID_PRECISIONS = ("bf16", "tf32", "fp32")


def get_precision(module):
    return getattr(module, "precision", "bf16")


def prepared(module, name, prec, params, build=None):
    """Return ops.prepare_weight(build(*params)) cached on ``module``."""
    cache = module.__dict__.setdefault("_nsp_cache", {})
    ver = tuple((p.data_ptr(), p._version) for p in params)
    hit = cache.get((name, prec))
    if hit is not None and hit[0] == ver:
        return hit[1]
    with torch.no_grad():
        w = build(*params) if build is not None else params[0]
        val = ops.prepare_weight(w, prec)
    cache[(name, prec)] = (ver, val)
    return val


def act_dtype(prec):
    return torch.bfloat16 if prec == "bf16" else torch.float32


def cached(module, name, params, build):
    """Generic per-module cache of a tensor derived from parameters (same invalidation rule as ``prepared``).
    Caches live on the module object: a global cache keyed by data_ptr would go stale when memory is reused."""
    cache = module.__dict__.setdefault("_nsp_cache", {})
    ver = tuple((p.data_ptr(), p._version) for p in params)
    hit = cache.get((name, "raw"))
    if hit is not None and hit[0] == ver:
        return hit[1]
    with torch.no_grad():
        val = build(*params)
    cache[(name, "raw")] = (ver, val)
    return val


def invalidate_caches(module):
    """Drop every cached operand (bf16 copies, transposes, folded BatchNorm weights, tap tables) under ``module``.

    The caches are keyed by ``(data_ptr, _version)`` of their source parameters.  Updates made through ``p.data`` (the
    reference's forget-gate initialisation models/base.py:74, legacy weight noise / clipping / EMA swaps) do NOT bump
    ``_version``: call this after such an update (``load_state_dict``, optimizer steps and every other in-place update through
    the parameter itself are tracked automatically)."""
    for m in module.modules():
        m.__dict__.pop("_nsp_cache", None)
