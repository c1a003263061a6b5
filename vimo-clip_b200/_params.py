"""Parameter access and packed-weight caching that also work on ``torch.nn.DataParallel`` REPLICAS.

The reference wraps every model in ``torch.nn.DataParallel`` (train.py:64, inference.py:80, TFAM/train_and_eval.py:392).
``replicate()`` builds, on every forward, one shallow copy of each module per GPU whose ``_parameters`` are EMPTY: the
broadcast copies of the parameters are set as plain tensor attributes (non-leaf, attached to autograd through ``Broadcast``)
and listed in ``_former_parameters``.  Two consequences for a module that keeps raw-pointer packed weights:

* ``named_parameters()`` of a replica yields nothing, so the training Functions must collect the replica's tensors by walking
  the modules (``named_tensors``) -- gradients then flow back to the real parameters through ``Broadcast``;
* a cache stored on the replica dies with it.  ``SharedCache`` lives in the module's ``__dict__``, which ``replicate()`` copies
  shallowly, so the replicas of every forward see the SAME cache object; entries are keyed per device and validated against
  the parameters of the module the cache was created in (the wrapped ``.module``), which is what actually changes when an
  optimiser steps or a checkpoint is loaded.
"""
from __future__ import annotations

import threading
import weakref


def _own_tensors(mod):
    params = mod._parameters
    if any(v is not None for v in params.values()):
        return params
    former = getattr(mod, "_former_parameters", None)
    return former if former else params


def named_tensors(module, prefix: str = ""):
    """``[(qualified name, tensor)]`` of every parameter below ``module``: the ``nn.Parameter`` objects on a normal module, the
    broadcast copies on a DataParallel replica (same names, same order as ``named_parameters()``)."""
    out = []
    for mod_name, mod in module.named_modules(prefix=prefix, remove_duplicate=False):
        for key, t in _own_tensors(mod).items():
            if t is not None:
                out.append((f"{mod_name}.{key}" if mod_name else key, t))
    return out


class SharedCache:
    """Per-device cache of packed weights, shared by a module and its DataParallel replicas (see the module docstring)."""

    def __init__(self, owner=None):
        self._owner = weakref.ref(owner) if owner is not None else None
        self._slots = {}
        self._lock = threading.Lock()

    def __deepcopy__(self, memo):
        return SharedCache()  # a deep copy of the module gets an empty cache, bound on first use

    def __getstate__(self):  # torch.save(module): packed weights are rebuilt after loading
        return {}

    def __setstate__(self, state):
        self._owner, self._slots, self._lock = None, {}, threading.Lock()

    def bind(self, module) -> None:
        if self._owner is None or self._owner() is None:
            self._owner = weakref.ref(module)

    def signature(self, module):
        owner = self._owner() if self._owner is not None else None
        src = owner if owner is not None else module
        return tuple((t.data_ptr(), t._version) for _, t in named_tensors(src))

    def get(self, key, sig, build):
        with self._lock:
            hit = self._slots.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        val = build()
        with self._lock:
            self._slots[key] = (sig, val)
        return val

    def peek(self, key):
        with self._lock:
            hit = self._slots.get(key)
        return None if hit is None else hit[1]

    def put(self, key, val) -> None:
        with self._lock:
            self._slots[key] = (None, val)
