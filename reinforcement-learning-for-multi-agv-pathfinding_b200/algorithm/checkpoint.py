"""Checkpoint compatibility with the reference.

The reference saves WHOLE MODULES -- torch.save(self.agent.policy_network, ".../policy_net.pt")
(MADQN_structure/Controller.py:161-172, DQN_structure/Controller.py:337-344) -- and `use_NN` loads them
with torch.load (Controller.py:57-60).  A whole-module pickle names its class by module path:
`algorithm.MADQN_structure.MADQN.Net` for AG-DQN (the controllers import `algorithm....` with src/ on
sys.path), `DQN_structure.DQN.Net` for the shipped DQN checkpoints (saved with src/algorithm on the path).
This package's classes live under `reinforcement-learning-for-multi-agv-pathfinding_b200.algorithm....`, so:

  save_module(net, path)   pickles `net` under the REFERENCE's class path (an alias class whose
                           __module__ / __qualname__ are the reference's), hence the reference's
                           `use_NN` torch.load works on files written here;
  load_module(path)        loads files written by the reference (or by save_module) in a process that
                           does not have the reference on sys.path: the reference's module paths are
                           aliased to this package's network classes for the duration of the load.

Layer names are identical (Net mirrors), so state_dicts interchange freely as well.
"""
from __future__ import annotations

import contextlib
import importlib
import sys
import types

import torch

_PKG = __name__.rsplit(".", 1)[0]          # "...b200.algorithm"

# reference module path -> module of this package that holds the same-named classes
REFERENCE_PATHS = {
    "algorithm.MADQN_structure.MADQN": _PKG + ".MADQN_structure.MADQN",
    "src.algorithm.MADQN_structure.MADQN": _PKG + ".MADQN_structure.MADQN",
    "DQN_structure.DQN": _PKG + ".DQN_structure.DQN",
    "algorithm.DQN_structure.DQN": _PKG + ".DQN_structure.DQN",
}


def _alias_class(cls, ref_module: str):
    """subclass of `cls` that pickles as `<ref_module>.<cls.__name__>`"""
    return type(cls.__name__, (cls,), {"__module__": ref_module, "__qualname__": cls.__name__})


@contextlib.contextmanager
def reference_aliases():
    """While active, the reference's module paths resolve (for pickle) to alias classes of this package's
    networks -- unless the real reference modules are already imported, which are then left alone."""
    added, created = [], {}
    try:
        for ref, ours in REFERENCE_PATHS.items():
            if ref in sys.modules:
                continue
            mod = importlib.import_module(ours)
            parts = ref.split(".")
            for i in range(1, len(parts)):          # parent packages
                name = ".".join(parts[:i])
                if name not in sys.modules:
                    sys.modules[name] = types.ModuleType(name)
                    added.append(name)
            alias = types.ModuleType(ref)
            for attr in ("Net",):
                if hasattr(mod, attr):
                    setattr(alias, attr, created.setdefault((ref, attr), _alias_class(getattr(mod, attr), ref)))
            sys.modules[ref] = alias
            added.append(ref)
        yield
    finally:
        for name in reversed(added):
            sys.modules.pop(name, None)


def save_module(net: torch.nn.Module, path: str, ref_module: str = "algorithm.MADQN_structure.MADQN"):
    """torch.save(whole module) in the reference's format: class path `<ref_module>.Net`, CPU tensors."""
    with reference_aliases():
        holder = sys.modules[ref_module]
        cls = getattr(holder, type(net).__name__ if hasattr(holder, type(net).__name__) else "Net")
        clone = cls.__new__(cls)
        torch.nn.Module.__init__(clone)
        # same construction-time attributes and submodules as the source module, parameters copied to the CPU
        src = net.module if hasattr(net, "module") and isinstance(net.module, torch.nn.Module) else net
        import copy
        cpu = copy.deepcopy(src).to("cpu").to(memory_format=torch.contiguous_format)
        clone.__dict__.update(cpu.__dict__)
        torch.save(clone, path)


def load_module(path: str, map_location="cpu") -> torch.nn.Module:
    """torch.load of a whole-module checkpoint written by the reference or by save_module"""
    with reference_aliases():
        return torch.load(path, map_location=map_location, weights_only=False)
