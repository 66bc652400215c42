"""Plugin loader of the B200 path.

Twin of ``runtime/energy_manager.py:11-33``: the reference imports ``modules.energy.<name>``;
here the hot-path names resolve to the B200 twins in ``membrane_solver_b200.modules.energy``.
Names without a B200 twin raise -- the new path has no CPU fallback -- unless
``allow_reference=True``, which imports the reference's own module for plugins outside the
hot path (the evaluation manager then calls them through the unchanged array contract).
"""

from __future__ import annotations

import importlib
import logging

from ..modules.energy import LEAFLET_NAMES, NAMES as _FUSED_NAMES

NAMES = _FUSED_NAMES + LEAFLET_NAMES

logger = logging.getLogger("membrane_solver")


class EnergyModuleManager:
    def __init__(self, module_names, *, allow_reference: bool = False):
        self.modules = {}
        for name in module_names:
            if name in self.modules:
                logger.warning("Energy module '%s' listed twice; loading once.", name)
                continue
            if name in NAMES:
                self.modules[name] = importlib.import_module(f"membrane_solver_b200.modules.energy.{name}")
            elif allow_reference:
                self.modules[name] = importlib.import_module(f"modules.energy.{name}")
            else:
                raise ImportError(f"energy module '{name}' has no B200 implementation (available: {', '.join(NAMES)}); "
                                  "pass allow_reference=True to load the reference's module for it")

    def get_module(self, mod):
        """``EnergyModuleManager.get_module`` (``energy_manager.py:27-33``)."""
        if mod not in self.modules:
            raise KeyError(f"Energy module '{mod}' not found.")
        return self.modules[mod]


def install(enforce: bool = False) -> list[str]:
    """Register the B200 twins under the reference's module names, so that an unmodified
    ``runtime.energy_manager.EnergyModuleManager`` (``importlib.import_module(f"modules.energy.{name}")``,
    ``energy_manager.py:21``) and ``runtime.constraint_manager`` load them.  ``enforce``: also replace the hard volume
    projection ``modules.constraints.volume.enforce_constraint`` by its array twin.  Off by default: the twin always
    uses the CURRENT ``dV/dx``, whereas the reference's projection can start from a stale cached gradient
    (``Body.compute_volume`` refreshes the cached version but not the cached gradient dict, ``geometry/body.py:70-148``
    vs ``:401-410``), so trajectories of the reference that went through that quirk are only reproduced with the
    reference's own function.  Returns the names bound."""
    import sys

    bound = []
    for name in NAMES:
        mod = importlib.import_module(f"membrane_solver_b200.modules.energy.{name}")
        sys.modules[f"modules.energy.{name}"] = mod
        pkg = sys.modules.get("modules.energy")
        if pkg is not None:
            setattr(pkg, name, mod)
        bound.append(f"modules.energy.{name}")
    cmod = importlib.import_module("membrane_solver_b200.modules.constraints.volume")
    ref = sys.modules.get("modules.constraints.volume")
    if ref is not None:
        ref.constraint_gradients_array = cmod.constraint_gradients_array
        ref.constraint_gradients = cmod.constraint_gradients
        bound.append("modules.constraints.volume.constraint_gradients[_array]")
        if enforce:
            ref.enforce_constraint = cmod.enforce_constraint
            bound.append("modules.constraints.volume.enforce_constraint")
    return bound
