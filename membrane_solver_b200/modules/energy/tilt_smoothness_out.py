"""Outer-leaflet tilt smoothness (Dirichlet) energy on the B200 path.

Twin of ``modules/energy/tilt_smoothness_out.py`` (which forwards to ``tilt_smoothness_leaflet.py:17-79; tilt gradient only, no shape gradient``): same contract
(``USES_TILT_LEAFLETS``, ``tilts_in`` / ``tilts_out`` / ``tilt_in_grad_arr`` / ``tilt_out_grad_arr`` keywords,
``+=`` into caller-owned arrays, ``grad_arr=None`` = tilt-only evaluation).  See ``_leaflet.py``.
"""

from . import _common as C
from . import _leaflet

USES_TILT_LEAFLETS = True
B200_LEAFLET = ("out", C.L.MOD_TILT_SMOOTHNESS)

compute_energy_and_gradient_array, compute_energy_array, compute_energy_and_gradient = _leaflet.make_module(*B200_LEAFLET)

__all__ = ["compute_energy_and_gradient_array", "compute_energy_array", "compute_energy_and_gradient"]
