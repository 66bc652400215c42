"""B200 twins of the reference's hot-path energy plugins.

Each module keeps the reference contract::

    compute_energy_and_gradient_array(mesh, global_params, param_resolver, *,
                                      positions, index_map, grad_arr[, tilts, tilt_grad_arr]) -> float

accumulating (``+=``) into the caller-owned arrays, plus the optional
``compute_energy_array`` and the legacy dict API ``compute_energy_and_gradient``.  On top
of that each module exposes ``B200_MODULE`` (its kernel bit), ``b200_configure`` and
``b200_energy`` so that ``runtime.evaluation_manager.EvaluationManager`` can evaluate all
loaded B200 modules in ONE fused device pass.
"""

NAMES = ("surface", "volume", "bending", "tilt", "bending_tilt")
# leaflet plugins: evaluated by their own device sweeps (csrc/ms_leaflet.cuh), not by the fused patch pass
LEAFLET_NAMES = ("tilt_in", "tilt_out", "bending_tilt_in", "bending_tilt_out", "tilt_smoothness_in",
                 "tilt_smoothness_out", "tilt_smoothness")
