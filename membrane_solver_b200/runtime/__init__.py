"""Host-side mirror of the reference's evaluation surface (runtime/energy_manager.py,
runtime/evaluation_manager.py, runtime/energy_context.py) for the B200 path."""
