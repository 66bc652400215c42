"""Energy / constraint plugins with the reference's module contract (modules/energy, modules/constraints)."""
