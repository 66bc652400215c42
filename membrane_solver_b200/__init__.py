"""membrane_solver_b200: B200-native energy + gradient path of membrane_solver."""
