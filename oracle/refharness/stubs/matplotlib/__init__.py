"""Import stub for `matplotlib` (highway_env/vehicle/dynamics.py imports pyplot)."""
