"""The shield / traffic / density / reward / lateral-control cases of soak_parity.py and soak_repro.py."""
CASES = [("cbf-cav", "cav", 3, "default", "steer"), ("cbf-cav", "mixed", 3, "srew", "steer"), ("cbf-avs_cint", "cav", 3, "default", "steer"),
         ("cbf-avs_cint", "mixed", 2, "mrew", "steer"), ("none", "mixed", 1, "default", "steer"), ("cbf-cav", "cav", 1, "mrew", "steer"),
         ("cbf-cav", "mixed", 3, "default", "steer_vel"), ("cbf-cav", "cav", 2, "srew", "steer"), ("cbf-avs_cint", "mixed", 3, "default", "steer_vel"),
         ("cbf-cav", "av", 3, "default", "steer"), ("cbf-avs_cint", "av", 2, "srew", "steer")]
