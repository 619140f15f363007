"""swraytracing_b200 -- B200-native engine for SWRaytracing's packet hot path."""
from .engine import (Engine, QGFlow, QG2Flow, SwrtError, load_library, interpolate_dev, g2k_dev, k2g_dev,
                     MODE_SPECTRAL, MODE_LAGRANGE6, MODE_NUFFT, SCHEME_LEAPFROG, SCHEME_RK4_PACKET, SCHEME_RK4_XKA,
                     HIST_INTRINSIC, HIST_ABSOLUTE, FLAG_RHS_GH)

__all__ = ["Engine", "QGFlow", "QG2Flow", "SwrtError", "load_library", "interpolate_dev", "g2k_dev", "k2g_dev",
           "MODE_SPECTRAL", "MODE_LAGRANGE6", "MODE_NUFFT", "SCHEME_LEAPFROG", "SCHEME_RK4_PACKET", "SCHEME_RK4_XKA",
           "HIST_INTRINSIC", "HIST_ABSOLUTE", "FLAG_RHS_GH"]
