"""swraytracing_b200 -- B200-native engine for SWRaytracing's packet hot path."""
from .engine import (Engine, QGFlow, SwrtError, load_library, interpolate_dev, g2k_dev, k2g_dev,
                     MODE_SPECTRAL, MODE_LAGRANGE6, SCHEME_LEAPFROG, SCHEME_RK4_PACKET, SCHEME_RK4_XKA,
                     HIST_INTRINSIC, HIST_ABSOLUTE)

__all__ = ["Engine", "QGFlow", "SwrtError", "load_library", "interpolate_dev", "g2k_dev", "k2g_dev",
           "MODE_SPECTRAL", "MODE_LAGRANGE6", "SCHEME_LEAPFROG", "SCHEME_RK4_PACKET", "SCHEME_RK4_XKA",
           "HIST_INTRINSIC", "HIST_ABSOLUTE"]
