"""polarimetry_b200 -- B200-native (sm_100a) per-pixel polarimetric analysis, drop-in for the hot path of
PyPOLAR (cchandre/Polarimetry): see DESIGN.md. Host code is Python; the work is done by libpolarb200.so."""
from .api import (AnalysisParams, Calibration, Engine, Result, Variable, compute_fields, binning, get_intensity, compute_dark,  # noqa: F401
                  analyze_stack, histogram, install, uninstall, harmonic_basis)

__all__ = ['AnalysisParams', 'Calibration', 'Engine', 'Result', 'Variable', 'compute_fields', 'binning', 'get_intensity', 'compute_dark', 'analyze_stack', 'histogram', 'install', 'uninstall']
