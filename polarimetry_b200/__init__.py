"""polarimetry_b200 -- B200-native (sm_100a) per-pixel polarimetric analysis, drop-in for the hot path of
PyPOLAR (cchandre/Polarimetry): see DESIGN.md. Host code is Python; the work is done by libpolarb200.so."""
from .api import AnalysisParams, Calibration, Engine, Result, Variable, compute_fields, binning, get_intensity, compute_dark, install, harmonic_basis  # noqa: F401

__all__ = ['AnalysisParams', 'Calibration', 'Engine', 'Result', 'Variable', 'compute_fields', 'binning', 'get_intensity', 'compute_dark', 'install']
