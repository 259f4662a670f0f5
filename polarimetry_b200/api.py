"""Host-side API of polarimetry_b200: the PyPOLAR analysis path on a B200, behind the reference's own names.

Two layers:
  * `Engine` -- one per GPU; owns a `pb_ctx` of libpolarb200.so. `Engine.analyze()` is the fused entry
    (raw stack in -> rho/psi/S2/S4/eta maps, intensity map, mask, histograms out), the equivalent of
    `Polarimetry.analyze_stack` (PyPOLAR.py:2948-2979) without the GUI. `analyze_batch` / `analyze_host`
    are the many-stack and host-buffer forms.
  * drop-in functions with the reference's signatures -- `compute_fields` (PP:2981), `binning` (PP:2942),
    `get_intensity` (PC:260), `histo` (PC:339-369 numerics) -- and `install()` which rebinds them on an
    imported reference module.
PyTorch is used for device memory and streams only; all arithmetic is in the CUDA library. There is no
CPU fallback: every entry point raises if the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field as _field
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from ._cabi import PbOutputs, PbParams

METHODS = ('1PF', 'CARS', 'SRS', 'SHG', '2PF', '4POLAR 2D', '4POLAR 3D')
VAR_NAMES = {'1PF': ['Rho', 'Psi'], 'CARS': ['Rho', 'S2', 'S4'], 'SRS': ['Rho', 'S2', 'S4'], '2PF': ['Rho', 'S2', 'S4'],
             'SHG': ['Rho', 'S2', 'S4', 'S_SHG'], '4POLAR 2D': ['Rho', 'Psi'], '4POLAR 3D': ['Rho', 'Psi', 'Eta']}
# default display ranges of the reference GUI (PyPOLAR.py:1719-1730)
DEFAULT_RANGES = {'Rho': (0.0, 180.0), 'Rho_contour': (0.0, 180.0), 'Rho_angle': (0.0, 180.0), 'Psi': (40.0, 180.0),
                  'S2': (0.0, 1.0), 'S4': (-1.0, 1.0), 'S_SHG': (-1.0, 1.0), 'Eta': (0.0, 90.0)}


@dataclass
class AnalysisParams:
    """The settings `analyze_stack` reads from the GUI (PP:2950-2968); defaults as in `startup` (PP:626-639)."""
    method: str = '1PF'
    dark: float = 0.0
    removed_intensity: float = 0.0
    bins: Tuple[int, int] = (1, 1)          # (size along H, size along W) = (bin_spinboxes[0], [1]), PP:2943-2945
    polar_dir: str = 'clockwise'
    offset_angle: float = 0.0
    rotation: Tuple[float, float, float] = (0.0, 0.0, 0.0)    # stick, figure, reference (PP:552)
    chi2threshold: float = 500.0
    ilow: float = 0.0
    nbins: int = 60
    ranges: Dict[str, Tuple[float, float]] = _field(default_factory=lambda: dict(DEFAULT_RANGES))
    out_dtype: str = 'float64'              # 'float64' (bit-comparable maps) or 'float32' (throughput mode)
    exact_chi2: bool = False                # disable the certified chi2 screening
    hist_match: bool = False                # warp-aggregate histogram updates with a match on the bin (aligned samples)
    strip_march: bool = False               # binned 1PF 18 angles 3x3: the strip-march kernel + redo pass (same results; experimental)
    force_unfused: bool = False             # run the separate kernels instead of the fused one


class Calibration:
    """Disk-cone table (1PF) or inverse polarization matrix (4POLAR): the data side of the reference's
    `Calibration` (PC:411-464). The tables themselves are the user's / the reference's data files."""

    _next_uid = 0

    def __init__(self, rho_table=None, psi_table=None, n: Optional[int] = None, invK=None, name: str = ''):
        # never-recycled identity for the engine's upload cache (id() is reused as soon as an object is freed)
        Calibration._next_uid += 1
        self.uid = Calibration._next_uid
        self.name = name
        self.rho_table = None if rho_table is None else np.ascontiguousarray(rho_table, dtype=np.float64)
        self.psi_table = None if psi_table is None else np.ascontiguousarray(psi_table, dtype=np.float64)
        if self.rho_table is not None:
            n = int(n if n is not None else self.rho_table.shape[0])
            self.grid = np.linspace(-1, 1, n, dtype=np.float64)           # PC:463
        self.invKmat = None if invK is None else np.ascontiguousarray(invK, dtype=np.float64)

    @classmethod
    def from_mat(cls, path) -> 'Calibration':
        """Disk-cone `.mat` with RoTest / PsiTest / NbMapValues (PC:460-464)."""
        from scipy.io import loadmat
        d = loadmat(str(path), variable_names=['RoTest', 'PsiTest', 'NbMapValues'], squeeze_me=True)
        return cls(d['RoTest'], d['PsiTest'], int(d['NbMapValues']), name=Path(path).stem)

    @classmethod
    def from_npz(cls, path) -> 'Calibration':
        d = np.load(str(path))
        return cls(d['RoTest'], d['PsiTest'], int(d['NbMapValues']), name=Path(path).stem)

    @classmethod
    def from_K(cls, K, name: str = '') -> 'Calibration':
        """4POLAR: invKmat = pinv(K) (PC:450); K is 4x3 (2D) or 4x4 (3D)."""
        return cls(invK=np.linalg.pinv(np.asarray(K, dtype=np.float64)), name=name)

    @classmethod
    def from_K_txt(cls, path) -> 'Calibration':
        return cls.from_K(np.genfromtxt(str(path), dtype=np.float64), name=Path(path).stem)


class Variable:
    """Result container with the reference's field names (PC:307-327): `.name`, `.values` [H,W]."""

    def __init__(self, name: str, values):
        self.name, self.values = name, values


@dataclass
class Result:
    vars: List[Variable]
    intmap: Optional[torch.Tensor]
    mask: Optional[torch.Tensor]
    intensity: Optional[torch.Tensor]
    hist: Dict[Tuple[str, Optional[int]], np.ndarray]
    roi_map: Optional[np.ndarray] = None
    plan: str = ''

    def get_var(self, name):
        return next((v for v in self.vars if v.name == name), None)

    @property
    def xmap(self):
        """Column index where the pixel is valid, NaN elsewhere (PP:3056-3058), built lazily."""
        H, W = self.mask.shape[-2:]
        xx = torch.arange(W, device=self.mask.device, dtype=torch.float64).expand(H, W)
        return torch.where(self.mask.bool(), xx, torch.full_like(xx, float('nan')))

    @property
    def ymap(self):
        H, W = self.mask.shape[-2:]
        yy = torch.arange(H, device=self.mask.device, dtype=torch.float64)[:, None].expand(H, W)
        return torch.where(self.mask.bool(), yy, torch.full_like(yy, float('nan')))


def harmonic_basis(nangle: int, polar_dir: str, offset_angle: float):
    """cos/sin of 2a and 4a computed by NumPy exactly as the reference does (PP:2983-2986, PP:2993)."""
    pd = -1 if polar_dir == 'clockwise' else 1
    alpha = pd * np.linspace(0, 180, nangle, endpoint=False) + offset_angle
    e2 = np.exp(2j * np.deg2rad(alpha))
    e4 = e2 ** 2
    return [np.ascontiguousarray(x) for x in (e2.real, e2.imag, e4.real, e4.imag)]


def rasterize_rois(rois: Sequence[dict], shape: Tuple[int, int], base_mask: Optional[np.ndarray] = None):
    """Selected ROI polygons -> (roi_bits uint32 [H,W], ilow list, indices). Same rasteriser and the same
    int32 truncation of the vertices as the reference (cv2.fillPoly, PP:2883-2889)."""
    import cv2
    sel = [r for r in (rois or []) if r.get('select')]
    bits = np.zeros(shape, dtype=np.uint32)
    if len(sel) > _cabi.PB_MAX_ROIS:
        raise ValueError(f'at most {_cabi.PB_MAX_ROIS} selected ROIs per call')
    tmp = np.zeros(shape, dtype=np.uint8)
    for k, r in enumerate(sel):
        verts = np.array([np.asarray(r['vertices']).T], dtype=np.int32)
        tmp.fill(0)
        cv2.fillPoly(tmp, verts, 1)
        bits |= (tmp == 1).astype(np.uint32) << np.uint32(k)
    if not sel:
        bits[:] = 1
    if base_mask is not None:
        bits[np.asarray(base_mask) == 0] = 0
    return bits, [float(r['ILow']) for r in sel], [r['indx'] for r in sel]


def _ptr(t) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr() if isinstance(t, torch.Tensor) else t.ctypes.data


class Engine:
    """One analysis context per GPU."""

    def __init__(self, device: Optional[int] = None):
        self.lib = _cabi.load()
        if not torch.cuda.is_available():
            raise RuntimeError('polarimetry_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device('cuda', self.device_index)
        self._ctx = C.c_void_p()
        rc = self.lib.pb_create(C.byref(self._ctx), self.device_index)
        if rc != 0:
            raise _cabi.PbError('pb_create failed: ' + self.lib.pb_last_error(None).decode())
        self._key_basis = self._key_calib = self._key_hist = self._key_rois = None

    def close(self):
        if getattr(self, '_ctx', None):
            self.lib.pb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state upload -----------------------------------------------------------------------------
    def _ck(self, rc, what):
        _cabi.check(self.lib, self._ctx, rc, what)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _set_basis(self, N, p: AnalysisParams):
        key = (N, p.polar_dir, float(p.offset_angle))
        if key != self._key_basis:
            c2, s2, c4, s4 = harmonic_basis(N, p.polar_dir, p.offset_angle)
            self._ck(self.lib.pb_set_basis(self._ctx, N, _ptr(c2), _ptr(s2), _ptr(c4), _ptr(s4)), 'pb_set_basis')
            self._key_basis = key

    def set_diskcones(self, calibs: Sequence[Calibration]) -> None:
        """Upload several disk cones at once (pb_set_diskcones): they stay resident and `_set_calib` merely selects among them --
        no copy, allocation or synchronisation when the calibration sweep (PP:850-913) moves to its next disk."""
        calibs = list(calibs)
        if not calibs or any(c.rho_table is None for c in calibs):
            raise ValueError('set_diskcones: disk-cone Calibrations required')
        n = calibs[0].grid.size
        if any(c.grid.size != n or not np.array_equal(c.grid, calibs[0].grid) for c in calibs):
            raise ValueError('set_diskcones: the disk cones must share one grid (NbMapValues)')
        rho = np.ascontiguousarray(np.stack([c.rho_table for c in calibs]))
        psi = np.ascontiguousarray(np.stack([c.psi_table for c in calibs]))
        self._key_calib, self._resident = None, {}
        self._ck(self.lib.pb_set_diskcones(self._ctx, len(calibs), _ptr(rho), _ptr(psi), n, _ptr(calibs[0].grid)), 'pb_set_diskcones')
        self._resident = {c.uid: k for k, c in enumerate(calibs)}
        self._key_calib = ('disk', calibs[0].uid)

    def _set_calib(self, p: AnalysisParams, calib: Optional[Calibration]):
        if p.method == '1PF':
            if calib is None or calib.rho_table is None:
                raise ValueError('1PF needs a disk-cone Calibration')
            key = ('disk', calib.uid)
            if key != self._key_calib and calib.uid in getattr(self, '_resident', {}):
                self._key_calib = None
                self._ck(self.lib.pb_select_diskcone(self._ctx, self._resident[calib.uid]), 'pb_select_diskcone')
                self._key_calib = key
            if key != self._key_calib:
                self._resident = {}
                n = calib.grid.size
                self._key_calib = None           # a failed upload must not look cached
                self._ck(self.lib.pb_set_diskcone(self._ctx, _ptr(calib.rho_table), _ptr(calib.psi_table), n, _ptr(calib.grid)), 'pb_set_diskcone')
                self._key_calib = key
        elif p.method.startswith('4POLAR'):
            if calib is None or calib.invKmat is None:
                raise ValueError('4POLAR needs a K-matrix Calibration')
            rows = int(p.method[-2]) + 1
            if calib.invKmat.shape != (rows, 4):
                raise ValueError('Error in the dimension of the calibration matrix')     # PC:451-452
            self._ck(self.lib.pb_set_invk(self._ctx, _ptr(calib.invKmat), rows), 'pb_set_invk')

    def _set_hist(self, p: AnalysisParams):
        names = VAR_NAMES[p.method]
        key = (p.method, p.nbins, tuple(p.ranges[n] for n in names))
        if key != self._key_hist:
            edges = np.ascontiguousarray(np.stack([np.linspace(p.ranges[n][0], p.ranges[n][1], p.nbins + 1) for n in names]))
            is_rho = np.array([int(n == 'Rho') for n in names], dtype=np.int32)
            self._key_hist = None                # a failed upload must not look cached
            self._ck(self.lib.pb_set_hist(self._ctx, len(names), p.nbins, _ptr(edges), _ptr(is_rho)), 'pb_set_hist')
            self._key_hist = key

    def _set_rois(self, ilows: Sequence[float], per_roi: bool):
        n = len(ilows)
        groups = [1 << k for k in range(n)] if (per_roi and n) else [0xffffffff]
        key = (tuple(ilows), tuple(groups))
        if key != self._key_rois:
            il = np.asarray(ilows, dtype=np.float64)
            gb = np.asarray(groups, dtype=np.uint32)
            self._ck(self.lib.pb_set_rois(self._ctx, n, _ptr(il) if n else None, len(groups), _ptr(gb)), 'pb_set_rois')
            self._key_rois = key
        return len(groups)

    def _params(self, p: AnalysisParams, dtype: str, N, H, W) -> PbParams:
        flags = 0
        if p.out_dtype == 'float64':
            flags |= _cabi.FLAG_OUT_F64
        if p.exact_chi2:
            flags |= _cabi.FLAG_EXACT_CHI2
        if p.force_unfused:
            flags |= _cabi.FLAG_FORCE_UNFUSED
        if p.hist_match:
            flags |= _cabi.FLAG_HIST_MATCH
        if p.strip_march:
            flags |= _cabi.FLAG_STRIP_MARCH
        return PbParams(_cabi.METHODS[p.method], _cabi.DTYPES[dtype], N, H, W, int(p.bins[0]), int(p.bins[1]), flags,
                        float(p.dark), float(p.removed_intensity), float(p.rotation[0]), float(p.rotation[1]),
                        float(p.rotation[2]), float(p.chi2threshold), float(p.ilow))

    def _to_device(self, a, dtype=None):
        if isinstance(a, torch.Tensor):
            t = a.to(self.device)
        else:
            a = np.ascontiguousarray(a)
            if a.dtype == np.uint16:      # torch has limited uint16 support: move the bytes
                t = torch.from_numpy(a.view(np.int16)).to(self.device).view(torch.uint16)
            elif a.dtype == np.uint32:
                t = torch.from_numpy(a.view(np.int32)).to(self.device)
            else:
                t = torch.from_numpy(a).to(self.device)
        return t.contiguous() if dtype is None else t.to(dtype).contiguous()

    @staticmethod
    def _dtype_name(t: torch.Tensor) -> str:
        name = str(t.dtype).replace('torch.', '')
        if name not in _cabi.DTYPES:
            raise TypeError(f'unsupported stack dtype {t.dtype}; use uint8, uint16, float32 or float64')
        return name

    # ---- the fused path ----------------------------------------------------------------------------
    def analyze_batch(self, stacks, p: AnalysisParams, calib: Optional[Calibration] = None, rois: Optional[Sequence[dict]] = None,
                      per_roi: bool = False, base_mask=None, edge_contours=None, want_intensity: bool = True,
                      want_hist: bool = True, hist_per_stack: bool = False, out: Optional[dict] = None) -> dict:
        """stacks [B,N,H,W] (device tensor or array). Returns dict of device tensors: 'vars' {name: [B,H,W]},
        'intmap', 'mask', 'intensity' and 'hist' int64 [B or 1, n_groups, n_vars, nbins] (summed over the
        batch unless hist_per_stack). Everything is enqueued on torch's current stream; nothing synchronises."""
        t = self._to_device(stacks)
        if t.dim() != 4:
            raise ValueError('stacks must be [B,N,H,W]')
        B, N, H, W = t.shape
        names = VAR_NAMES[p.method]
        if not p.method.startswith('4POLAR'):
            self._set_basis(N, p)
        self._set_calib(p, calib)
        if want_hist:
            self._set_hist(p)
        bits_np, ilows, indices = (None, [], [])
        roi_dev = None
        if rois or base_mask is not None:
            bits_np, ilows, indices = rasterize_rois(rois, (H, W), base_mask)
            roi_dev = self._to_device(bits_np)
        ng = self._set_rois(ilows, per_roi)
        edge_dev = None if edge_contours is None else self._to_device(edge_contours, torch.float64)
        pp = self._params(p, self._dtype_name(t), N, H, W)
        if not want_hist:
            pp.flags |= _cabi.FLAG_NO_HIST
        mdt = torch.float64 if p.out_dtype == 'float64' else torch.float32
        o = out if out is not None else {}
        dev = self.device
        if 'vars' not in o:
            o['vars'] = {n: torch.empty((B, H, W), dtype=mdt, device=dev) for n in names}
            o['intmap'] = torch.empty((B, H, W), dtype=mdt, device=dev)
            o['mask'] = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
            o['intensity'] = torch.empty((B, H, W), dtype=torch.float64, device=dev) if want_intensity else None
            if float(p.rotation[2]) != 0.0:
                o['rho_angle'] = torch.empty((B, H, W), dtype=mdt, device=dev)
            if edge_dev is not None:
                o['rho_contour'] = torch.empty((B, H, W), dtype=mdt, device=dev)
        nb = B if hist_per_stack else 1
        if want_hist:
            if 'hist' not in o:
                o['hist'] = torch.zeros((nb, ng, len(names), p.nbins), dtype=torch.int64, device=dev)
            else:
                o['hist'].zero_()
        outs = (PbOutputs * B)()
        for b in range(B):
            for k, n in enumerate(names):
                outs[b].var[k] = o['vars'][n][b].data_ptr()
            outs[b].intmap = o['intmap'][b].data_ptr()
            outs[b].mask = o['mask'][b].data_ptr()
            if o.get('intensity') is not None:
                outs[b].intensity = o['intensity'][b].data_ptr()
            if 'rho_angle' in o:
                outs[b].rho_angle = o['rho_angle'][b].data_ptr()
            if 'rho_contour' in o:
                outs[b].rho_contour = o['rho_contour'][b].data_ptr()
            if want_hist:
                outs[b].hist = o['hist'][b if hist_per_stack else 0].data_ptr()
        self._ck(self.lib.pb_analyze(self._ctx, C.byref(pp), t.data_ptr(), B, _ptr(roi_dev), _ptr(edge_dev), outs, self._stream()), 'pb_analyze')
        o['plan'] = self.lib.pb_plan(self._ctx, C.byref(pp)).decode()
        o['roi_indices'] = indices if (per_roi and indices) else [None]
        o['roi_bits'] = bits_np
        o['_keep'] = (t, roi_dev, edge_dev)
        return o

    def analyze(self, stack, p: AnalysisParams, calib: Optional[Calibration] = None, rois: Optional[Sequence[dict]] = None,
                per_roi: bool = False, base_mask=None, edge_contours=None, want_hist: bool = True) -> Result:
        """One stack [N,H,W] -> Result (maps on the device, histograms on the host)."""
        t = self._to_device(stack)
        o = self.analyze_batch(t[None], p, calib, rois, per_roi, base_mask, edge_contours, True, want_hist)
        names = VAR_NAMES[p.method]
        vars_ = [Variable(n, o['vars'][n][0]) for n in names]
        rho = vars_[0].values
        if 'rho_contour' in o:                                            # PP:3059-3068
            edge = o['_keep'][2]
            both = torch.isfinite(rho) & torch.isfinite(edge)
            vars_.append(Variable('Rho_contour', o['rho_contour'][0]))
            for v in list(vars_[1:-1]):
                vars_.append(Variable(v.name + '_contour', torch.where(both, v.values, torch.full_like(v.values, float('nan')))))
        if 'rho_angle' in o:                                              # PP:3069-3072
            vars_.append(Variable('Rho_angle', o['rho_angle'][0]))
        hist = {}
        if want_hist:
            h = o['hist'][0].cpu().numpy()
            for g, idx in enumerate(o['roi_indices']):
                for k, n in enumerate(names):
                    hist[(n, idx)] = h[g, k]
            # extra variables: standalone histogram kernel (PC:340-369)
            for v in vars_[len(names):]:
                if v.name in p.ranges:
                    for g, idx in enumerate(o['roi_indices']):
                        sel = None
                        if o['roi_bits'] is not None:
                            gb = (1 << g) if (per_roi and idx is not None) else 0xffffffff
                            sel = self._to_device(((o['roi_bits'] & np.uint32(gb)) != 0).astype(np.uint8))
                        hist[(v.name, idx)] = self.histo(v.values, sel, v.name, *p.ranges[v.name], float(p.rotation[1]), p.nbins)
        return Result(vars_, o['intmap'][0], o['mask'][0], o['intensity'][0] if o.get('intensity') is not None else None,
                      hist, o['roi_bits'], o['plan'])

    def analyze_host(self, stacks_host, p: AnalysisParams, calib: Optional[Calibration] = None, out: Optional[dict] = None,
                     want_hist: bool = True, want_maps: bool = True) -> dict:
        """Host buffers in, host buffers out, through `pb_analyze_host` (H2D / kernels / D2H overlapped over two
        streams). stacks_host: numpy array or pinned CPU tensor [B,N,H,W]. Returns dict of host arrays.
        want_maps=False: only the reduced histograms come back (batch mode of BASELINE configs[3]: no map D2H)."""
        a = stacks_host if isinstance(stacks_host, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(stacks_host).view(np.int16) if np.asarray(stacks_host).dtype == np.uint16 else np.ascontiguousarray(stacks_host))
        src_dtype = 'uint16' if (not isinstance(stacks_host, torch.Tensor) and np.asarray(stacks_host).dtype == np.uint16) else self._dtype_name(a)
        B, N, H, W = a.shape
        names = VAR_NAMES[p.method]
        if not p.method.startswith('4POLAR'):
            self._set_basis(N, p)
        self._set_calib(p, calib)
        if want_hist:
            self._set_hist(p)
        ng = self._set_rois([], False)
        pp = self._params(p, src_dtype, N, H, W)
        if not want_hist:
            pp.flags |= _cabi.FLAG_NO_HIST
        mdt = torch.float64 if p.out_dtype == 'float64' else torch.float32
        o = out if out is not None else {}
        if want_maps and 'vars' not in o:
            pin = dict(pin_memory=True)
            o['vars'] = {n: torch.empty((B, H, W), dtype=mdt, **pin) for n in names}
            o['intmap'] = torch.empty((B, H, W), dtype=mdt, **pin)
            o['mask'] = torch.empty((B, H, W), dtype=torch.uint8, **pin)
        if 'hist' not in o:
            o['hist'] = torch.zeros((1, ng, len(names), p.nbins), dtype=torch.int64)
        o['hist'].zero_()
        outs = (PbOutputs * B)()
        for b in range(B):
            if want_maps:
                for k, n in enumerate(names):
                    outs[b].var[k] = o['vars'][n][b].data_ptr()
                outs[b].intmap = o['intmap'][b].data_ptr()
                outs[b].mask = o['mask'][b].data_ptr()
            if want_hist:
                outs[b].hist = o['hist'][0].data_ptr()
        self._ck(self.lib.pb_analyze_host(self._ctx, C.byref(pp), a.data_ptr(), B, None, None, outs), 'pb_analyze_host')
        return o

    # ---- the individual seams ----------------------------------------------------------------------
    def get_intensity(self, values, dark: float = 0.0, bin: Sequence[int] = (1, 1)) -> torch.Tensor:
        """Stack.get_intensity (PC:260-264)."""
        t = self._to_device(values)
        N, H, W = t.shape
        out = torch.empty((H, W), dtype=torch.float64, device=self.device)
        self._ck(self.lib.pb_get_intensity(self._ctx, t.data_ptr(), _cabi.DTYPES[self._dtype_name(t)], N, H, W, float(dark),
                                           int(bin[0]), int(bin[1]), out.data_ptr(), self._stream()), 'pb_get_intensity')
        return out

    def compute_dark(self, values, size_cell: int = 20) -> float:
        """Stack.compute_dark (PC:266-278) on the device (integer stacks)."""
        t = self._to_device(values)
        N, H, W = t.shape
        out = C.c_double(0.0)
        self._ck(self.lib.pb_compute_dark(self._ctx, t.data_ptr(), _cabi.DTYPES[self._dtype_name(t)], N, H, W, int(size_cell),
                                          C.byref(out), self._stream()), 'pb_compute_dark')
        return float(out.value)

    def from_hwn(self, stack_hwn) -> torch.Tensor:
        """Angle-minor stack(s) [..., H, W, N] (the N samples of a pixel contiguous) -> the reference's angle-major [..., N, H, W]
        (PP:1926) on the device, same dtype; every analysis entry point takes the result."""
        t = self._to_device(stack_hwn)
        if t.dim() < 3:
            raise ValueError('from_hwn: [H, W, N] or [B, H, W, N] required')
        lead = t.shape[:-3]
        H, W, N = (int(x) for x in t.shape[-3:])
        src = t.reshape(-1, H, W, N)
        dst = torch.empty((src.shape[0], N, H, W), dtype=torch.int16 if t.dtype == torch.uint16 else t.dtype, device=self.device)
        if t.dtype == torch.uint16:
            dst = dst.view(torch.uint16)
        for b in range(src.shape[0]):
            self._ck(self.lib.pb_stack_from_hwn(self._ctx, src[b].data_ptr(), _cabi.DTYPES[self._dtype_name(t)], N, H, W, dst[b].data_ptr(),
                                                self._stream()), 'pb_stack_from_hwn')
        return dst.reshape(*lead, N, H, W)

    def define_stack(self, source, default_dark: Optional[float] = None, size_cell: int = 20) -> dict:
        """Polarimetry.define_stack (PP:1923-1942, "next" row N3) up to the device: `source` is a multi-page TIFF path (read with
        cv2.imreadmulti(..., IMREAD_ANYDEPTH) exactly as the reference does) or the pages [N,H,W] themselves. Pages go through a
        pinned staging buffer to the device; a non-integer stack is rescaled to uint16 there (pb_rescale_stack, PP:1929-1930);
        dark = default_dark or Stack.compute_dark on the device (PC:266-278, via set_dark PP:2247-2252).
        Returns dict(values=uint16/uint8 device tensor [N,H,W], nangle, height, width, dark)."""
        if isinstance(source, (str, Path)):
            import cv2
            ok, pages = cv2.imreadmulti(str(source), [], cv2.IMREAD_ANYDEPTH)
            if not ok or len(pages) == 0:
                raise FileNotFoundError(f'define_stack: cannot read {source}')
            a = np.asarray(pages)
        elif isinstance(source, torch.Tensor):
            a = source
        else:
            a = np.asarray(source)
        if a.ndim != 3:
            raise ValueError('define_stack: a stack [nangle, height, width] is required (the reference fails on single pages too, PP:1925-1928)')
        N, H, W = (int(x) for x in a.shape)
        if isinstance(a, torch.Tensor):
            t = a.to(self.device).contiguous()
        else:
            a = np.ascontiguousarray(a)
            host = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a)
            t = host.pin_memory().to(self.device, non_blocking=True)
            t = t.view(torch.uint16) if a.dtype == np.uint16 else t
        if t.dtype in (torch.float32, torch.float64):
            out = torch.empty((N, H, W), dtype=torch.int16, device=self.device).view(torch.uint16)
            mm = (C.c_double * 2)()
            self._ck(self.lib.pb_rescale_stack(self._ctx, t.data_ptr(), _cabi.DTYPES[self._dtype_name(t)], t.numel(), out.data_ptr(), mm,
                                               self._stream()), 'pb_rescale_stack')
            t = out
        elif t.dtype not in (torch.uint8, torch.uint16):
            raise TypeError(f'define_stack: {t.dtype} stacks are not supported (uint8 / uint16 / float32 / float64 TIFF pages)')
        dark = float(default_dark) if default_dark is not None else self.compute_dark(t, size_cell)
        return dict(values=t, nangle=N, height=H, width=W, dark=dark)

    def project(self, stacks, p: AnalysisParams, rois: Optional[Sequence[dict]] = None, per_roi: bool = False, base_mask=None) -> dict:
        """The disk-independent part of the 1PF analysis (pre-processing, binning, harmonics, fit / chi2 mask: PP:2953-3000) of
        stacks [B,N,H,W]: dict of device tensors a0, a2r, a2i (float64 [B,H,W]; a2 un-normalised, PP:2989) and mask (uint8).
        Input of `lut_apply`; nothing here depends on the disk cone."""
        if p.method != '1PF':
            raise ValueError('project / lut_apply: the disk-cone inversion exists for 1PF only (PP:3017)')
        t = self._to_device(stacks)
        if t.dim() == 3:
            t = t[None]
        B, N, H, W = t.shape
        self._set_basis(N, p)
        bits_np, ilows, indices = (None, [], [])
        roi_dev = None
        if rois or base_mask is not None:
            bits_np, ilows, indices = rasterize_rois(rois, (H, W), base_mask)
            roi_dev = self._to_device(bits_np)
        self._set_rois(ilows, per_roi)
        pp = self._params(p, self._dtype_name(t), N, H, W)
        pp.flags |= _cabi.FLAG_PROJECT_ONLY
        o = {k: torch.empty((B, H, W), dtype=torch.float64, device=self.device) for k in ('a0', 'a2r', 'a2i')}
        o['mask'] = torch.empty((B, H, W), dtype=torch.uint8, device=self.device)
        outs = (PbOutputs * B)()
        for b in range(B):
            outs[b].var[0] = o['a2r'][b].data_ptr()
            outs[b].var[1] = o['a2i'][b].data_ptr()
            outs[b].intmap = o['a0'][b].data_ptr()
            outs[b].mask = o['mask'][b].data_ptr()
        self._ck(self.lib.pb_analyze(self._ctx, C.byref(pp), t.data_ptr(), B, _ptr(roi_dev), None, outs, self._stream()), 'pb_analyze')
        o.update(plan=self.lib.pb_plan(self._ctx, C.byref(pp)).decode(), roi_indices=indices if (per_roi and indices) else [None],
                 roi_bits=bits_np, ilows=ilows, per_roi=per_roi, shape=(B, N, H, W), src_dtype=self._dtype_name(t), _keep=(t, roi_dev))
        return o

    def lut_apply(self, proj: dict, p: AnalysisParams, calib: Calibration, want_hist: bool = True, out: Optional[dict] = None) -> dict:
        """The disk-dependent tail (PP:2991, PP:3017-3026, PP:3052, PC:340-369) on the maps of `project`, with disk cone `calib`:
        the same dict `analyze_batch` returns (vars, intmap, mask, hist summed over the batch)."""
        B, N, H, W = proj['shape']
        names = VAR_NAMES['1PF']
        self._set_basis(N, p)
        self._set_calib(p, calib)
        if want_hist:
            self._set_hist(p)
        ng = self._set_rois(proj['ilows'], proj['per_roi'])
        pp = self._params(p, proj['src_dtype'], N, H, W)
        if not want_hist:
            pp.flags |= _cabi.FLAG_NO_HIST
        mdt = torch.float64 if p.out_dtype == 'float64' else torch.float32
        o = out if out is not None else {}
        if 'vars' not in o:
            o['vars'] = {n: torch.empty((B, H, W), dtype=mdt, device=self.device) for n in names}
            o['intmap'] = torch.empty((B, H, W), dtype=mdt, device=self.device)
            o['mask'] = torch.empty((B, H, W), dtype=torch.uint8, device=self.device)
        if want_hist:
            if 'hist' not in o:
                o['hist'] = torch.zeros((1, ng, len(names), p.nbins), dtype=torch.int64, device=self.device)
            else:
                o['hist'].zero_()
        outs = (PbOutputs * B)()
        for b in range(B):
            for k, n in enumerate(names):
                outs[b].var[k] = o['vars'][n][b].data_ptr()
            outs[b].intmap = o['intmap'][b].data_ptr()
            outs[b].mask = o['mask'][b].data_ptr()
            if want_hist:
                outs[b].hist = o['hist'][0].data_ptr()
        roi_dev = proj['_keep'][1]
        self._ck(self.lib.pb_lut_apply(self._ctx, C.byref(pp), proj['a0'].data_ptr(), proj['a2r'].data_ptr(), proj['a2i'].data_ptr(),
                                       proj['mask'].data_ptr(), B, _ptr(roi_dev), outs, self._stream()), 'pb_lut_apply')
        o['plan'] = 'lut_apply'
        o['roi_indices'], o['roi_bits'] = proj['roi_indices'], proj['roi_bits']
        return o

    def calibration_sweep(self, stacks: Sequence, disks: Sequence['Calibration'], p: AnalysisParams, rois: Optional[Sequence[dict]] = None,
                          per_roi: bool = False, base_mask=None, project_once: bool = True) -> List[dict]:
        """The 1PF calibration sweep (PP:850-913, "next" row N1): every disk cone against every stack, one statistics row per
        (disk, stack, ROI) as `analyze_stack(for_calib=True)` returns them (numeric columns of return_vecexcel, PP:2739-2779).
        Only the LUT changes between disks: each stack is projected ONCE (`project`, everything up to the mask of PP:3000) and
        the disk-dependent tail (`lut_apply`) runs per disk on the resident maps. project_once=False runs the whole fused
        analysis per (disk, stack) instead (same rows; kept as the cross-check)."""
        dev_stacks = [self._to_device(s) for s in stacks]
        sel_rois = [r for r in (rois or []) if r.get('select')] if per_roi else []
        targets = [(r['indx'], g) for g, r in enumerate(sel_rois)] or [(None, 0)]
        rows = []
        same_shape = len({tuple(st.shape) for st in dev_stacks}) == 1 and len({st.dtype for st in dev_stacks}) == 1
        if project_once and same_shape:
            # the batched form: ONE projection launch for all stacks, the disk cones resident on the device, and per disk one
            # lut_apply launch + two statistics launches (per-stack accumulators); nothing synchronises until the rows are read
            from .batch import finalize_stats
            batch = torch.stack(dev_stacks) if len(dev_stacks) > 1 else dev_stacks[0][None]
            proj = self.project(batch, p, rois, per_roi, base_mask)
            S, ng, D = batch.shape[0], len(proj['roi_indices']), len(disks)
            if any(c.uid not in getattr(self, '_resident', {}) for c in disks):
                self.set_diskcones(disks)
            acc1 = torch.zeros((D, S, ng, _cabi.PB_MAX_VARS + 1, 4), dtype=torch.float64, device=self.device)
            acc2 = torch.zeros((D, S, ng, _cabi.PB_MAX_VARS + 1, 2), dtype=torch.float64, device=self.device)
            out = {}
            for d, cal in enumerate(disks):
                o = self.lut_apply(proj, p, cal, want_hist=False, out=out)
                o['_keep'] = proj['_keep']
                self.batch_stats_accumulate(o, p, None, per_stack=True, acc=(acc1[d], acc2[d]))
            a1, a2 = acc1.cpu().numpy(), acc2.cpu().numpy()
            for d, cal in enumerate(disks):
                for k in range(S):
                    fin = finalize_stats(a1[d, k], a2[d, k], VAR_NAMES['1PF'])
                    for idx, g in targets:
                        row = dict(fin[g])
                        row.update(disk=cal.name or d, stack=k, roi=idx if idx is not None else 1)
                        rows.append(row)
            return rows
        projs = [self.project(st, p, rois, per_roi, base_mask) for st in dev_stacks] if project_once else None
        for d, cal in enumerate(disks):
            for k, st in enumerate(dev_stacks):
                if project_once:
                    o = self.lut_apply(projs[k], p, cal, want_hist=False)
                    res = Result([Variable(n, o['vars'][n][0]) for n in VAR_NAMES['1PF']], o['intmap'][0], o['mask'][0], None, {},
                                 o['roi_bits'], o['plan'])
                else:
                    res = self.analyze(st, p, cal, rois, per_roi, base_mask, want_hist=False)
                for idx, g in targets:
                    row = self.roi_stats(res, p, idx, g)
                    row.update(disk=cal.name or d, stack=k, roi=idx if idx is not None else 1)
                    rows.append(row)
        return rows

    def field(self, values, p: AnalysisParams) -> torch.Tensor:
        """analyze_stack's pre-processing + binning (PP:2953-2962) -> double field [N,H,W]."""
        t = self._to_device(values)
        N, H, W = t.shape
        out = torch.empty((N, H, W), dtype=torch.float64, device=self.device)
        pp = self._params(p, self._dtype_name(t), N, H, W)
        self._ck(self.lib.pb_field(self._ctx, C.byref(pp), t.data_ptr(), out.data_ptr(), self._stream()), 'pb_field')
        return out

    def binning(self, field, bins: Sequence[int]) -> torch.Tensor:
        """Polarimetry.binning (PP:2942-2946) on a double field [N,H,W]; returns a new tensor."""
        t = self._to_device(field, torch.float64).clone()
        N, H, W = t.shape
        self._ck(self.lib.pb_binning(self._ctx, t.data_ptr(), N, H, W, int(bins[0]), int(bins[1]), self._stream()), 'pb_binning')
        return t

    def compute_fields(self, field, mask, p: AnalysisParams, calib: Optional[Calibration] = None, edge_contours=None,
                       want_fit: bool = False, want_hist: bool = False) -> Result:
        """compute_fields (PP:2981-3078) on an already pre-processed, binned double field and an incoming mask."""
        f = self._to_device(field, torch.float64)
        N, H, W = f.shape
        m = self._to_device(np.asarray(mask, dtype=np.uint8) if not isinstance(mask, torch.Tensor) else mask.to(torch.uint8)).clone()
        names = VAR_NAMES[p.method]
        if not p.method.startswith('4POLAR'):
            self._set_basis(N, p)
        self._set_calib(p, calib)
        if want_hist:
            self._set_hist(p)
        self._set_rois([], False)
        pp = self._params(p, 'float64', N, H, W)
        if not want_hist:
            pp.flags |= _cabi.FLAG_NO_HIST
        mdt = torch.float64 if p.out_dtype == 'float64' else torch.float32
        dev = self.device
        vars_ = [Variable(n, torch.empty((H, W), dtype=mdt, device=dev)) for n in names]
        intmap = torch.empty((H, W), dtype=mdt, device=dev)
        edge_dev = None if edge_contours is None else self._to_device(edge_contours, torch.float64)
        o = PbOutputs()
        for k, v in enumerate(vars_):
            o.var[k] = v.values.data_ptr()
        o.intmap = intmap.data_ptr()
        o.mask = m.data_ptr()
        extra = {}
        if float(p.rotation[2]) != 0.0:
            extra['Rho_angle'] = torch.empty((H, W), dtype=mdt, device=dev)
            o.rho_angle = extra['Rho_angle'].data_ptr()
        if edge_dev is not None:
            extra['Rho_contour'] = torch.empty((H, W), dtype=mdt, device=dev)
            o.rho_contour = extra['Rho_contour'].data_ptr()
        hist_t = None
        if want_hist:
            hist_t = torch.zeros((1, len(names), p.nbins), dtype=torch.int64, device=dev)
            o.hist = hist_t.data_ptr()
        self._ck(self.lib.pb_compute_fields(self._ctx, C.byref(pp), f.data_ptr(), m.data_ptr(), None, _ptr(edge_dev), C.byref(o),
                                            self._stream()), 'pb_compute_fields')
        if 'Rho_contour' in extra:
            both = torch.isfinite(vars_[0].values) & torch.isfinite(edge_dev)
            base = list(vars_[1:])
            vars_.append(Variable('Rho_contour', extra['Rho_contour']))
            for v in base:
                vars_.append(Variable(v.name + '_contour', torch.where(both, v.values, torch.full_like(v.values, float('nan')))))
        if 'Rho_angle' in extra:
            vars_.append(Variable('Rho_angle', extra['Rho_angle']))
        res = Result(vars_, intmap, m, None, {}, None, 'compute_fields')
        if want_hist:
            h = hist_t.cpu().numpy()
            res.hist = {(n, None): h[0, k] for k, n in enumerate(names)}
        if want_fit and not p.method.startswith('4POLAR'):
            fit = torch.empty_like(f)
            chi2 = torch.empty((H, W), dtype=torch.float64, device=dev)
            self._ck(self.lib.pb_fit_maps(self._ctx, C.byref(pp), f.data_ptr(), m.data_ptr(), fit.data_ptr(), chi2.data_ptr(),
                                          self._stream()), 'pb_fit_maps')
            res.field_fit, res.chi2 = fit, chi2
        return res

    def histo(self, values, sel, name: str, vmin: float, vmax: float, rotation: float = 0.0, nbins: int = 60,
              htype: str = 'normal') -> np.ndarray:
        """Numeric part of Variable.histo (PC:340-369): int64 counts[nbins]."""
        v = self._to_device(values)
        if v.dtype not in (torch.float32, torch.float64):
            v = v.to(torch.float64)
        s = None if sel is None else self._to_device(sel).to(torch.uint8).contiguous()
        fold = htype == 'polar3'
        edges = np.linspace(vmin, 90 if fold else vmax, nbins + 1)
        counts = torch.zeros(nbins, dtype=torch.int64, device=self.device)
        self._ck(self.lib.pb_histo(self._ctx, v.data_ptr(), int(v.dtype == torch.float64), _ptr(s), v.numel(),
                                   int(name in ('Rho', 'Rho_angle')), float(rotation), int(fold), _ptr(edges), nbins,
                                   counts.data_ptr(), self._stream()), 'pb_histo')
        return counts.cpu().numpy()

    def stats(self, values, rho, sel, circular: bool, rotation_fig: float = 0.0, rewrap: bool = False, fold90: bool = False):
        """Reductions of return_vecexcel (PP:2744-2779) for one variable: dict(N, mean, std, mean_delta)."""
        v = self._to_device(values)
        r = None if rho is None else self._to_device(rho).to(v.dtype).contiguous()
        s = None if sel is None else self._to_device(sel).to(torch.uint8).contiguous()
        out = np.zeros(5)
        code = int(circular) | (int(rewrap) << 1)
        self._ck(self.lib.pb_stats(self._ctx, v.data_ptr(), _ptr(r), int(v.dtype == torch.float64), _ptr(s), v.numel(), code,
                                   float(rotation_fig), int(fold90), _ptr(out), self._stream()), 'pb_stats')
        return dict(N=int(out[0]), mean=out[1], std=out[2], mean_delta=out[3])

    def roi_stats(self, res: Result, p: AnalysisParams, roi_index: Optional[int] = None, group: int = 0) -> Dict[str, float]:
        """The numeric columns of return_vecexcel (PP:2739-2779) for one ROI (or the whole image)."""
        sel = None
        if res.roi_map is not None:
            gb = np.uint32(1 << group) if roi_index is not None else np.uint32(0xffffffff)
            sel = ((res.roi_map & gb) != 0).astype(np.uint8)
        rho = res.vars[0].values
        st = self.stats(rho, rho, sel, True, float(p.rotation[1]), rewrap=True)
        out = {'N': st['N'], 'MeanRho': st['mean'], 'StdRho': st['std'], 'MeanDeltaRho': st['mean_delta']}
        for v in res.vars[1:]:
            if v.name in ('Rho_contour', 'Rho_angle'):
                for tag, fold in (('180', False), ('90', True)):
                    s2 = self.stats(v.values, rho, sel, True, 0.0, rewrap=False, fold90=fold)
                    out[f'Mean{v.name}_{tag}'], out[f'Std{v.name}_{tag}'], out[f'MeanDelta{v.name}_{tag}'] = s2['mean'], s2['std'], s2['mean_delta']
            else:
                s2 = self.stats(v.values, rho, sel, False)
                out['Mean' + v.name], out['Std' + v.name] = s2['mean'], s2['std']
        s3 = self.stats(res.intmap, rho, sel, False)
        out['MeanInt'], out['StdInt'] = s3['mean'], s3['std']
        return out

    def batch_stats_accumulate(self, o: dict, p: AnalysisParams, reduce_fn=None, per_stack: bool = False,
                               acc: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """The moment sums of return_vecexcel (PP:2739-2779) over EVERY map of the batch `o` (the dict analyze_batch returned):
        two launches (pb_stats_batch pass 1 / pass 2), nothing synchronises, the means are finalised on the device. `reduce_fn`
        (e.g. batch.allreduce_sum_) is applied to each accumulator right after its pass: that is the whole exchange of a batch
        sharded over several GPUs. Returns device tensors acc1 [groups, PB_MAX_VARS+1, 4], acc2 [groups, PB_MAX_VARS+1, 2]
        (with per_stack: one leading [B] axis more, a set per stack); batch.finalize_stats turns them into N / mean / std per group
        and variable. `acc` = caller-provided zeroed accumulators (e.g. slices of a larger buffer)."""
        names = VAR_NAMES[p.method]
        B, H, W = o['intmap'].shape
        ng = len(o['roi_indices'])
        N = o['_keep'][0].shape[1]
        pp = self._params(p, self._dtype_name(o['_keep'][0]), N, H, W)
        lead = (B,) if per_stack else ()
        if acc is not None:
            acc1, acc2 = acc
        else:
            acc1 = torch.zeros(lead + (ng, _cabi.PB_MAX_VARS + 1, 4), dtype=torch.float64, device=self.device)
            acc2 = torch.zeros(lead + (ng, _cabi.PB_MAX_VARS + 1, 2), dtype=torch.float64, device=self.device)
        outs = (PbOutputs * B)()
        for b in range(B):
            for k, n in enumerate(names):
                outs[b].var[k] = o['vars'][n][b].data_ptr()
            outs[b].intmap = o['intmap'][b].data_ptr()
        roi_dev = o['_keep'][1]
        for npass, acc_ in ((1, acc1), (2, acc2)):
            self._ck(self.lib.pb_stats_batch(self._ctx, C.byref(pp), npass, int(per_stack), outs, B, _ptr(roi_dev), acc1.data_ptr(), acc2.data_ptr(),
                                             self._stream()), 'pb_stats_batch')
            if reduce_fn is not None:
                reduce_fn(acc_)
        return acc1, acc2

    # ---- "next" row N4: the result writers' boolean-mask gathers, on the device ---------------------------
    def compact(self, values, filter=None, sel=None, roi_bits=None, group_bits: int = 0xffffffff, want: Sequence[str] = ('f16',)) -> dict:
        """`values[keep]` in row-major order with keep = sel & (roi_bits & group_bits) & isfinite(filter or values): dict with
        any of 'f16' (numpy's astype(float16) of the fp64 value), 'val', 'index' (linear pixel index), and 'count'."""
        v = self._to_device(values)
        if v.dtype not in (torch.float32, torch.float64):
            v = v.to(torch.float64)
        v = v.contiguous()
        f = None if filter is None else self._to_device(filter).to(v.dtype).contiguous()
        s = None if sel is None else self._to_device(sel).to(torch.uint8).contiguous()
        b = None if roi_bits is None else self._to_device(np.ascontiguousarray(roi_bits, dtype=np.uint32).view(np.int32)).contiguous()
        n = v.numel()
        bufs = {'f16': torch.empty(n, dtype=torch.float16, device=self.device) if 'f16' in want else None,
                'val': torch.empty(n, dtype=v.dtype, device=self.device) if 'val' in want else None,
                'index': torch.empty(n, dtype=torch.int32, device=self.device) if 'index' in want else None}
        cnt = C.c_long(0)
        self._ck(self.lib.pb_compact(self._ctx, v.data_ptr(), _ptr(f), int(v.dtype == torch.float64), _ptr(s), _ptr(b), int(group_bits) & 0xffffffff,
                                     n, _ptr(bufs['f16']), _ptr(bufs['val']), _ptr(bufs['index']), C.byref(cnt), self._stream()), 'pb_compact')
        out = {k: t[:cnt.value] for k, t in bufs.items() if t is not None}
        out['count'] = int(cnt.value)
        return out

    @staticmethod
    def _group_bits(res: Result, roi_index: Optional[int], group: int) -> int:
        # get_slice_roi (PP:2894-2904): one ROI's slice, or any ROI when roi_index is None
        return (1 << group) if roi_index is not None else 0xffffffff

    def mat_arrays(self, res: Result, roi_index: Optional[int] = None, group: int = 0) -> Dict[str, np.ndarray]:
        """The arrays save_mat stores (PP:2829-2842): X, Y (uint16 index maps, 0 outside the mask) and, per variable,
        `var.values[roi & isfinite(var.values)].astype(float16)`, gathered on the device."""
        H, W = res.mask.shape[-2:]
        m = res.mask.bool()
        xx = torch.arange(W, device=m.device, dtype=torch.int32).expand(H, W)
        yy = torch.arange(H, device=m.device, dtype=torch.int32)[:, None].expand(H, W)
        zero = torch.zeros((), dtype=torch.int32, device=m.device)
        d = {'X': torch.where(m, xx, zero).cpu().numpy().astype(np.uint16), 'Y': torch.where(m, yy, zero).cpu().numpy().astype(np.uint16)}
        ct = res.get_var('Rho_contour')
        if ct is not None:                               # PP:2835-2836: xct / yct are NaN outside isfinite(Rho_contour) (PP:3063)
            mc = torch.isfinite(ct.values)
            d.update({'X_contour': torch.where(mc, xx, zero).cpu().numpy().astype(np.uint16),
                      'Y_contour': torch.where(mc, yy, zero).cpu().numpy().astype(np.uint16)})
        gb = self._group_bits(res, roi_index, group)
        for var in res.vars:
            d[var.name] = self.compact(var.values, None, None, res.roi_map, gb, want=('f16',))['f16'].cpu().numpy()
        return d

    def csv_columns(self, res: Result, roi_index: Optional[int] = None, group: int = 0) -> Tuple[List[str], List[np.ndarray]]:
        """The columns save_csv writes (PP:2794-2811) for the non-contour table: X (uint16), Y (int16) and every non-contour
        variable under `roi & isfinite(vars[0])`, gathered on the device in row-major order."""
        W = res.mask.shape[-1]
        gb = self._group_bits(res, roi_index, group)
        v0 = res.vars[0].values
        first = self.compact(v0, None, None, res.roi_map, gb, want=('val', 'index'))
        idx = first['index'].cpu().numpy().astype(np.int64)
        names, cols = ['X', 'Y'], [(idx % W).astype(np.uint16), (idx // W).astype(np.int16)]
        # with edge contours the reference rebinds `filter` to the Rho_contour one before the variable loop (PP:2803-2806)
        ct = res.get_var('Rho_contour')
        filt = ct.values if ct is not None else v0
        for var in res.vars:
            if '_contour' in var.name:
                continue
            col = first['val'] if (var is res.vars[0] and ct is None) else self.compact(var.values, filt, None, res.roi_map, gb, want=('val',))['val']
            names.append(var.name)
            cols.append(col.cpu().numpy())
        return names, cols

    def csv_contour_columns(self, res: Result, roi_index: Optional[int] = None, group: int = 0) -> Tuple[List[str], List[np.ndarray]]:
        """The contour table of save_csv (PP:2803-2810, PP:2819-2826): X_ct, Y_ct and every *_contour variable under
        `roi & isfinite(Rho_contour)`."""
        ct = res.get_var('Rho_contour')
        if ct is None:
            raise ValueError('csv_contour_columns: the analysis had no edge contours')
        W = res.mask.shape[-1]
        gb = self._group_bits(res, roi_index, group)
        first = self.compact(ct.values, None, None, res.roi_map, gb, want=('val', 'index'))
        idx = first['index'].cpu().numpy().astype(np.int64)
        names, cols = ['X_ct', 'Y_ct'], [(idx % W).astype(np.uint16), (idx // W).astype(np.int16)]
        for var in res.vars:
            if '_contour' in var.name:
                col = first['val'] if var is ct else self.compact(var.values, ct.values, None, res.roi_map, gb, want=('val',))['val']
                names.append(var.name)
                cols.append(col.cpu().numpy())
        return names, cols

    def save_mat(self, res: Result, path, method: str, source: str = '', roi_index: Optional[int] = None, group: int = 0) -> None:
        """save_mat (PP:2829-2842) on device-resident results."""
        from datetime import date
        from scipy.io import savemat
        d = {'polarimetry': method, 'file': source, 'date': date.today().strftime('%B %d, %Y')}
        d.update(self.mat_arrays(res, roi_index, group))
        savemat(str(path), d)

    def save_csv(self, res: Result, path, method: str, source: str = '', roi_index: Optional[int] = None, group: int = 0) -> None:
        """save_csv (PP:2794-2818, non-contour table) on device-resident results; floats as '%.3f' (PP:2791-2792)."""
        import csv
        from datetime import date
        tables = [(Path(path), self.csv_columns(res, roi_index, group))]
        if res.get_var('Rho_contour') is not None:       # second file '<stem>_contour.csv' (PP:2819-2826)
            pth = Path(path)
            tables.append((pth.with_name(pth.stem + '_contour' + pth.suffix), self.csv_contour_columns(res, roi_index, group)))
        for pth, (names, cols) in tables:
            fmt = [[f'{x:.3f}' if isinstance(x, float) else x for x in (c.tolist() if c.dtype.kind == 'f' else list(c))] for c in cols]
            with open(pth, 'w', newline='') as f:
                w = csv.writer(f, delimiter=',', quoting=csv.QUOTE_MINIMAL)
                w.writerow([f"{method}, file: {source}, date: {date.today().strftime('%B %d %Y')}"])
                w.writerow(names)
                w.writerows(zip(*fmt))

    def launch_count(self) -> int:
        return int(self.lib.pb_launch_count(self._ctx))


# ---- drop-in functions with the reference's signatures ---------------------------------------------------
_default_engine: Optional[Engine] = None


def default_engine() -> Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine()
    return _default_engine


def get_intensity(self, dark: float = 0.0, bin: List[int] = [1, 1]) -> np.ndarray:
    """Drop-in for Stack.get_intensity (PC:260-264); `self` is a reference Stack."""
    return default_engine().get_intensity(self.values, dark, [int(b) for b in bin]).cpu().numpy()


def compute_dark(self, size_cell: int = 20) -> float:
    """Drop-in for Stack.compute_dark (PC:266-278); `self` is a reference Stack (integer stacks, as read from the TIFF)."""
    return default_engine().compute_dark(self.values, int(size_cell))


def binning(self, field: np.ndarray) -> np.ndarray:
    """Drop-in for Polarimetry.binning (PP:2942-2946); `self` is the reference application object."""
    bins = [int(sb.get()) for sb in self.bin_spinboxes]
    if any(s > 1 for s in bins):
        return default_engine().binning(field, bins).cpu().numpy()
    return field


def _pb_calibration(calibration, method: str):
    """The reference's Calibration object (PC:411-464) -> this package's Calibration, cached on the object."""
    if method == '1PF':
        cal = getattr(calibration, '_pb_calib', None)
        if cal is None:
            cal = Calibration(calibration.RhoPsi[:, :, 0], calibration.RhoPsi[:, :, 1], calibration.xy[0].size, name=getattr(calibration, 'name', ''))
            try:
                calibration._pb_calib = cal
            except Exception:
                pass
        return cal
    if method.startswith('4POLAR'):
        return Calibration(invK=calibration.invKmat)
    return None


def compute_fields(field: np.ndarray, datastack, mask: np.ndarray, method: str = '1PF', polar_dir: str = 'clockwise',
                   offset_angle: float = 0, calibration=None, edge_contours=None, reference_angle: float = 0,
                   rotation: float = 0, for_calib: bool = False, chi2threshold: float = 500, _Variable=None):
    """Drop-in for the module-level compute_fields (PP:2981-3078): same arguments, same side effects
    (`mask` updated in place; `datastack.vars/.intmap/.xmap/.ymap/.field/.field_fit/.chi2` filled)."""
    eng = default_engine()
    p = AnalysisParams(method=method, polar_dir=polar_dir, offset_angle=float(offset_angle),
                       rotation=(float(rotation), 0.0, float(reference_angle)), chi2threshold=float(chi2threshold))
    cal = _pb_calibration(calibration, method)
    res = eng.compute_fields(field, mask, p, cal, edge_contours, want_fit=not for_calib)
    mk = (lambda n, v: _Variable(n, values=v)) if _Variable is not None else Variable
    datastack.vars = [mk(v.name, v.values.cpu().numpy()) for v in res.vars]
    final = res.mask.cpu().numpy().astype(bool)
    mask[...] = final
    datastack.intmap = res.intmap.cpu().numpy()
    if for_calib:
        return datastack
    datastack.xmap, datastack.ymap = res.xmap.cpu().numpy(), res.ymap.cpu().numpy()
    if edge_contours is not None:
        both = np.isfinite(datastack.vars[0].values) & np.isfinite(edge_contours)
        datastack.xct[~both], datastack.yct[~both] = np.nan, np.nan
    if not method.startswith('4POLAR'):
        field[:, ~final] = np.nan
        datastack.field, datastack.field_fit, datastack.chi2 = field, res.field_fit.cpu().numpy(), res.chi2.cpu().numpy()
    return datastack


class LazyMaps:
    """`datastack.field / .field_fit / .chi2` of the fused S4 path (PP:3073-3077). The reference materialises two float64 [N,H,W]
    arrays that only the individual-fit viewer reads, one clicked pixel at a time (PP:2485-2494); here they are produced on first
    use from the raw stack (pb_field + pb_fit_maps) and then behave like the NumPy arrays the reference stores."""

    def __init__(self, engine: Engine, values, params: AnalysisParams, mask: torch.Tensor):
        self._eng, self._values, self._p, self._mask, self._cache = engine, values, params, mask, None

    def _get(self):
        if self._cache is None:
            f = self._eng.field(self._values, self._p)
            N, H, W = f.shape
            self._eng._set_basis(N, self._p)
            pp = self._eng._params(self._p, 'float64', N, H, W)
            fit = torch.empty_like(f)
            chi2 = torch.empty((H, W), dtype=torch.float64, device=f.device)
            m = self._mask.to(torch.uint8).contiguous()
            self._eng._ck(self._eng.lib.pb_fit_maps(self._eng._ctx, C.byref(pp), f.data_ptr(), m.data_ptr(), fit.data_ptr(), chi2.data_ptr(),
                                                    self._eng._stream()), 'pb_fit_maps')
            f = torch.where(m.bool()[None], f, torch.full_like(f, float('nan')))          # PP:3074
            self._cache = {'field': f.cpu().numpy(), 'field_fit': fit.cpu().numpy(), 'chi2': chi2.cpu().numpy()}
            self._values = None
        return self._cache

    def view(self, key: str):
        return _LazyArray(self, key)


class _LazyArray:
    """NumPy-like view on one of the LazyMaps arrays: indexing, np.asarray(), .shape / .ndim / .dtype trigger the computation."""

    def __init__(self, owner: LazyMaps, key: str):
        self._owner, self._key = owner, key

    def _a(self) -> np.ndarray:
        return self._owner._get()[self._key]

    def __array__(self, dtype=None, copy=None):
        a = self._a()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, idx):
        return self._a()[idx]

    def __setitem__(self, idx, value):
        self._a()[idx] = value

    def __len__(self):
        return len(self._a())

    def __getattr__(self, name):
        return getattr(self._a(), name)


def analyze_stack(self, datastack, for_calib: bool = False, _Variable=None):
    """Drop-in for Polarimetry.analyze_stack (PP:2948-2979), the FUSED seam S4: the GUI state is pulled exactly where the reference
    pulls it (PP:2949-2968), the raw `self.stack.values` goes through the fused kernels (one pass over the stack: pre-processing,
    binning, compute_fields), `datastack` is filled as compute_fields fills it, and the reference's own `return_vecexcel` /
    `plot_data` / `save_data` run on the result. `self` is the reference application object."""
    eng = default_engine()
    roi_map, _mask = self.compute_roi_map(datastack)                       # PP:2949 (also the per_roi.deselect() side effect, PP:2873-2874)
    method = self.method.get()                                             # PP:2950
    datastack.dark = float(self.dark.get())                                # PP:2951
    datastack.method = method                                              # PP:2952
    calibration = self.CD if method in ['1PF', '4POLAR 2D', '4POLAR 3D'] else None      # PP:2958-2961
    if hasattr(self, 'edge_contours'):                                     # PP:2963-2966
        edge_contours, datastack.xct, datastack.yct = self.define_rho_ct(self.edge_contours)
    else:
        edge_contours = None
    p = AnalysisParams(method=method, dark=datastack.dark, removed_intensity=float(self.removed_intensity),
                       bins=tuple(int(sb.get()) for sb in self.bin_spinboxes),                     # PP:2943
                       polar_dir=self.polar_dir.get(), offset_angle=float(self.offset_angle.get()),
                       rotation=(float(self.rotation[0].get()), 0.0, float(self.rotation[2].get())),
                       ilow=float(self.ilow.get()))                                                # PP:2876
    base = self.get_mask(datastack, False)                                 # PP:2890 (the PNG mask or ones; no second pop-up)
    res = eng.analyze(self.stack.values, p, _pb_calibration(calibration, method), datastack.rois, bool(self.per_roi.get()),
                      None if np.all(np.asarray(base) != 0) else base, edge_contours, want_hist=False)
    mk = (lambda n, v: _Variable(n, values=v)) if _Variable is not None else Variable
    datastack.vars = [mk(v.name, v.values.cpu().numpy()) for v in res.vars]
    datastack.intmap = res.intmap.cpu().numpy()
    if not for_calib:
        datastack.xmap, datastack.ymap = res.xmap.cpu().numpy(), res.ymap.cpu().numpy()       # PP:3056-3058
        if edge_contours is not None:                                                          # PP:3063
            both = np.isfinite(datastack.vars[0].values) & np.isfinite(edge_contours)
            datastack.xct[~both], datastack.yct[~both] = np.nan, np.nan
        if not method.startswith('4POLAR'):                                                    # PP:3073-3077, on demand
            lazy = LazyMaps(eng, self.stack.values, p, res.mask)
            datastack.field, datastack.field_fit, datastack.chi2 = lazy.view('field'), lazy.view('field_fit'), lazy.view('chi2')
    if for_calib:                                                          # PP:2969-2977
        if self.per_roi.get():
            result = []
            for roi in datastack.rois:
                if roi['select']:
                    result.append(self.return_vecexcel(datastack, roi_map, roi=roi, simplify=True)[0])
        else:
            result = [self.return_vecexcel(datastack, roi_map, roi=[], simplify=True)[0]]
        return result
    self.plot_data(datastack, roi_map=roi_map)                             # PP:2978
    self.save_data(datastack, roi_map=roi_map)                             # PP:2979


def histogram(a, bins=10, range=None, *args, **kwargs):
    """Drop-in for the np.histogram call of Variable.histo (PC:369, seam S5): explicit uniform edges (np.linspace) and a float
    array go to the device (pb_histo: [e_k, e_k+1), last bin closed, values outside dropped -- numpy's rule); anything else is
    numpy's business. Returns (counts int64, bin_edges) like numpy."""
    edges = np.asarray(bins)
    arr = np.asarray(a)
    if args or kwargs or edges.ndim != 1 or edges.size < 2 or edges.size - 1 > _cabi.PB_MAX_BINS or arr.dtype.kind != 'f' or arr.size == 0:
        return np.histogram(a, bins, range, *args, **kwargs)
    edges = np.ascontiguousarray(edges, dtype=np.float64)
    step = (edges[-1] - edges[0]) / (edges.size - 1)
    if not (step > 0) or np.any(np.abs(edges - (edges[0] + step * np.arange(edges.size))) > 0.25 * step):
        return np.histogram(a, bins, range, *args, **kwargs)
    eng = default_engine()
    v = eng._to_device(np.ascontiguousarray(arr.ravel(), dtype=np.float64))
    counts = torch.zeros(edges.size - 1, dtype=torch.int64, device=eng.device)
    eng._ck(eng.lib.pb_histo(eng._ctx, v.data_ptr(), 1, None, v.numel(), 0, 0.0, 0, _ptr(edges), edges.size - 1, counts.data_ptr(), eng._stream()),
            'pb_histo')
    return counts.cpu().numpy(), edges


class _NumpyWithDeviceHistogram:
    """Stands in for the `np` name of pypolar_classes: every attribute is numpy's, except `histogram` (seam S5)."""

    def __init__(self, numpy_module):
        self.__dict__['_np'] = numpy_module

    def __getattr__(self, name):
        return histogram if name == 'histogram' else getattr(self.__dict__['_np'], name)


_SEAMS = (('module', 'compute_fields'), ('Polarimetry', 'binning'), ('Polarimetry', 'analyze_stack'), ('Stack', 'get_intensity'),
          ('Stack', 'compute_dark'), ('classes', 'np'))


def install(pypolar_module, classes_module=None) -> dict:
    """Rebind the hot-path seams of an imported reference (`import PyPOLAR, pypolar_classes`) to this package:
    S1 compute_fields, S2 Polarimetry.binning, S3 Stack.get_intensity (+ Stack.compute_dark), S4 Polarimetry.analyze_stack (the
    fused entry: what the Analysis button, the folder loop and the calibration sweep call, PP:898 / PP:2128 / PP:2133) and S5 the
    np.histogram call of Variable.histo. Nothing touches the GPU until a rebound function is called. Returns the original
    attributes for `uninstall`."""
    import functools
    import inspect
    Var = getattr(pypolar_module, 'Variable', None) or getattr(classes_module, 'Variable', None)
    stack_cls = getattr(classes_module, 'Stack', None) or getattr(pypolar_module, 'Stack', None)
    saved = {'pypolar': pypolar_module, 'classes': classes_module, 'stack_cls': stack_cls, 'attrs': {}}

    def rebind(owner, name, new, key):
        old = getattr(owner, name, None)
        saved['attrs'][key] = (owner, name, old)
        if callable(old) and callable(new):
            try:                                          # present the reference's own signature / name to introspection
                new.__signature__ = inspect.signature(old)
                new.__name__, new.__qualname__ = old.__name__, getattr(old, '__qualname__', old.__name__)
            except (TypeError, ValueError):
                pass
        setattr(owner, name, new)

    def _cf(*a, **k):
        return compute_fields(*a, _Variable=Var, **k)

    def _as(self, datastack, for_calib=False):
        return analyze_stack(self, datastack, for_calib, _Variable=Var)

    def _binning(self, field):
        return binning(self, field)

    def _gi(self, dark=0.0, bin=[1, 1]):
        return get_intensity(self, dark, bin)

    def _cd(self, size_cell=20):
        return compute_dark(self, size_cell)

    functools.update_wrapper(_cf, compute_fields)
    rebind(pypolar_module, 'compute_fields', _cf, 'S1')
    rebind(pypolar_module.Polarimetry, 'binning', _binning, 'S2')
    rebind(pypolar_module.Polarimetry, 'analyze_stack', _as, 'S4')
    if stack_cls is not None:
        rebind(stack_cls, 'get_intensity', _gi, 'S3')
        rebind(stack_cls, 'compute_dark', _cd, 'N2')
    if classes_module is not None and hasattr(classes_module, 'np'):
        rebind(classes_module, 'np', _NumpyWithDeviceHistogram(classes_module.np), 'S5')
    return saved


def uninstall(saved: dict) -> None:
    """Undo `install` (restores the reference's own functions)."""
    for owner, name, old in saved['attrs'].values():
        if old is None:
            try:
                delattr(owner, name)
            except AttributeError:
                pass
        else:
            setattr(owner, name, old)
