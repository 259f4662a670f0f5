"""TEST INFRASTRUCTURE ONLY -- headless loader for the *unmodified* reference (PyPOLAR v2.9.4).

The reference is a Tk GUI; its numerical hot path (PyPOLAR.py:2942-3078, pypolar_classes.py:260-264,
339-369) is plain NumPy/SciPy. This module imports the reference's two source files from where they lie
(`/root/reference/pypolar`, read-only, never copied) with every GUI / plotting / file-format dependency
replaced by a permissive stub, builds a fake `Polarimetry` application object holding only the state the
hot path reads, and exposes `run_reference()` which calls the reference's own
`compute_intensity` -> `analyze_stack(for_calib=True)` / `compute_fields` and the `np.histogram` call of
`Variable.histo` (PC:368-369).

It is used (a) by `oracle/make_golden.py` to generate the committed fixtures under `tests/golden/` and
(b) by the CPU tests that pin `oracle/restate.py` against the reference. It cannot travel to the GPU box
(`/root/reference` does not exist there): nothing run under `-m gpu`, `smoke()` or `bench.py` imports it.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

import numpy as np

REFERENCE_DIRS = [Path('/root/reference/pypolar'), Path(__file__).resolve().parent.parent / 'baseline' / '_ref' / 'pypolar']

_STUBBED = [
    'customtkinter', 'tkinter', 'tkinter.filedialog', 'tkinter.messagebox', 'tksheet', 'roifile',
    'matplotlib', 'matplotlib.pyplot', 'matplotlib.figure', 'matplotlib.patches', 'matplotlib.collections',
    'matplotlib.backend_bases', 'matplotlib.backends', 'matplotlib.backends.backend_tkagg',
    'matplotlib.backends.backend_pdf', 'matplotlib.ticker', 'matplotlib._pylab_helpers',
    'matplotlib.font_manager', 'matplotlib.image', 'matplotlib.colors', 'mpl_toolkits', 'mpl_toolkits.axes_grid1',
    'colorcet', 'skimage', 'skimage.exposure', 'skimage.measure', 'tifffile', 'openpyxl', 'requests',
    'generate_json', 'joblib', 'PIL', 'PIL.Image', 'PIL.ImageTk',
]


class _Meta(type):
    """Metaclass so that *class-level* attribute access on a dummy (e.g. annotations such as
    `mpl.image.AxesImage`, PC:329) yields another dummy instead of raising."""

    def __getattr__(cls, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _make_dummy(name)


def _make_dummy(name: str):
    # NOTE: no instance-level __getattr__: the reference relies on hasattr(self, 'edge_contours') (PP:2963).
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return None

    return _Meta(name, (), {'__init__': __init__, '__call__': __call__})


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        d = _make_dummy(name)
        setattr(self, name, d)
        return d


def _install_stubs() -> None:
    for full in _STUBBED:
        if full in sys.modules and not isinstance(sys.modules[full], _StubModule):
            # a real module of that name is importable here (e.g. joblib, PIL): keep the real one
            continue
        try:
            if full.split('.')[0] in ('joblib', 'PIL', 'requests'):
                importlib.import_module(full)
                continue
        except Exception:
            pass
        m = _StubModule(full)
        m.__path__ = []  # behave as a package
        sys.modules[full] = m
    for full in _STUBBED:
        if '.' in full and isinstance(sys.modules.get(full), _StubModule):
            parent, child = full.rsplit('.', 1)
            if isinstance(sys.modules.get(parent), _StubModule):
                setattr(sys.modules[parent], child, sys.modules[full])
    gj = sys.modules['generate_json']
    # constants imported at PC:26 / PP:53 (the real module writes ~/.pypolar at import: stubbed)
    for k, v in dict(orange=('#FF7F4F', '#ffb295'), gray=('#7F7F7F', '#A6A6A6'), red=('#B54D42', '#d5928b'),
                     green=('#ADD1AD', '#cee3ce'), blue=('#0047AB', '#0047AB'), text_color='black',
                     header_fontsize=18, default_fontsize=14, default_fontname='Nunito', tooltip_fontsize=13,
                     os_name='Linux').items():
        setattr(gj, k, v)
    mpl = sys.modules['matplotlib']
    mpl.use = lambda *a, **k: None
    plt = sys.modules['matplotlib.pyplot']
    plt.rcParams = {}
    plt.ion = lambda *a, **k: None
    ctk = sys.modules['customtkinter']
    for fn in ('deactivate_automatic_dpi_awareness', 'set_appearance_mode', 'set_window_scaling', 'set_widget_scaling'):
        setattr(ctk, fn, lambda *a, **k: None)
    # class bases used by the reference must be plain classes
    ctk.CTk = type('CTk', (), {})
    ctk.CTkToplevel = type('CTkToplevel', (), {})
    ctk.CTkFrame = type('CTkFrame', (), {})
    ctk.CTkTabview = type('CTkTabview', (), {})
    tk = sys.modules['tkinter']
    tk.Frame = type('Frame', (), {})
    tk.Event = type('Event', (), {})


_cache = {}


def reference_dir() -> Path:
    for d in REFERENCE_DIRS:
        if (d / 'PyPOLAR.py').exists():
            return d
    raise FileNotFoundError('reference not available (looked in %s)' % ', '.join(map(str, REFERENCE_DIRS)))


def available() -> bool:
    try:
        reference_dir()
        return True
    except FileNotFoundError:
        return False


def load():
    """Import the reference modules (unmodified) and return (PyPOLAR, pypolar_classes)."""
    if 'mods' in _cache:
        return _cache['mods']
    d = reference_dir()
    _install_stubs()
    sys.path.insert(0, str(d))
    try:
        pc = importlib.import_module('pypolar_classes')
        pp = importlib.import_module('PyPOLAR')
    finally:
        sys.path.remove(str(d))
    _cache['mods'] = (pp, pc)
    return pp, pc


class _Var:
    """Stand-in for a Tk StringVar / CheckBox / SpinBox: .get/.set/.select/.deselect."""

    def __init__(self, v):
        self.v = v

    def get(self):
        return self.v

    def set(self, v):
        self.v = v

    def select(self):
        self.v = True

    def deselect(self):
        self.v = False


def make_app(method='1PF', dark=0.0, ilow=0.0, offset_angle=0.0, polar_dir='clockwise', bins=(1, 1),
             rotation=(0.0, 0.0, 0.0), removed_intensity=0, disk=None, calib_label='no distortions',
             per_roi=False, option='Thresholding'):
    """Fake `Polarimetry` instance carrying exactly the state read by PP:2869-2979."""
    pp, pc = load()
    app = pp.Polarimetry.__new__(pp.Polarimetry)
    app.method = _Var(method)
    app.dark = _Var(repr(float(dark)))
    app.ilow = _Var(repr(float(ilow)))
    app.offset_angle = _Var(repr(float(offset_angle)))
    app.polar_dir = _Var(polar_dir)
    app.per_roi = _Var(bool(per_roi))
    app.option = _Var(option)
    app.polar_dropdown = _Var('UL90-UR0-LR135-LL45')
    app.bin_spinboxes = [_Var(int(bins[0])), _Var(int(bins[1]))]
    app.rotation = [_Var(repr(float(r))) for r in rotation]
    app.removed_intensity = removed_intensity
    app.beads_name, app.contrastThreshold, app.sigma = '', 0, 0
    if method == '1PF':
        cd = pc.Calibration.__new__(pc.Calibration)
        disk_file = Path(disk) if disk is not None else reference_dir() / 'diskcones' / 'Disk_Ga0_Pa0_Ta0_Gb0_Pb0_Tb0_Gc0_Pc0_Tc0.mat'
        cd.define_disk(disk_file)
        cd.offset_default = 0
        app.CD = cd
    elif method.startswith('4POLAR'):
        app.CD = pc.Calibration(method, calib_label)
    else:
        app.CD = None
    return app


def run_reference(values: np.ndarray, method='1PF', dark=0.0, ilow=0.0, offset_angle=0.0,
                  polar_dir='clockwise', bins=(1, 1), rotation=(0.0, 0.0, 0.0), removed_intensity=0,
                  disk=None, calib_label='no distortions', rois=None, per_roi=False, for_calib=True,
                  reference_angle=None, edge_contours=None, nbins=60, var_ranges=None, mask_image=None):
    """Run the reference's own code on one stack `values` [N,H,W]. Returns a dict of results.

    for_calib=True goes through `Polarimetry.analyze_stack` (PP:2948) and returns the stats rows;
    for_calib=False replays analyze_stack's statements PP:2949-2968 and calls the reference's
    `compute_fields` directly (plot/save are GUI-only), returning the non-calib outputs as well.
    """
    pp, pc = load()
    if reference_angle is not None:
        rotation = (rotation[0], rotation[1], reference_angle)
    app = make_app(method, dark, ilow, offset_angle, polar_dir, bins, rotation, removed_intensity, disk,
                   calib_label, per_roi, 'ROI' if rois else 'Thresholding')
    stack = pc.Stack(Path('/tmp/synthetic.tif'))
    if mask_image is not None:
        # option 'Mask' (PP:2084): the reference opens <maskfolder>/<stack name>.png with PIL and divides by its maximum
        import tempfile
        import cv2
        folder = Path(tempfile.mkdtemp(prefix='pb_mask_'))
        cv2.imwrite(str(folder / 'synthetic.png'), np.asarray(mask_image, dtype=np.uint8))
        app.option = _Var('Mask')
        app.maskfolder = folder
    stack.values = values
    stack.nangle, stack.height, stack.width = values.shape
    app.stack = stack
    pp.Polarimetry.compute_intensity(app, stack)          # PP:2242 -> PC:260
    ds = pc.DataStack(stack)
    ds.rois = rois or []
    out = {}
    if for_calib:
        out['rows'] = pp.Polarimetry.analyze_stack(app, ds, for_calib=True)
        roi_map, _ = pp.Polarimetry.compute_roi_map(app, ds)
    else:
        roi_map, mask = pp.Polarimetry.compute_roi_map(app, ds)
        ds.dark = float(app.dark.get())
        ds.method = method
        field = np.maximum(stack.values - (ds.dark + float(app.removed_intensity)), 0)
        if method == 'CARS':
            field = np.sqrt(field)
        elif method.startswith('4POLAR'):
            field[[1, 2]] = field[[2, 1]]
        calibration = app.CD if method in ['1PF', '4POLAR 2D', '4POLAR 3D'] else None
        field = pp.Polarimetry.binning(app, field)
        global_mask = np.any(mask, axis=0) if mask.ndim == 3 else mask
        if edge_contours is not None:
            ds.xct, ds.yct = np.indices(values.shape[1:], dtype=float)[::-1].copy()
        ds = pp.compute_fields(field, ds, global_mask, method=method, polar_dir=polar_dir,
                               offset_angle=float(offset_angle), calibration=calibration,
                               edge_contours=edge_contours, reference_angle=float(rotation[2]),
                               rotation=float(rotation[0]), for_calib=False)
        out['xmap'], out['ymap'] = ds.xmap, ds.ymap
        out['field'], out['field_fit'], out['chi2'] = ds.field, ds.field_fit, ds.chi2
    out['intensity'] = np.asarray(stack.intensity)
    out['intmap'] = ds.intmap
    out['vars'] = {v.name: v.values for v in ds.vars}
    out['roi_map'] = roi_map
    # histograms: the numeric part of Variable.histo (PC:340-369) with the same arguments as PP:2279
    ranges = dict(Rho=(0, 180), Rho_contour=(0, 180), Rho_angle=(0, 180), Psi=(40, 180), S2=(0, 1), S4=(-1, 1),
                  S_SHG=(-1, 1), Eta=(0, 90))
    ranges.update(var_ranges or {})
    hist = {}
    slices = [None] if not (per_roi and rois) else [r['indx'] for r in rois if r.get('select')]
    for v in ds.vars:
        if v.name not in ranges:
            continue
        for s in slices:
            m = pp.Polarimetry.get_slice_roi(app, roi_map, s)
            data_vals = v.values[m * np.isfinite(v.values)]
            if v.name in ['Rho', 'Rho_angle']:
                data_vals = np.mod(2 * (data_vals + float(rotation[1])), 360) / 2
            vmin, vmax = ranges[v.name]
            bins_ = np.linspace(vmin, vmax, nbins + 1)
            hist[(v.name, s)] = np.histogram(data_vals, bins=bins_, range=(vmin, vmax))[0]
    out['hist'] = hist
    return out
