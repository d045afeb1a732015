"""ORACLE — TEST INFRASTRUCTURE ONLY (build container only).

Imports the UNMODIFIED reference from /root/reference with the three shims of SURVEY.md §8c so that
tests/golden/make_golden.py and tests/test_oracle.py can pin oracle/pointpath_oracle.py against the
reference's own code. /root/reference does not exist on the GPU box: nothing that runs there
(`-m gpu` tests, smoke(), bench.py) may call `load()`; use `available()` to gate.

Shims:
  1. CWD/argv : modules/config/Config.py:4 opens ./config.yml and Parser.py:12 parses sys.argv at import.
  2. network  : modules/imhead/Pipe.py:8 downloads Faster-RCNN weights at import -> force weights=None.
  3. device   : config.yml:1 says 'cuda'; VoxelNet.reindex allocates on cfg.device -> set 'cpu'.
`modules.Extension` (a torch JIT build that writes under ~/.cache) is replaced by the reference's own
cpp/voxelutil.cpp compiled in place by oracle/Makefile into oracle/_ref/.
"""
from __future__ import annotations

import importlib
import importlib.util
import glob
import os
import subprocess
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
SOURCE = '/root/reference'                                   # read-only, build container only
STAGED = os.path.join(_ROOT, 'baseline', '_ref')             # git-ignored copy that travels to the GPU box
REF = os.environ.get('MVX_REFERENCE_ROOT') or (SOURCE if os.path.isfile(os.path.join(SOURCE, 'MVXNet.py')) else STAGED)
_mods = None


def stage() -> bool:
    """Copy the reference's Python tree (+ config.yml, cpp/) into the git-ignored baseline/_ref/ so that the GPU box - which
    has no /root/reference - can run the UNMODIFIED reference modules next to the drop-in (tests/test_gpu_dropin.py).
    Never part of the product or of the git history; called by __graft_entry__.build() where /root/reference exists."""
    import shutil
    if not os.path.isfile(os.path.join(SOURCE, 'MVXNet.py')):
        return False
    keep = ('.py', '.yml', '.cpp', '.md', '.txt')
    for dirpath, dirnames, files in os.walk(SOURCE):
        dirnames[:] = [d for d in dirnames if not d.startswith('.') and d != '__pycache__']
        rel = os.path.relpath(dirpath, SOURCE)
        for f in files:
            if f.endswith(keep):
                dst = os.path.join(STAGED, rel, f)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(os.path.join(dirpath, f), dst)
    return True


def available() -> bool:
    return os.path.isfile(os.path.join(REF, 'config.yml')) and os.path.isfile(os.path.join(REF, 'MVXNet.py'))


def load_voxelutil():
    """The reference's native module, compiled from /root/reference/cpp/voxelutil.cpp (oracle/_ref)."""
    if os.path.isfile(os.path.join(SOURCE, 'cpp', 'voxelutil.cpp')):      # the Makefile compiles it where it lies
        subprocess.run(['make', '-s', '-C', _HERE, 'ref'], check=True)
    so = glob.glob(os.path.join(_HERE, '_ref', 'voxelutil*.so'))
    if not so:
        raise RuntimeError('oracle/_ref/voxelutil*.so missing (needs /root/reference to build)')
    spec = importlib.util.spec_from_file_location('voxelutil', so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load(device: str = 'cpu'):
    """Returns a namespace with the reference modules: cfg, pre, calib, imhead_pipe, layers, vpipe, VoxelNet, cpp."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(f'reference tree not found at {REF}')
    cwd, argv = os.getcwd(), sys.argv
    try:
        os.chdir(REF)
        sys.argv = ['refshim']
        sys.path.insert(0, REF)
        ext = types.ModuleType('modules.Extension')
        ext.cpp = load_voxelutil()
        import torchvision.models.detection.faster_rcnn as frcnn
        orig = frcnn.fasterrcnn_resnet50_fpn_v2
        frcnn.fasterrcnn_resnet50_fpn_v2 = lambda *a, **k: orig(weights=None, weights_backbone=None)
        try:
            import modules                                   # noqa: F401  (reference package)
            sys.modules['modules.Extension'] = ext
            cfg = importlib.import_module('modules.config')
            cfg.config['device'] = device
            pre = importlib.import_module('modules.data.Preprocessing')
            calib = importlib.import_module('modules.utils.Calib')
            layers = importlib.import_module('modules.layers')
            imhead_pipe = importlib.import_module('modules.imhead.Pipe')
            vpipe = importlib.import_module('modules.voxelnet.Pipe')
            vnet = importlib.import_module('modules.voxelnet.VoxelNet')
            mvxnet = importlib.import_module('MVXNet')       # the model assembly (MVXNet.py:13-27); needs the patched Pipe import above
        finally:
            frcnn.fasterrcnn_resnet50_fpn_v2 = orig
    finally:
        os.chdir(cwd)
        sys.argv = argv
    _mods = types.SimpleNamespace(cfg=cfg, pre=pre, calib=calib, layers=layers, imhead_pipe=imhead_pipe,
                                  vpipe=vpipe, VoxelNet=vnet.VoxelNet, cpp=ext.cpp, MVXNet=mvxnet.MVXNet)
    return _mods


def load_calc():
    """modules/Calc.py (bbox3d2bev, classifyAnchors). Its `from shapely.geometry import Polygon` serves only the shapely variants
    (`iou2d`, `getPolygons`, `classifyAnchors_`), which are never called here; shapely is not installed, so an empty stand-in
    module satisfies the import."""
    load()
    if 'shapely' not in sys.modules:
        try:
            importlib.import_module('shapely.geometry')
        except ImportError:
            sh, geo = types.ModuleType('shapely'), types.ModuleType('shapely.geometry')
            geo.Polygon = object
            sh.geometry = geo
            sys.modules['shapely'], sys.modules['shapely.geometry'] = sh, geo
    return importlib.import_module('modules.Calc')
