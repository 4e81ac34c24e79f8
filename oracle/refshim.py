"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* Python reference.

Imports the reference's ``tools``/``pack``/``generate``/``model`` modules so that (a) the C restatement in
``oracle/tap_oracle.c`` can be validated against the real thing, (b) golden vectors can be generated
(``tests/golden/make_golden.py``), (c) the model-in-the-loop GPU tests and bench.py's reference arm can run the real
``model.DRL`` / ``tools.Container``.  Where the modules come from: ``$TAPNET_REFERENCE`` if set, else /root/reference (the
build container), else oracle/_ref/ -- the byte-for-byte staged copy ``oracle/stage_ref.py`` makes (git-ignored; it is what
travels to the GPU box, where /root/reference does not exist).  Nothing under tap-net_b200/ imports this.

Three shims are needed:
  1. ``numpy.math`` was removed in NumPy 2 (used at pack.py:109,309,365).
  2. matplotlib / mpl_toolkits are not installed: stub modules.  The only stub
     with behaviour is ``matplotlib.path.Path.contains_point`` (tools.py:764),
     restated from matplotlib's src/_path.h ``point_in_path_impl`` -- this is the
     one UNPINNED point of the 3D parity.
  3. the reference writes datasets relative to cwd -> callers chdir to a scratch dir.
"""
import math
import os
import sys
import types

STAGED_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _resolve():
    env = os.environ.get("TAPNET_REFERENCE")
    if env:
        return env
    for cand in ("/root/reference", STAGED_DIR):
        if os.path.isfile(os.path.join(cand, "tools.py")):
            return cand
    return "/root/reference"


REFERENCE_DIR = _resolve()


def staged():
    """True when the modules come from the staged copy (oracle/_ref/), i.e. on the GPU box."""
    return os.path.abspath(REFERENCE_DIR) == os.path.abspath(STAGED_DIR)


def checkpoint(name):
    """Path of a shipped pretrained actor, e.g. '2d-bot-C+P+S-lb-soft-width-5-note-sh-R-diff'."""
    return os.path.join(REFERENCE_DIR, "pretrain_model", name, "actor.pt")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "tools.py"))


class _Path(object):
    """Restatement of matplotlib.path.Path for straight, code-less polygons.

    ``contains_point`` follows matplotlib src/_path.h point_in_path_impl (crossing
    test with the ``vy >= ty`` half-open rule, polygon implicitly closed), radius 0,
    no transform."""

    def __init__(self, vertices, codes=None):
        self.vertices = [(float(v[0]), float(v[1])) for v in vertices]

    def contains_point(self, point, transform=None, radius=0.0):
        tx, ty = float(point[0]), float(point[1])
        v = self.vertices
        n = len(v)
        if n < 3:
            return False
        inside = False
        sx, sy = v[0]
        vtx0, vty0 = sx, sy
        yflag0 = vty0 >= ty
        for k in range(1, n + 1):
            vtx1, vty1 = v[k] if k < n else (sx, sy)
            yflag1 = vty1 >= ty
            if yflag0 != yflag1:
                if ((vty1 - ty) * (vtx0 - vtx1) >= (vtx1 - tx) * (vty0 - vty1)) == yflag1:
                    inside = not inside
            yflag0 = yflag1
            vtx0, vty0 = vtx1, vty1
        return inside


def _install_stubs():
    import numpy as np
    if not hasattr(np, "math"):
        np.math = math
    if "matplotlib" in sys.modules and not getattr(sys.modules["matplotlib"], "_tap_stub", False):
        return  # a real matplotlib is present: use it

    def mod(name):
        m = types.ModuleType(name)
        m._tap_stub = True
        sys.modules[name] = m
        return m

    mpl = mod("matplotlib")
    mpl.use = lambda *a, **k: None
    mpl.rcParams = {}
    mpl.rc = lambda *a, **k: None
    pyplot = mod("matplotlib.pyplot")

    def _noop_attr(name):                                # trainer.train's plt.close / title / plot / savefig: no-ops
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None
    pyplot.__getattr__ = _noop_attr
    patches = mod("matplotlib.patches")
    path = mod("matplotlib.path")
    path.Path = _Path
    mpl.pyplot, mpl.patches, mpl.path = pyplot, patches, path
    tk = mod("mpl_toolkits")
    m3 = mod("mpl_toolkits.mplot3d")
    m3.Axes3D = type("Axes3D", (), {})
    art = mod("mpl_toolkits.mplot3d.art3d")
    art.Poly3DCollection = type("Poly3DCollection", (), {})
    ax3 = mod("mpl_toolkits.mplot3d.axis3d")
    ax3.Axis = type("Axis", (), {"_get_coord_info": lambda self, renderer: None})
    tk.mplot3d = m3
    m3.art3d, m3.axis3d = art, ax3


_cache = {}


def load(names=("tools", "pack", "generate")):
    """Return the reference modules (dict name -> module)."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_DIR)
    _install_stubs()
    if REFERENCE_DIR not in sys.path:
        sys.path.append(REFERENCE_DIR)
    out = {}
    for n in names:
        if n not in _cache:
            # never shadowed by our drop-in 'pack'/'tools': load by file path
            import importlib.util
            spec = importlib.util.spec_from_file_location(n, os.path.join(REFERENCE_DIR, n + ".py"))
            m = importlib.util.module_from_spec(spec)
            prev = sys.modules.get(n)
            sys.modules[n] = m
            try:
                spec.loader.exec_module(m)
            except Exception:
                if prev is not None:
                    sys.modules[n] = prev
                else:
                    sys.modules.pop(n, None)
                raise
            _cache[n] = m
        out[n] = _cache[n]
    return out
