"""Load the UNMODIFIED reference package from /root/reference for golden generation.

Test infrastructure only (never imported by the product).  The reference imports
``pastas`` and ``matplotlib`` which are not installed in this image and uses a few
numpy/pandas names that no longer exist; this module provides import shims only --
no arithmetic on the Kalman path is replaced.

Only usable in the build container: /root/reference does not exist on the GPU box.
"""
import importlib
import logging
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("METRAN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "metran", "kalmanfilter.py"))


def _install_stubs():
    import numpy as np

    # numpy names removed in numpy>=1.24 / 2.0 (kalmanfilter.py:99, solver.py:257)
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "NaN"):
        np.NaN = np.nan

    if "pastas" not in sys.modules:
        import numba

        pastas = types.ModuleType("pastas")
        pastas.__version__ = "1.4.0"
        pastas.__path__ = []

        dec = types.ModuleType("pastas.decorators")
        dec.njit = numba.njit  # reference: kalmanfilter.py:8 wraps numba.njit

        utils = types.ModuleType("pastas.utils")
        utils.initialize_logger = lambda logger=None, level=logging.INFO: None
        utils.validate_name = lambda name, raise_error=False: name

        ts = types.ModuleType("pastas.timeseries")

        class TimeSeries:  # only used in isinstance checks (metran.py:541)
            pass

        ts.TimeSeries = TimeSeries

        tsu = types.ModuleType("pastas.timeseries_utils")
        tsu._frequency_is_supported = lambda freq: freq

        ver = types.ModuleType("pastas.version")
        ver.__version__ = "1.4.0"

        plotting = types.ModuleType("pastas.plotting")
        plotting.__path__ = []
        plotutil = types.ModuleType("pastas.plotting.plotutil")
        plotutil._get_height_ratios = lambda ylims: [1.0 for _ in ylims]

        mods = {
            "pastas": pastas,
            "pastas.decorators": dec,
            "pastas.utils": utils,
            "pastas.timeseries": ts,
            "pastas.timeseries_utils": tsu,
            "pastas.version": ver,
            "pastas.plotting": plotting,
            "pastas.plotting.plotutil": plotutil,
        }
        for k, v in mods.items():
            sys.modules[k] = v
        pastas.decorators, pastas.utils, pastas.timeseries = dec, utils, ts
        pastas.timeseries_utils, pastas.version, pastas.plotting = tsu, ver, plotting
        plotting.plotutil = plotutil

    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            mpl.__path__ = []
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt


def load_reference():
    """Return the reference ``metran`` package (imported from REFERENCE_ROOT)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    try:
        metran = importlib.import_module("metran")
    finally:
        if REFERENCE_ROOT in sys.path:
            sys.path.remove(REFERENCE_ROOT)
    # pandas>=3: Timedelta(to_offset("D")) raises (metran.py:262); the grid is always
    # daily (metran.py:571) so a = freq/1D = 1 for the default freq "D".
    import numpy as np

    def _phi(self, alpha):
        from pandas import Timedelta

        freq = self.settings["freq"]
        try:
            from pandas.tseries.frequencies import to_offset

            a = Timedelta(to_offset(freq)) / Timedelta(1, "D")
        except Exception:
            a = Timedelta(freq if any(c.isdigit() for c in freq) else "1" + freq) / Timedelta(1, "D")
        return np.exp(-a / alpha)

    metran.metran.Metran._phi = _phi
    return metran


def load_example_series():
    """The five residual series of examples/data (tests/conftest.py:11-25), sorted by name."""
    from pathlib import Path

    from pandas import read_csv

    path = Path(REFERENCE_ROOT) / "examples" / "data"
    out = []
    for fi in sorted(path.glob("*_res.csv")):
        s = read_csv(fi, header=0, index_col=0, parse_dates=True, date_format="%Y-%m-%d",
                     names=[fi.stem.split("_")[0]]).squeeze()
        out.append(s)
    return out
