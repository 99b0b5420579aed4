"""Small host helpers (reference ``core/misc.py:8-22``)."""
import glob
import os
import re


def find_latest_iter_file(dst_path: str):
    """Returns (iteration, path) of the newest ``{iter:06d}-rho.npz``."""
    best = None
    for path in glob.glob(os.path.join(dst_path, "*-rho.npz")):
        m = re.match(r"(\d+)-rho\.npz$", os.path.basename(path))
        if m and (best is None or int(m.group(1)) > best[0]):
            best = (int(m.group(1)), path)
    if best is None:
        raise FileNotFoundError(f"no *-rho.npz checkpoint under {dst_path}")
    return best


def str2bool(value) -> bool:
    """argparse helper of the reference's example scripts (``core/misc.py``)."""
    if isinstance(value, bool):
        return value
    v = str(value).strip().lower()
    if v in ("true", "1", "yes", "y", "t"):
        return True
    if v in ("false", "0", "no", "n", "f"):
        return False
    import argparse
    raise argparse.ArgumentTypeError(f"boolean value expected, got {value!r}")


def float_or_none(x):
    """``"none"`` (any case) -> None, anything else -> float."""
    if x is None or str(x).strip().lower() == "none":
        return None
    return float(x)
