"""Small host helpers (reference ``core/misc.py:8-22``)."""
import glob
import os
import re


def find_latest_iter_file(data_dir: str):
    """Returns (iteration, path) of the newest ``{iter:06d}-rho.npz``."""
    best = None
    for path in glob.glob(os.path.join(data_dir, "*-rho.npz")):
        m = re.match(r"(\d+)-rho\.npz$", os.path.basename(path))
        if m and (best is None or int(m.group(1)) > best[0]):
            best = (int(m.group(1)), path)
    if best is None:
        raise FileNotFoundError(f"no *-rho.npz checkpoint under {data_dir}")
    return best
