#!/usr/bin/env python
"""One-off conversion of the reference's HDF5 cross-section files into the packaged .npz copies.

    python tools/convert_xs_data.py [/root/reference/alpha_eigen_modules/data]

Reads <src>/<G>_group/<material>.hdf5 with alpha_eigen_b200.hdf5_lite (no h5py in this image) and writes
alpha_eigen_b200/data/<G>_group/<material>.npz holding the same datasets under the same names, untouched
(material.py:36-46 decides what to do with them: transpose of scat_xs, chi ignored, ...).  The .npz files are
what Material falls back on when ./data/<G>_group/<material>.hdf5 does not exist in the working directory.
Also copies the 5 KB one-group file as the reader's test fixture (tests/golden/dahl_problem.hdf5).
"""
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alpha_eigen_b200.hdf5_lite import File  # noqa: E402


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/alpha_eigen_modules/data"
    for group_dir in sorted(os.listdir(src)):
        for name in sorted(os.listdir(os.path.join(src, group_dir))):
            if not name.endswith(".hdf5"):
                continue
            with File(os.path.join(src, group_dir, name)) as hf:
                data = {k: hf[k][:] for k in hf.keys()}
            out_dir = os.path.join(ROOT, "alpha_eigen_b200", "data", group_dir)
            os.makedirs(out_dir, exist_ok=True)
            out = os.path.join(out_dir, name[:-5] + ".npz")
            np.savez_compressed(out, **data)
            print(out, {k: v.shape for k, v in data.items()})
    shutil.copyfile(os.path.join(src, "1_group", "dahl_problem.hdf5"), os.path.join(ROOT, "tests", "golden", "dahl_problem.hdf5"))


if __name__ == "__main__":
    main()
