"""Materials (mirror of reference material.py) plus an explicit multigroup table material.

SimpleMaterial reproduces the hard-coded one-group Kornreich-Parsons data (material.py:9-20).
Material reads the reference's HDF5 cross-section files (material.py:36-46) with h5py or, when that is not
installed, the built-in reader hdf5_lite; the reference's own data set is packaged as .npz.
TableMaterial carries explicit multigroup arrays; it is what the synthetic 12- and 70-group
benchmark slabs use (the reference has no runnable multigroup path).
"""
import os

import numpy as np

_KP = {  # name -> (scatter_xs, nu_sigmaf)      material.py:12-20
    "fuel": (0.8, 0.7),
    "moderator": (0.8, 0.0),
    "absorber": (0.1, 0.0),
}


class SimpleMaterial(object):
    def __init__(self, material_type, num_groups):
        self.name = material_type
        self.num_groups = num_groups
        self.total_xs = 1.0
        self.chi = 1.0
        self.inv_speed = 1.0
        self.scatter_xs, self.nu_sigmaf = _KP.get(material_type, (None, None))


class TableMaterial(SimpleMaterial):
    """Explicit G-group data: total_xs, nu_sigmaf, chi, inv_speed (G,), scatter_xs (G,G) with
    scatter_xs[g_from, g_to] (the orientation material.py:38 produces by transposing the file)."""

    def __init__(self, material_type, total_xs, scatter_xs, nu_sigmaf, chi, inv_speed):
        total_xs = np.atleast_1d(np.asarray(total_xs, dtype=np.float64))
        super().__init__(material_type, len(total_xs))
        G = self.num_groups
        self.total_xs = total_xs
        self.scatter_xs = np.asarray(scatter_xs, dtype=np.float64).reshape(G, G)
        self.nu_sigmaf = np.asarray(nu_sigmaf, dtype=np.float64).reshape(G)
        self.chi = np.asarray(chi, dtype=np.float64).reshape(G)
        self.inv_speed = np.asarray(inv_speed, dtype=np.float64).reshape(G)


_PACKAGED_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _open_hdf5(path):
    """h5py when it is installed, else the built-in reader (same calls: hf[name][:])."""
    try:
        import h5py
        return h5py.File(path, "r")
    except ImportError:
        from .hdf5_lite import File
        return File(path, "r")


class Material(SimpleMaterial):
    """Multigroup material read from the reference's cross-section files (material.py:26-46).

    File lookup: `<data_dir>/<G>_group/<name>.hdf5` relative to the working directory, exactly as
    material.py:30 does; when that file does not exist, the copy of the reference's data set packaged with this
    module (alpha_eigen_b200/data/<G>_group/<name>.npz, written by tools/convert_xs_data.py: same datasets, same
    names) is used, so `Material('metal', 70)` works from any directory.

    material.py:43 discards the file's fission spectrum (chi = zeros), which makes every fission source vanish;
    that behaviour is kept by default and `use_file_chi=True` wires the `chi` dataset in.
    """

    def __init__(self, material_type, num_groups, data_dir="./data", use_file_chi=False):
        super().__init__(material_type, num_groups)
        self.name = material_type
        self.filename = f"{data_dir}/{num_groups}_group/{material_type}.hdf5"
        self.use_file_chi = use_file_chi
        self.group_edges = None
        self.group_centers = None
        self.num_energy_groups = num_groups
        self.parse_file()

    def _datasets(self):
        if os.path.exists(self.filename):
            with _open_hdf5(self.filename) as hf:
                return {k: np.asarray(hf[k][:]) for k in hf.keys()}
        packaged = os.path.join(_PACKAGED_DATA, f"{self.num_groups}_group", f"{self.name}.npz")
        if not os.path.exists(packaged):
            raise FileNotFoundError(f"{self.filename} not found and no packaged copy at {packaged}")
        with np.load(packaged) as z:
            return {k: z[k] for k in z.files}

    def parse_file(self):
        hf = self._datasets()
        as_f64 = lambda a: np.atleast_1d(np.asarray(a, dtype=np.float64))
        self.scatter_xs = as_f64(hf["scat_xs"]).transpose()         # material.py:38: [g_from, g_to] after the transpose
        self.total_xs = as_f64(hf["total_xs"])
        self.nu_sigmaf = as_f64(hf["nusigma_f"])
        self.inv_speed = as_f64(hf["inverse_speed"])
        self.group_edges = as_f64(hf["group_edges"])
        self.chi = np.zeros(self.num_energy_groups)                 # material.py:43 (the file's chi is ignored)
        if self.use_file_chi and "chi" in hf:
            self.chi = as_f64(hf["chi"])
        self.nusigma_f = np.zeros(self.num_energy_groups)           # material.py:44
        self.group_centers = (self.group_edges[0:-1] + self.group_edges[1:]) * 0.5
        self.num_energy_groups = np.size(self.group_centers)
        if self.scatter_xs.ndim == 1 and self.num_energy_groups == 1:
            self.scatter_xs = self.scatter_xs.reshape(1, 1)


def material_row(mat, G):
    """(total, scatter[G,G], chi, nusf, inv_speed) as arrays for a material of either kind."""
    one = lambda v: np.full(G, float(v)) if np.ndim(v) == 0 else np.asarray(v, dtype=np.float64).reshape(G)
    sc = mat.scatter_xs
    sc = np.full((1, 1), float(sc)) if np.ndim(sc) == 0 else np.asarray(sc, dtype=np.float64).reshape(G, G)
    return one(mat.total_xs), sc, one(mat.chi), one(mat.nu_sigmaf), one(mat.inv_speed)
