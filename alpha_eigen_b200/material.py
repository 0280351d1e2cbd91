"""Materials (mirror of reference material.py) plus an explicit multigroup table material.

SimpleMaterial reproduces the hard-coded one-group Kornreich-Parsons data (material.py:9-20).
Material reads the reference's HDF5 layout (material.py:36-46) when h5py is available.
TableMaterial carries explicit multigroup arrays; it is what the synthetic 12- and 70-group
benchmark slabs use (the reference has no runnable multigroup path).
"""
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


class Material(SimpleMaterial):
    """HDF5-backed multigroup material, file layout of material.py:30,36-46."""

    def __init__(self, material_type, num_groups, data_dir="./data"):
        super().__init__(material_type, num_groups)
        self.name = material_type
        self.filename = f"{data_dir}/{num_groups}_group/{material_type}.hdf5"
        self.group_edges = None
        self.group_centers = None
        self.num_energy_groups = num_groups
        self.parse_file()

    def parse_file(self):
        try:
            import h5py
        except ImportError as e:  # the reference imports h5py at module scope (material.py:2)
            raise ImportError("Material needs h5py to read " + self.filename) from e
        with h5py.File(self.filename, "r") as hf:
            self.scatter_xs = hf["scat_xs"][:].transpose()
            self.total_xs = hf["total_xs"][:]
            self.nu_sigmaf = hf["nusigma_f"][:]
            self.inv_speed = hf["inverse_speed"][:]
            self.group_edges = hf["group_edges"][:]
            self.chi = np.zeros(self.num_energy_groups)         # material.py:43 (file's chi is ignored)
            self.nusigma_f = self.chi.copy()
        self.group_centers = 0.5 * (self.group_edges[:-1] + self.group_edges[1:])
        self.num_energy_groups = np.size(self.group_centers)


def material_row(mat, G):
    """(total, scatter[G,G], chi, nusf, inv_speed) as arrays for a material of either kind."""
    one = lambda v: np.full(G, float(v)) if np.ndim(v) == 0 else np.asarray(v, dtype=np.float64).reshape(G)
    sc = mat.scatter_xs
    sc = np.full((1, 1), float(sc)) if np.ndim(sc) == 0 else np.asarray(sc, dtype=np.float64).reshape(G, G)
    return one(mat.total_xs), sc, one(mat.chi), one(mat.nu_sigmaf), one(mat.inv_speed)
