"""Cross-section data path: the built-in HDF5 reader, the packaged copy of the
reference's multigroup data and Material.parse_file (material.py:26-46).  CPU only."""
import os
import shutil

import numpy as np
import pytest

from alpha_eigen_b200.hdf5_lite import File, Hdf5Error
from alpha_eigen_b200.material import Material

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DATA = "/root/reference/alpha_eigen_modules/data"


def test_reader_on_reference_fixture(golden_dir):
    """tests/golden/dahl_problem.hdf5 is the reference's own 1-group data file (tests/data/1_group)."""
    with File(os.path.join(golden_dir, "dahl_problem.hdf5")) as hf:
        assert sorted(hf.keys()) == ["group_centers", "group_edges", "inv_speed", "inverse_speed", "nusigma_f", "scat_xs",
                                     "total_xs"]
        assert "total_xs" in hf and "chi" not in hf
        assert hf["total_xs"].shape == (1,) and hf["total_xs"].dtype == np.int64
        np.testing.assert_array_equal(hf["group_edges"][:], [0, 1])
        np.testing.assert_array_equal(hf["scat_xs"][:], [1])
        np.testing.assert_array_equal(hf["nusigma_f"][:], [0])
        with pytest.raises(KeyError):
            hf["nope"]


def test_reader_rejects_what_it_cannot_read(tmp_path):
    bad = tmp_path / "bad.hdf5"
    bad.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(Hdf5Error, match="not an HDF5 file"):
        File(str(bad))
    newer = tmp_path / "v2.hdf5"
    newer.write_bytes(b"\x89HDF\r\n\x1a\n" + bytes([2]) + bytes(100))
    with pytest.raises(Hdf5Error, match="superblock version 2"):
        File(str(newer))
    with pytest.raises(Hdf5Error, match="read-only"):
        File(str(bad), "w")


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("rel", ["70_group/metal", "70_group/polyethylene", "1_group/dahl_problem"])
def test_packaged_data_equals_reference_files(rel):
    """The .npz copies packaged with the module hold exactly what the reference's HDF5 files hold."""
    with File(os.path.join(REF_DATA, rel + ".hdf5")) as hf, \
            np.load(os.path.join(ROOT, "alpha_eigen_b200", "data", rel + ".npz")) as z:
        assert sorted(hf.keys()) == sorted(z.files)
        for k in z.files:
            a = hf[k][:]
            assert a.dtype == z[k].dtype and a.shape == z[k].shape
            np.testing.assert_array_equal(a, z[k])


def test_material_from_packaged_data():
    m = Material("metal", 70)
    assert m.num_energy_groups == 70 and m.filename == "./data/70_group/metal.hdf5"
    assert m.scatter_xs.shape == (70, 70) and m.total_xs.shape == m.nu_sigmaf.shape == m.inv_speed.shape == (70,)
    with np.load(os.path.join(ROOT, "alpha_eigen_b200", "data", "70_group", "metal.npz")) as z:
        np.testing.assert_array_equal(m.scatter_xs, z["scat_xs"].T)            # material.py:38
        np.testing.assert_array_equal(m.total_xs, z["total_xs"])
        np.testing.assert_array_equal(m.group_centers, 0.5 * (z["group_edges"][:-1] + z["group_edges"][1:]))
    assert m.group_edges[0] == 17.0 and m.group_edges.shape == (71,)
    assert not m.chi.any() and not m.nusigma_f.any()                           # material.py:43-44
    # physically sensible: scattering never exceeds the total cross-section
    assert np.all(m.scatter_xs.sum(axis=1) <= m.total_xs * (1 + 1e-12))
    p = Material("polyethylene", 70, use_file_chi=True)
    assert abs(p.chi.sum() - 1) < 1e-9
    d = Material("dahl_problem", 1)
    assert d.scatter_xs.shape == (1, 1) and d.total_xs[0] == 1.0 and d.num_energy_groups == 1


def test_material_reads_working_directory_file_first(tmp_path, monkeypatch, golden_dir):
    """material.py:30: './data/<G>_group/<name>.hdf5' relative to the working directory."""
    os.makedirs(tmp_path / "data" / "1_group")
    shutil.copyfile(os.path.join(golden_dir, "dahl_problem.hdf5"), tmp_path / "data" / "1_group" / "mine.hdf5")
    monkeypatch.chdir(tmp_path)
    m = Material("mine", 1)
    assert m.total_xs[0] == 1.0 and m.inv_speed[0] == 1.0
    with pytest.raises(FileNotFoundError):
        Material("missing", 1)
