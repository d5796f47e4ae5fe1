"""CPU-only: occupancy-grid interchange files (the reference writes legacy VTK through pyvista,
/root/reference/nerf/run_nerf_acc.py:200-204,359-367, and reads it back at visualization/visualization.py:158-162)."""
import numpy as np
import pytest
import torch

from nerf_for_angiography_b200 import gridio


def _grid(shape=(6, 5, 4), seed=0):
    return np.random.default_rng(seed).random(shape) < 0.3


@pytest.mark.parametrize("ascii", [False, True])
def test_vtk_round_trip(tmp_path, ascii):
    b = _grid()
    f = tmp_path / "coarsegrid.vtk"
    gridio.save_grid_vtk(f, b, ascii=ascii)
    head = f.read_bytes()[:200].decode("latin-1")
    assert head.startswith("# vtk DataFile Version") and "DATASET STRUCTURED_POINTS" in head
    assert "DIMENSIONS 7 6 5" in head and f"CELL_DATA {b.size}" in head        # dimensions = shape + 1 (run_nerf_acc.py:201)
    out = gridio.load_grid_vtk(f)
    assert out.dtype == bool and out.shape == b.shape and np.array_equal(out, b)
    # a torch tensor / 128^3 grid takes the same path
    big = torch.from_numpy(_grid((16, 16, 16), 3))
    gridio.save_grid_vtk(f, big)
    assert np.array_equal(gridio.load_grid_vtk(f), big.numpy())


def test_reads_the_field_form_and_skips_metadata(tmp_path):
    """pyvista / VTK >= 9 may write the cell array inside a FIELD block, follow arrays with METADATA blocks and use other
    integer types; the flat array is reshaped in C order exactly as the reference does."""
    b = _grid((3, 2, 2), 1)
    flat = b.astype(int).reshape(-1)
    txt = ("# vtk DataFile Version 5.1\nvtk output\nASCII\nDATASET STRUCTURED_POINTS\nDIMENSIONS 4 3 3\nSPACING 1 1 1\nORIGIN 0 0 0\n"
           "CELL_DATA 12\nFIELD FieldData 2\nother 1 12 float\n" + " ".join(["0.5"] * 12) + "\nMETADATA\nINFORMATION 0\n\n"
           "values 1 12 int\n" + " ".join(str(v) for v in flat) + "\nMETADATA\nINFORMATION 0\n\n"
           "POINT_DATA 36\nSCALARS values float\nLOOKUP_TABLE default\n" + " ".join(["7"] * 36) + "\n")
    f = tmp_path / "g.vtk"
    f.write_text(txt)
    assert np.array_equal(gridio.load_grid_vtk(f), b)
    # big-endian binary payload, 32-bit ints
    f.write_bytes(b"# vtk DataFile Version 3.0\nvtk output\nBINARY\nDATASET STRUCTURED_POINTS\nDIMENSIONS 4 3 3\nSPACING 1 1 1\nORIGIN 0 0 0\n"
                  b"CELL_DATA 12\nSCALARS values int 1\nLOOKUP_TABLE default\n" + flat.astype(">i4").tobytes() + b"\n")
    assert np.array_equal(gridio.load_grid_vtk(f), b)


def test_malformed_files_are_rejected(tmp_path):
    f = tmp_path / "bad.vtk"
    f.write_text("hello\n")
    with pytest.raises(ValueError):
        gridio.load_grid_vtk(f)
    gridio.save_grid_vtk(f, _grid())
    with pytest.raises(ValueError):
        gridio.load_grid_vtk(f, name="missing")
    data = f.read_bytes()
    f.write_bytes(data[:-40])                                                     # truncated payload
    with pytest.raises(ValueError):
        gridio.load_grid_vtk(f)
    with pytest.raises(ValueError):
        gridio.save_grid_vtk(f, np.zeros((4, 4)))


def test_npy_and_assignment_to_an_occupancy_grid(tmp_path):
    import nerf_for_angiography_b200 as A
    b = _grid((8, 8, 8), 2)
    gridio.save_grid_npy(tmp_path / "g.npy", b)
    assert np.array_equal(gridio.load_grid_npy(tmp_path / "g.npy"), b)
    g = A.OccupancyGrid(torch.tensor([-100, -100, -100, 100, 100, 100.0]), 8, A.ContractionType.AABB)
    gridio.assign_binary(g, b)
    assert g.binary.dtype == torch.bool and np.array_equal(g.binary.numpy(), b)
    gridio.save_grid_vtk(tmp_path / "g.vtk", g)                                   # an OccupancyGrid is accepted directly
    assert np.array_equal(gridio.load_grid_vtk(tmp_path / "g.vtk"), b)
    with pytest.raises(ValueError):
        gridio.assign_binary(g, _grid((4, 4, 4)))
