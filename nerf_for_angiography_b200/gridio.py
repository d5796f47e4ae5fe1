"""Occupancy-grid interchange with the reference's files.

The reference keeps its occupancy grids as legacy-VTK image data written through pyvista
(/root/reference/nerf/run_nerf_acc.py:200-204,359-367,384-385): a `UniformGrid` with `dimensions = binary.shape + 1`, unit
spacing, origin 0 and ONE cell array `values = binary.astype('int').flatten()` (C order of the [x, y, z] array).  The
inference script reads it back as `grid['values'].reshape(dimensions - 1)` and assigns it to `acc_grid._binary`
(/root/reference/visualization/visualization.py:158-162) -- so the flat array is taken as is, whatever VTK's own cell order is.

pyvista / vtk are not needed here: `save_grid_vtk` writes the same legacy "STRUCTURED_POINTS + CELL_DATA" file (readable by
`pv.get_reader(...).read()`), `load_grid_vtk` parses what pyvista writes (ASCII or BINARY; the array as SCALARS or inside a
FIELD block; VTK >= 9 METADATA blocks are skipped).  `save_grid_npy` / `load_grid_npy` are the dependency-free alternative.
Host-side I/O only; nothing here is on the hot path."""
import re

import numpy as np
import torch

# legacy-VTK type names -> numpy dtypes (files are big-endian in BINARY mode)
_VTK_TYPES = {"bit": None, "unsigned_char": "u1", "char": "i1", "unsigned_short": "u2", "short": "i2", "unsigned_int": "u4", "int": "i4",
              "unsigned_long": "u8", "long": "i8", "vtktypeint64": "i8", "vtktypeuint64": "u8", "vtktypeint32": "i4", "vtktypeuint32": "u4",
              "vtktypeint8": "i1", "vtktypeuint8": "u1", "vtktypeint16": "i2", "vtktypeuint16": "u2", "float": "f4", "double": "f8",
              "vtkidtype": "i8"}


def _as_numpy(binary):
    if hasattr(binary, "binary"):                               # an OccupancyGrid
        binary = binary.binary
    if isinstance(binary, torch.Tensor):
        binary = binary.detach().cpu().numpy()
    b = np.asarray(binary)
    if b.ndim != 3:
        raise ValueError("occupancy grid must be a 3-d array")
    return b


def save_grid_vtk(filename, binary, name="values", ascii=False):
    """Write `binary` ([nx, ny, nz] bool / int, or an OccupancyGrid) the way `pv_grid.save(...)` does in the reference."""
    b = _as_numpy(binary)
    flat = np.ascontiguousarray(b.astype(np.int64).reshape(-1))             # .astype('int').flatten()
    nx, ny, nz = b.shape
    head = ("# vtk DataFile Version 3.0\nvtk output\n%s\nDATASET STRUCTURED_POINTS\nDIMENSIONS %d %d %d\nSPACING 1 1 1\nORIGIN 0 0 0\n"
            "CELL_DATA %d\nSCALARS %s vtktypeint64\nLOOKUP_TABLE default\n") % ("ASCII" if ascii else "BINARY", nx + 1, ny + 1, nz + 1, flat.size, name)
    with open(filename, "wb") as f:
        f.write(head.encode("ascii"))
        if ascii:
            for i in range(0, flat.size, 9):
                f.write((" ".join(str(int(v)) for v in flat[i:i + 9]) + "\n").encode("ascii"))
        else:
            f.write(flat.astype(">i8").tobytes())
            f.write(b"\n")


class _Reader:
    def __init__(self, data):
        self.d, self.pos = data, 0

    def line(self):
        """next non-empty line as text (None at the end of the file)"""
        while self.pos < len(self.d):
            end = self.d.find(b"\n", self.pos)
            end = len(self.d) if end < 0 else end
            s = self.d[self.pos:end].decode("latin-1").strip()
            self.pos = end + 1
            if s:
                return s
        return None

    def values(self, n, vtk_type, binary):
        dt = _VTK_TYPES.get(vtk_type.lower())
        if dt is None:
            raise ValueError(f"unsupported VTK data type {vtk_type!r}")
        if binary:
            nbytes = n * np.dtype(dt).itemsize
            raw = self.d[self.pos:self.pos + nbytes]
            if len(raw) != nbytes:
                raise ValueError("truncated VTK file")
            self.pos += nbytes
            return np.frombuffer(raw, dtype=">" + dt).astype(dt)
        out = []
        while len(out) < n:
            s = self.line()
            if s is None:
                raise ValueError("truncated VTK file")
            out.extend(s.split())
        if len(out) != n:
            raise ValueError("VTK array length does not match its header")
        return np.array([float(v) for v in out]).astype(dt)


def load_grid_vtk(filename, name="values"):
    """Read a legacy-VTK image-data file with a cell array `name` and return it as a bool array of shape DIMENSIONS - 1
    (`grid['values'].reshape(np.array(grid.dimensions) - 1)`, visualization.py:160)."""
    with open(filename, "rb") as f:
        rd = _Reader(f.read())
    if not (rd.line() or "").startswith("# vtk DataFile"):
        raise ValueError("not a legacy VTK file")
    rd.line()                                                            # title
    mode = (rd.line() or "").upper()
    if mode not in ("ASCII", "BINARY"):
        raise ValueError("VTK file: expected ASCII or BINARY")
    binary = mode == "BINARY"
    dims, n_cells, found, in_cells = None, None, None, False
    while True:
        s = rd.line()
        if s is None:
            break
        tok = s.split()
        key = tok[0].upper()
        if key == "DATASET":
            if tok[1].upper() != "STRUCTURED_POINTS":
                raise ValueError(f"VTK dataset {tok[1]} is not image data (STRUCTURED_POINTS)")
        elif key == "DIMENSIONS":
            dims = tuple(int(v) for v in tok[1:4])
        elif key == "CELL_DATA":
            n_cells, in_cells = int(tok[1]), True
        elif key == "POINT_DATA":
            in_cells = False
        elif key == "SCALARS":
            ncomp = int(tok[3]) if len(tok) > 3 else 1
            nxt = rd.line()
            if nxt is None or not nxt.upper().startswith("LOOKUP_TABLE"):
                raise ValueError("VTK file: SCALARS without LOOKUP_TABLE")
            n = (n_cells if in_cells else int(np.prod(dims))) * ncomp
            arr = rd.values(n, tok[2], binary)
            if in_cells and tok[1] == name:
                found = arr
        elif key == "FIELD":
            for _ in range(int(tok[2])):
                h = rd.line()
                while h is not None and re.match(r"(METADATA|INFORMATION|NAME |DATA )", h.upper()):
                    h = rd.line()
                an, ncomp, ntup, ty = h.split()
                arr = rd.values(int(ncomp) * int(ntup), ty, binary)
                if in_cells and an == name:
                    found = arr
        # SPACING / ORIGIN / ASPECT_RATIO / METADATA / INFORMATION ...: nothing to keep
    if dims is None or found is None:
        raise ValueError(f"VTK file has no cell array {name!r}")
    shape = tuple(d - 1 for d in dims)
    if found.size != int(np.prod(shape)):
        raise ValueError("VTK cell array does not match DIMENSIONS - 1")
    return found.reshape(shape) != 0


def save_grid_npy(filename, binary):
    np.save(filename, _as_numpy(binary).astype(bool))


def load_grid_npy(filename):
    b = np.load(filename)
    if b.ndim != 3:
        raise ValueError("occupancy grid must be a 3-d array")
    return b.astype(bool)


def assign_binary(grid, binary):
    """`acc_grid._binary = grid_occupancy` (visualization.py:162) with the shape / device checks the reference leaves out."""
    b = torch.as_tensor(np.asarray(binary)).bool()
    if tuple(b.shape) != (grid._resolution,) * 3:
        raise ValueError(f"grid file holds {tuple(b.shape)} cells, the OccupancyGrid has {(grid._resolution,) * 3}")
    grid._binary = b.to(grid.occs.device).contiguous()
    return grid
