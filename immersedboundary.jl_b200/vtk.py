"""Host-side output: ``vtk_grid`` / ``export_vtk`` (``src/mesher.jl:304-345, 1138-1185``, ``src/ImmersedBoundary.jl:1277-1329``).

The reference writes through WriteVTK.jl (an un-vendored dependency): one rectilinear grid per octree block collected
in a multi-block file, and one unstructured grid per surface.  This module writes the same structure as plain VTK XML
(ASCII data arrays, no third-party package): ``<fname>/VOLUME.vtm`` -> ``<fname>/VOLUME/block_<i>.vtr`` with the cell
data of each block (cells are block-major, first dimension fastest, ``src/mesher.jl:1079-1092``), and
``<fname>/SURFACE.vtm`` -> ``<fname>/<surface name>.vtu`` with line / triangle cells and the field values interpolated
to the surface (``surf(v)``).  Pure I/O on host copies: device arrays are downloaded, nothing is computed here."""
import os
import shutil
import warnings

import numpy as np

F32 = np.float32


def _host(a):
    return a.to_host() if hasattr(a, "to_host") else np.asarray(a)


def _fmt(a):
    a = np.asarray(a)
    return " ".join(repr(float(x)) if a.dtype.kind == "f" else str(int(x)) for x in a.ravel())


def _data_array(name, a, ncomp=None):
    a = np.asarray(a)
    vt = "Float32" if a.dtype.kind == "f" else "Int32"
    nc = "" if ncomp in (None, 1) else f' NumberOfComponents="{ncomp}"'
    return f'<DataArray type="{vt}" Name="{name}"{nc} format="ascii">{_fmt(a)}</DataArray>'


def _fix_export(v):
    """``_fix_export`` (``src/mesher.jl:1113-1123``): point index last -> rows are points here: (points, components)."""
    v = _host(v)
    return v if v.ndim == 1 else v.reshape(v.shape[0], -1)


def _write_vtm(path, files):
    with open(path, "w") as fh:
        fh.write('<?xml version="1.0"?>\n<VTKFile type="vtkMultiBlockDataSet" version="1.0" byte_order="LittleEndian">\n'
                 "<vtkMultiBlockDataSet>\n")
        for i, (name, rel) in enumerate(files):
            fh.write(f'<DataSet index="{i}" name="{name}" file="{rel}"/>\n')
        fh.write("</vtkMultiBlockDataSet>\n</VTKFile>\n")


def vtk_grid_mesh(fname, msh, block_indices=None, make_folder=True, **fields):
    """``vtk_grid(fname, msh, partition_indices; kwargs...)`` (``src/mesher.jl:1138-1185``): one rectilinear grid per
    block (points ``LinRange(o, o + w, block_size + 1)`` per dimension) with ``fields`` as cell data; 0-based block ids."""
    nd, bs = msh.nd, msh.block_size
    bo, bw = msh.block_origins, msh.block_widths
    nper = bs ** nd
    blocks = range(bo.shape[0]) if block_indices is None else list(block_indices)
    if make_folder:
        if os.path.isdir(fname):
            warnings.warn(f"Overwriting volume output in folder {fname}.")
            shutil.rmtree(fname)
        os.makedirs(fname)
    os.makedirs(os.path.join(fname, "VOLUME"), exist_ok=True)
    host = {k: _fix_export(v) for k, v in fields.items()}
    files = []
    for b in blocks:
        rel = os.path.join("VOLUME", f"block_{b}.vtr")
        ext = " ".join(f"0 {bs}" for _ in range(nd)) + " 0 0" * (3 - nd)
        with open(os.path.join(fname, rel), "w") as fh:
            fh.write('<?xml version="1.0"?>\n<VTKFile type="RectilinearGrid" version="1.0" byte_order="LittleEndian">\n')
            fh.write(f'<RectilinearGrid WholeExtent="{ext}">\n<Piece Extent="{ext}">\n<CellData>\n')
            for k, v in host.items():
                blk = v[b * nper:(b + 1) * nper]
                fh.write(_data_array(k, blk, None if blk.ndim == 1 else blk.shape[1]) + "\n")
            fh.write("</CellData>\n<Coordinates>\n")
            for d in range(3):
                c = np.linspace(F32(bo[b, d]), F32(bo[b, d]) + F32(bw[b, d]), bs + 1, dtype=F32) if d < nd else np.zeros(1, F32)
                fh.write(_data_array("xyz"[d], c) + "\n")
            fh.write("</Coordinates>\n</Piece>\n</RectilinearGrid>\n</VTKFile>\n")
        files.append((f"block_{b}", rel))
    vtm = os.path.join(fname, "VOLUME.vtm")
    _write_vtm(vtm, files)
    return vtm


def vtk_grid_stl(path, stl, **fields):
    """``vtk_grid(fname, stl; kwargs...)`` (``src/mesher.jl:304-345``): line (2-D) or triangle (3-D) cells; an array whose
    length matches the points becomes point data, one matching the simplices cell data."""
    pts, simp = np.asarray(stl.points, dtype=F32), np.asarray(stl.simplices)
    npts, nd = pts.shape
    ncell, nvert = simp.shape
    p3 = np.zeros((npts, 3), F32)
    p3[:, :nd] = pts
    pd, cd = {}, {}
    for k, v in fields.items():
        v = _fix_export(v)
        (cd if v.shape[0] == ncell else pd)[k] = v
    with open(path, "w") as fh:
        fh.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian">\n<UnstructuredGrid>\n')
        fh.write(f'<Piece NumberOfPoints="{npts}" NumberOfCells="{ncell}">\n<Points>\n{_data_array("Points", p3, 3)}\n</Points>\n<Cells>\n')
        fh.write(_data_array("connectivity", simp.astype(np.int32)) + "\n")
        fh.write(_data_array("offsets", (np.arange(ncell, dtype=np.int32) + 1) * nvert) + "\n")
        fh.write(f'<DataArray type="UInt8" Name="types" format="ascii">{" ".join(["3" if nd == 2 else "5"] * ncell)}</DataArray>\n</Cells>\n')
        for tag, dd in (("PointData", pd), ("CellData", cd)):
            fh.write(f"<{tag}>\n")
            for k, v in dd.items():
                fh.write(_data_array(k, v, None if v.ndim == 1 else v.shape[1]) + "\n")
            fh.write(f"</{tag}>\n")
        fh.write("</Piece>\n</UnstructuredGrid>\n</VTKFile>\n")
    return path


def export_vtk(fname, dom, block_indices=None, surface_data=None, export_volume=True, export_surface=True, **fields):
    """``export_vtk(fname, dom, block_indices; surface_data, export_volume, export_surface, kwargs...)``
    (``src/ImmersedBoundary.jl:1277-1329``): volume fields per block, and per surface the same fields interpolated to
    the surface (``surf(v)``) plus the entries of ``surface_data[name]`` (a dict of arrays)."""
    if os.path.isdir(fname):
        warnings.warn(f"Overwriting output in folder {fname}.")
        shutil.rmtree(fname)
    os.makedirs(fname)
    out = {}
    if export_volume:
        out["volume"] = vtk_grid_mesh(fname, dom.mesh, block_indices, make_folder=False, **fields)
    if export_surface:
        files = []
        for sname, surf in dom.surfaces.items():
            stl = dom.mesh.distance_fields[sname].stl
            if stl is None:      # analytic surface: no simplices to write
                continue
            data = {k: _fix_export(surf(_host(v).astype(F32))) for k, v in fields.items()}
            for k, v in (surface_data or {}).get(sname, {}).items():
                data[k] = _fix_export(v)
            vtk_grid_stl(os.path.join(fname, f"{sname}.vtu"), stl, **data)
            files.append((sname, f"{sname}.vtu"))
        out["surface"] = os.path.join(fname, "SURFACE.vtm")
        _write_vtm(out["surface"], files)
    return out
