"""export_vtk / vtk_grid (src/mesher.jl:304-345, 1138-1185; src/ImmersedBoundary.jl:1277-1329): structure and content of
the files written for test/rae2822.jl's last lines (`export_vtk("rae2822", dom; ny = ny)`), read back with an XML
parser.  Volume only on CPU (surface values go through a device kernel; covered in the GPU suite)."""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

F32 = np.float32


def _arr(node):
    return np.array(node.text.split(), dtype=np.float64)


def test_volume_export_round_trip(get_case, ib, tmp_path):
    c = get_case("advection")
    msh = c.msh
    N = len(msh)
    rng = np.random.default_rng(3)
    u, U = rng.random(N).astype(F32), rng.random((N, 2)).astype(F32)
    out = ib.export_vtk(str(tmp_path / "adv"), c.dom, export_surface=False, u=u, U=U)
    root = ET.parse(out["volume"]).getroot()
    sets = root.findall(".//DataSet")
    assert len(sets) == msh.nblocks == 187
    nper = msh.block_size ** msh.nd
    centers, widths = ib.get_cells(msh)
    for b in (0, 17, msh.nblocks - 1):
        g = ET.parse(os.path.join(str(tmp_path / "adv"), sets[b].get("file"))).getroot()
        coords = [_arr(a) for a in g.find(".//Coordinates")]
        assert len(coords[0]) == len(coords[1]) == msh.block_size + 1 and len(coords[2]) == 1
        assert np.allclose(coords[0][[0, -1]], [msh.block_origins[b, 0], msh.block_origins[b, 0] + msh.block_widths[b, 0]])
        # cell centres implied by the grid lines are the mesh's own cell centres, first dimension fastest
        xc = 0.5 * (coords[0][1:] + coords[0][:-1])
        yc = 0.5 * (coords[1][1:] + coords[1][:-1])
        blk = centers[b * nper:(b + 1) * nper]
        assert np.allclose(blk[:, 0], np.tile(xc, msh.block_size), atol=1e-6) and np.allclose(blk[:, 1], np.repeat(yc, msh.block_size), atol=1e-6)
        data = {a.get("Name"): a for a in g.find(".//CellData")}
        assert np.array_equal(_arr(data["u"]).astype(F32), u[b * nper:(b + 1) * nper])
        assert data["U"].get("NumberOfComponents") == "2"
        assert np.array_equal(_arr(data["U"]).astype(F32).reshape(-1, 2), U[b * nper:(b + 1) * nper])
    # a subset of blocks, written again into the same folder (overwrite warning)
    import warnings
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        out = ib.export_vtk(str(tmp_path / "adv"), c.dom, [3, 5], export_surface=False, u=u)
    assert any("Overwriting" in str(x.message) for x in w)
    assert [s.get("name") for s in ET.parse(out["volume"]).getroot().findall(".//DataSet")] == ["block_3", "block_5"]


def test_surface_grid_of_a_stereolitography(ib, tmp_path):
    stl = ib.merge_points(ib.Stereolitography(os.path.join(os.path.dirname(__file__), "golden", "rae2822.dat")))
    pts, simp = stl.points, stl.simplices
    val = np.arange(len(simp), dtype=F32)
    path = ib.vtk_grid_stl(str(tmp_path / "wall.vtu"), stl, cellval=val, xy=pts.astype(F32))
    piece = ET.parse(path).getroot().find(".//Piece")
    assert int(piece.get("NumberOfPoints")) == len(pts) and int(piece.get("NumberOfCells")) == len(simp)
    cells = {a.get("Name"): a for a in piece.find("Cells")}
    assert np.array_equal(_arr(cells["connectivity"]).astype(int).reshape(-1, 2), simp)
    assert set(cells["types"].text.split()) == {"3"}                 # VTK_LINE in 2-D
    named = {a.get("Name"): a for tag in ("PointData", "CellData") for a in piece.find(tag)}
    assert np.array_equal(_arr(named["cellval"]).astype(F32), val)
    # a closed 2-D curve has as many segments as points: WriteVTK-style length matching cannot tell the two apart
    assert np.array_equal(_arr(named["xy"]).astype(F32).reshape(-1, 2), pts.astype(F32))


def test_mesh_save_load_round_trip(ib, tmp_path):
    """Mesh (de)serialisation: a saved and re-loaded mesh gives the same blocks, cells, refined STL and
    -- through rebuilt distance fields -- the same Domain tables (ghosts, donors, weights), 2-D STL and 3-D analytic."""
    import os
    rae = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rae2822.dat")
    stl = ib.merge_points(ib.Stereolitography(rae))
    m2 = ib.Mesh(np.array([-25, -25], np.float32), np.array([50, 50], np.float32), ("wall", stl, np.float32(2e-2)))
    m3 = ib.Mesh([-2, -2, -2], [4, 4, 4], ("wall", ib.Sphere([0, 0, 0], 0.5), np.float32(0.12)))
    for k, (m, fams) in enumerate(((m2, [("farfield", [(0, False), (0, True), (1, False), (1, True)])]),
                                   (m3, [("farfield", [(d, s) for d in range(3) for s in (False, True)])]))):
        path = tmp_path / f"mesh{k}.ibx"
        m.save(path)
        r = ib.Mesh.load(path)
        assert (r.nd, r.block_size, r.nblocks, r.ncells) == (m.nd, m.block_size, m.nblocks, m.ncells)
        assert np.array_equal(r.block_origins, m.block_origins) and np.array_equal(r.block_widths, m.block_widths)
        assert sorted(r.distance_fields) == sorted(m.distance_fields)
        da = ib.Domain(m, hypercube_families=fams, upload=False)
        db = ib.Domain(r, hypercube_families=fams, upload=False)
        assert np.array_equal(da.faces(), db.faces())
        for name in da.boundaries:
            for key, b in da.boundaries[name].items():
                c = db.boundaries[name][key]
                assert np.array_equal(b.ghost_indices, c.ghost_indices) and np.array_equal(b.projections, c.projections)
                assert np.array_equal(b.image_domain, c.image_domain) and np.array_equal(b.interp_w, c.interp_w)
    with pytest.raises(ib.IbxError):
        ib.Mesh.load(rae)                      # not a mesh file
