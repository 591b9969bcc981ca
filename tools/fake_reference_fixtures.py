"""Self-test of the pinning kit's file format: writes, FROM THE ORACLE, the same files tools/gen_reference_fixtures.jl
writes from the real package (case `advection` only), so that tests/test_reference_fixtures.py can be exercised without
Julia.  Output goes to a scratch directory -- never to tests/golden/reference, which is reserved for real reference data.

    python tools/fake_reference_fixtures.py /tmp/fake_ref && IBX_REFERENCE_FIXTURES=/tmp/fake_ref pytest tests/test_reference_fixtures.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from oracle import cfd, euler  # noqa: E402
from oracle.domain import JST_sensor, MUSCL, at_faces, cell_gradient, green_gauss, unsigned_green_gauss  # noqa: E402

F32 = np.float32
out = sys.argv[1]
os.makedirs(out, exist_ok=True)
man = {}
NAMES = {np.dtype(np.float32): "Float32", np.dtype(np.float64): "Float64", np.dtype(np.int64): "Int64"}


def dump(name, a):
    a = np.asarray(a)
    np.asfortranarray(a).ravel(order="F").tofile(os.path.join(out, name + ".bin"))
    man[name] = {"eltype": NAMES[a.dtype], "size": list(a.shape)}


def dump_lists(name, lists):
    ptr = np.cumsum([0] + [len(l) for l in lists]).astype(np.int64)
    val = np.concatenate(lists) if lists else np.zeros(0)
    dump(name + "_ptr", ptr)
    dump(name + "_val", val)


import immersedboundary_jl_b200 as ib  # noqa: E402  (only for conftest.Case, which builds both sides)
from conftest import Case  # noqa: E402
c = Case("advection", ib, oracle, 100_000)
tag, msh, dom = "advection", c.omsh, c.odom
N, nd = dom.centers.shape
dump(tag + "_block_origins", msh.block_origins.T)      # Julia keeps (nd, nblocks)
dump(tag + "_block_widths", msh.block_widths.T)
dump(tag + "_centers", dom.centers)
dump(tag + "_widths", dom.widths)
pids = sorted(dom.partitions)
dump(tag + "_partition_ids", np.array(pids, np.int64))
for pid in pids:
    p, part = f"{tag}_p{pid}", dom.partitions[pid]
    dump(p + "_domain", part.domain.astype(np.int64) + 1)
    dump(p + "_image", part.image.astype(np.int64) + 1)
    dump(p + "_image_in_domain", part.image_in_domain.astype(np.int64) + 1)
    for d in range(nd):
        o, n = part.face_owners_neighbors[d]
        dump(f"{p}_own{d + 1}", o.astype(np.int64) + 1)
        dump(f"{p}_nei{d + 1}", n.astype(np.int64) + 1)
for bname, chunks in dom.boundaries.items():
    dump(f"{tag}_b_{bname}_chunks", np.array(sorted(chunks), np.int64))
    for cid in sorted(chunks):
        b, p = chunks[cid], f"{tag}_b_{bname}_{cid}"
        dump(p + "_ghost", b.ghost_indices.astype(np.int64) + 1)
        dump(p + "_proj", b.projections.astype(F32))
        dump(p + "_normals", b.normals.astype(F32))
        dump(p + "_image_dist", b.image_distances.astype(F32))
        dump(p + "_ghost_dist", b.ghost_distances.astype(F32))
        dump(p + "_image_domain", b.image_domain.astype(np.int64) + 1)
        ptr, idx, w = b.image_interpolator.to_csr()
        dump_lists(p + "_donors", [idx[ptr[g]:ptr[g + 1]].astype(np.int64) + 1 for g in range(len(ptr) - 1)])
        dump_lists(p + "_weights", [w[ptr[g]:ptr[g + 1]].astype(np.float64) for g in range(len(ptr) - 1)])
X = dom.centers
u = (np.sin(F32(3) * X[:, 0]) + np.cos(F32(2) * X[:, 1])).astype(F32)
dump(tag + "_u", u)
D = np.zeros(N, F32)
dom(lambda part, u_, D_: D_.__setitem__(slice(None), JST_sensor(part, u_)), u.copy(), D)
dump(tag + "_jst", D)
for dim in range(nd):
    g, gg, ugg = np.zeros(N, F32), np.zeros(N, F32), np.zeros(N, F32)

    def f(part, u_, g_, gg_, ugg_):
        du = cell_gradient(part, u_, dim)
        g_[...] = du
        l, r = MUSCL(part, u_, du, dim, D=JST_sensor(part, u_), high_order=True)
        gg_[...] = green_gauss(part, (l + r) / F32(2), dim)
        ugg_[...] = unsigned_green_gauss(part, at_faces(part, u_, dim), dim)

    dom(f, u.copy(), g, gg, ugg)
    dump(f"{tag}_grad{dim + 1}", g)
    dump(f"{tag}_gg_muscl{dim + 1}", gg)
    dump(f"{tag}_ugg_faces{dim + 1}", ugg)
fl = cfd.Fluid()
a_inf = np.sqrt(F32(1.4) * F32(283.0) * F32(288.15))
P = np.zeros((N, nd + 2), F32)
P[:, 0] = F32(101325.0) * (1 + F32(0.02) * u / 3)
P[:, 1] = F32(288.15) * (1 + F32(0.01) * np.cos(X[:, 0]))
P[:, 2] = F32(0.5) * a_inf * (1 + F32(0.05) * np.sin(X[:, 1]))
Q = cfd.primitive2state(fl, P)
dump(tag + "_P", P)
dump(tag + "_Q", Q)
R, cf = np.zeros_like(Q), np.zeros(N, F32)
dom(euler.euler_residual(fl), Q.copy(), R, cf)
dump(tag + "_R", R)
dump(tag + "_cfl", cf)
dump("accumulator_kat", np.array([3.0, 38.0]))
json.dump(man, open(os.path.join(out, "manifest.json"), "w"))
print(f"wrote {len(man)} arrays to {out}")
