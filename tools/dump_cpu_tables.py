"""Untimed helper of bench.py's CPU arms: build the sphere-octree recipe with the product's HOST-side C++ builder and
dump the partition / boundary tables + the synthetic state the CPU restatement (oracle/cpu_ref.c) runs on.

It runs as a separate process so that the process timing the reference arm never maps libibx.so (the
reference arm must not depend on the product); the tables themselves are identical to the oracle's own builder output
(tests/test_builder_parity.py, tests/test_oracle_cpu_ref.py) -- the NumPy builder is just too slow at this size.

    python tools/dump_cpu_tables.py OUTDIR LEVEL RADIUS MAX_PARTITION_SIZE [stl SUBDIV]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
F32, I64 = np.float32, np.int64


def main():
    out, level, radius, mps = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
    stl_sub = int(sys.argv[6]) if len(sys.argv) > 6 and sys.argv[5] == "stl" else None
    import immersedboundary_jl_b200 as ib
    os.makedirs(out, exist_ok=True)
    h = F32(32.0 / 2 ** level / 8 * 1.01)
    if stl_sub is None:
        surf = ib.Sphere([0, 0, 0], 0.5)
    else:
        pts, tri = ib.synthetic.icosphere(stl_sub, 0.5)
        surf = ib.Stereolitography(pts, tri)
    msh = ib.Mesh([-16, -16, -16], [32, 32, 32], ("wall", surf, h), refinement_regions=[(ib.Ball([0, 0, 0], radius), h)])
    fams = [("farfield", [(d, s) for d in range(3) for s in (False, True)])]
    dom = ib.Domain(msh, max_partition_size=mps, hypercube_families=fams, build_partitions=True, build_surfaces=False, upload=False)
    n = len(dom)
    centers, widths = dom.cells()

    def save(name, a, dtype=None):
        np.save(os.path.join(out, name + ".npy"), np.ascontiguousarray(a, dtype=dtype))

    meta = {"nd": 3, "ncells": n, "nblocks": int(msh.nblocks), "parts": [], "boundaries": {}, "level": level, "radius": radius,
            "h": float(h), "surface": "analytic sphere" if stl_sub is None else f"icosphere STL ({stl_sub} subdivisions)"}
    for i in sorted(dom.partitions):
        t = dom.partitions[i].tables()
        tag = f"p{i}"
        save(tag + "_domain", t["domain"], I64)
        save(tag + "_image", t["image"], I64)
        save(tag + "_iid", t["image_in_domain"], I64)
        save(tag + "_spacing", np.asfortranarray(widths[t["domain"]]).T, F32)   # stored (nd, n_dom) C-order == (n_dom, nd) F-order
        for d in range(3):
            o, q = t["faces"][d]
            save(f"{tag}_own{d}", o, I64)
            save(f"{tag}_nei{d}", q, I64)
            for side in (False, True):
                pp, ii = t["lists"][(d, side)]
                save(f"{tag}_ptr{d}{int(side)}", pp, I64)
                save(f"{tag}_idx{d}{int(side)}", ii, I64)
        dom.partitions[i]._tables = None
        meta["parts"].append(tag)
    for name, chunks in dom.boundaries.items():
        meta["boundaries"][name] = []
        for k in sorted(chunks):
            b = chunks[k]
            tag = f"b_{name}_{k}"
            save(tag + "_ghost", b.ghost_indices, I64)
            save(tag + "_imdom", b.image_domain, I64)
            save(tag + "_ptr", b.interp_ptr, I64)
            save(tag + "_idx", b.interp_idx, I64)
            save(tag + "_w", b.interp_w, F32)
            save(tag + "_normals", np.asfortranarray(b.normals_host).T, F32)
            save(tag + "_eta", b.ghost_distances / b.image_distances, F32)
            meta["boundaries"][name].append(tag)
    Q = ib.synthetic.primitive2state_host(ib.synthetic.euler_state(centers))
    save("Q", np.asfortranarray(Q).T, F32)      # (nv, N) C-order == (N, nv) column-major
    json.dump(meta, open(os.path.join(out, "meta.json"), "w"))


if __name__ == "__main__":
    main()
