"""Oracle (TEST INFRASTRUCTURE, CPU baseline): driver of ``cpu_ref.c`` -- the compiled, OpenMP-threaded restatement of
the reference's partitioned Euler residual + ghost update (one task per partition like ``ThreadTools.tmap``,
``src/ImmersedBoundary.jl:834``; array-at-a-time operators with temporaries).

Used only by ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs and by ``tests/test_oracle_cpu_ref.py``, which
pins it bit for bit to the NumPy oracle.  The partition / boundary tables can come from the oracle's own builder
(``from_oracle``) or, for bench-sized meshes the NumPy builder is too slow for, from host-side table dumps of any
builder with the same layout (``from_builder``; bench.py passes the product's C++ builder tables, which
``tests/test_builder_parity.py`` shows to be identical)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_ref.c")
LIB = os.path.join(HERE, "_build", "libibxref.so")
F32, I64 = np.float32, np.int64
P64 = C.POINTER(C.c_int64)
PF = C.POINTER(C.c_float)


class _Part(C.Structure):
    _fields_ = [("n_dom", C.c_int64), ("n_img", C.c_int64), ("domain", P64), ("image", P64), ("image_in_domain", P64),
                ("spacing", PF), ("nf", C.c_int64 * 3), ("own", P64 * 3), ("nei", P64 * 3),
                ("lptr", P64 * 3), ("lidx", P64 * 3), ("rptr", P64 * 3), ("ridx", P64 * 3)]


class _Bdry(C.Structure):
    _fields_ = [("n_ghost", C.c_int64), ("n_img_dom", C.c_int64), ("nnz", C.c_int64), ("ghost", P64), ("image_domain", P64),
                ("ptr", P64), ("idx", P64), ("w", PF), ("normals", PF), ("eta", PF), ("normal_flow", C.c_int),
                ("n_pinf", C.c_int), ("pinf", C.c_float * 5)]


def build(force=False):
    """gcc -O2 -fopenmp -ffp-contract=off (no FMA contraction: the reference's separate roundings) -> _build/libibxref.so"""
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-std=c11", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", SRC,
                        "-o", LIB, "-lm"], check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ibxref_residual.restype = C.c_int
        _lib.ibxref_residual.argtypes = [C.c_int, C.POINTER(_Part), C.c_int, C.c_float, C.c_float, C.c_int64, PF, PF, PF, C.c_int]
        _lib.ibxref_residual_ex.restype = C.c_int
        _lib.ibxref_residual_ex.argtypes = [C.c_int, C.POINTER(_Part), C.c_int, C.c_float, C.c_float, C.c_int64, PF, PF, PF, C.c_int, C.c_int]
        _lib.ibxref_ghost_update.restype = C.c_int
        _lib.ibxref_ghost_update.argtypes = [C.c_int, C.POINTER(_Bdry), C.c_int, C.c_float, C.c_float, C.c_int64, PF, C.c_int]
    return _lib


def _i64(a):
    return np.ascontiguousarray(a, dtype=I64)


def _p64(a):
    return a.ctypes.data_as(P64)


def _pf(a):
    return a.ctypes.data_as(PF)


class CpuRef:
    """Compiled CPU restatement bound to one set of partition / boundary tables."""

    def __init__(self, nd, ncells, parts, boundaries):
        """parts: list of dicts (domain, image, image_in_domain, spacing[n_dom, nd], faces{dim: (own, nei)},
        lists{(dim, side): (ptr, idx)}); boundaries: {name: list of dicts (ghost, image_domain, ptr, idx, w,
        normals[G, nd], eta)} in chunk order."""
        self.nd, self.N = nd, ncells
        self._keep = []
        self.parts = (_Part * len(parts))()
        for s, t in zip(self.parts, parts):
            dom, img, iid = _i64(t["domain"]), _i64(t["image"]), _i64(t["image_in_domain"])
            sp = np.asfortranarray(np.asarray(t["spacing"], dtype=F32))
            self._keep += [dom, img, iid, sp]
            s.n_dom, s.n_img = dom.size, img.size
            s.domain, s.image, s.image_in_domain, s.spacing = _p64(dom), _p64(img), _p64(iid), _pf(sp)
            for d in range(nd):
                o, n = (_i64(x) for x in t["faces"][d])
                lp, li = (_i64(x) for x in t["lists"][(d, False)])
                rp, ri = (_i64(x) for x in t["lists"][(d, True)])
                self._keep += [o, n, lp, li, rp, ri]
                s.nf[d] = o.size
                s.own[d], s.nei[d] = _p64(o), _p64(n)
                s.lptr[d], s.lidx[d], s.rptr[d], s.ridx[d] = _p64(lp), _p64(li), _p64(rp), _p64(ri)
        self.bdry = {}
        for name, chunks in boundaries.items():
            arr = (_Bdry * len(chunks))()
            for s, t in zip(arr, chunks):
                g, im, p, ix = _i64(t["ghost"]), _i64(t["image_domain"]), _i64(t["ptr"]), _i64(t["idx"])
                w = np.ascontiguousarray(t["w"], dtype=F32)
                nr = np.asfortranarray(np.asarray(t["normals"], dtype=F32))
                eta = np.ascontiguousarray(t["eta"], dtype=F32)
                self._keep += [g, im, p, ix, w, nr, eta]
                s.n_ghost, s.n_img_dom, s.nnz = g.size, im.size, ix.size
                s.ghost, s.image_domain, s.ptr, s.idx, s.w, s.normals, s.eta = _p64(g), _p64(im), _p64(p), _p64(ix), _pf(w), _pf(nr), _pf(eta)
            self.bdry[name] = arr

    @classmethod
    def from_oracle(cls, dom):
        """Tables of an ``oracle.domain.Domain``."""
        parts = []
        for i in sorted(dom.partitions):
            p = dom.partitions[i]
            parts.append(dict(domain=p.domain, image=p.image, image_in_domain=p.image_in_domain, spacing=p.spacing,
                              faces=p.face_owners_neighbors, lists=p.face_lists))
        bd = {}
        for name, chunks in dom.boundaries.items():
            out = []
            for k in sorted(chunks):
                b = chunks[k]
                ptr, idx, w = b.image_interpolator.to_csr()
                out.append(dict(ghost=b.ghost_indices, image_domain=b.image_domain, ptr=ptr, idx=idx, w=w, normals=b.normals,
                                eta=b.ghost_distances / b.image_distances))
            bd[name] = out
        return cls(dom.centers.shape[1], dom.centers.shape[0], parts, bd)

    @classmethod
    def from_builder(cls, dom):
        """Tables of a domain built by the product's host-side C++ builder (``immersedboundary.jl_b200.Domain`` with
        ``build_partitions=True``; nothing here touches a GPU).  Duck-typed: this module never imports the product."""
        nd = dom.ndims
        widths = dom.cells()[1]
        parts = []
        for i in sorted(dom.partitions):
            t = dom.partitions[i].tables()
            parts.append(dict(domain=t["domain"], image=t["image"], image_in_domain=t["image_in_domain"],
                              spacing=widths[t["domain"]], faces=t["faces"], lists=t["lists"]))
        bd = {}
        for name, chunks in dom.boundaries.items():
            out = []
            for k in sorted(chunks):
                b = chunks[k]
                out.append(dict(ghost=b.ghost_indices, image_domain=b.image_domain, ptr=b.interp_ptr, idx=b.interp_idx, w=b.interp_w,
                                normals=b.normals_host, eta=b.ghost_distances / b.image_distances))
            bd[name] = out
        return cls(nd, len(dom), parts, bd)

    @classmethod
    def from_dump(cls, path):
        """Tables written by ``tools/dump_cpu_tables.py`` (a separate, untimed process runs the host-side builder, so the
        process that times this CPU path never loads the product library).  Returns (CpuRef, Q0 column-major, meta)."""
        import json
        meta = json.load(open(os.path.join(path, "meta.json")))
        ld = lambda name: np.load(os.path.join(path, name + ".npy"), mmap_mode="r")
        nd = meta["nd"]
        parts = []
        for tag in meta["parts"]:
            parts.append(dict(domain=ld(tag + "_domain"), image=ld(tag + "_image"), image_in_domain=ld(tag + "_iid"),
                              spacing=ld(tag + "_spacing").T,
                              faces={d: (ld(f"{tag}_own{d}"), ld(f"{tag}_nei{d}")) for d in range(nd)},
                              lists={(d, side): (ld(f"{tag}_ptr{d}{int(side)}"), ld(f"{tag}_idx{d}{int(side)}"))
                                     for d in range(nd) for side in (False, True)}))
        bd = {name: [dict(ghost=ld(t + "_ghost"), image_domain=ld(t + "_imdom"), ptr=ld(t + "_ptr"), idx=ld(t + "_idx"),
                          w=ld(t + "_w"), normals=ld(t + "_normals").T, eta=ld(t + "_eta")) for t in tags]
              for name, tags in meta["boundaries"].items()}
        Q = np.array(ld("Q")).T     # (N, nv) column-major, writable copy
        return cls(nd, meta["ncells"], parts, bd), Q, meta

    def residual(self, fluid, Q, R, cfl, n_threads=0, use_sensor=True):
        """``dom(f, Q, R, cfl)`` with f = the canonical Euler residual (HLL); Q, R column-major (N, nv) float32.
        ``use_sensor=False``: MUSCL without the ``D`` blend (``D = nothing``, src/ImmersedBoundary.jl:1141)."""
        assert Q.flags.f_contiguous and R.flags.f_contiguous and Q.dtype == F32 and R.dtype == F32 and cfl.dtype == F32
        rc = lib().ibxref_residual_ex(len(self.parts), self.parts, self.nd, float(fluid.R), float(fluid.gamma), self.N, _pf(Q), _pf(R),
                                      _pf(cfl), int(n_threads), int(bool(use_sensor)))
        assert rc == 0

    def ghost_update(self, fluid, Q, bcs, n_threads=0):
        """``euler_ghost_update``: for every (name, FlowBC) in order, one Jacobi ghost update of the family."""
        assert Q.flags.f_contiguous and Q.dtype == F32
        for name, bc in bcs:
            arr = self.bdry[name]
            pinf = [float(bc.p_inf), float(bc.T_inf)] + [float(x) for x in bc.u_inf]
            for s in arr:
                s.normal_flow, s.n_pinf = int(bc.normal_flow), len(pinf)
                for k, v in enumerate(pinf):
                    s.pinf[k] = v
            rc = lib().ibxref_ghost_update(len(arr), arr, self.nd, float(fluid.R), float(fluid.gamma), self.N, _pf(Q), int(n_threads))
            assert rc == 0
