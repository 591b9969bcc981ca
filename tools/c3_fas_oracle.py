"""CPU experiment (oracle side): the reference's `FAS!` (src/solver.jl:39-91) with the multigrid(dom) hierarchy as the
driver of the RAE2822 Euler case -- f(l, Q) = (R CFL / cfl on the cells that are not ghosts, 1) on every level, compiled
restatement oracle/cpu_ref.c on the host tables of every level, IDW transfer operators of multigrid(dom).

    python tools/c3_fas_oracle.py LEVELS CYCLES N_ITER CFL        e.g.  3 1000 3 0.4

Finding: with 2 or 3 levels the cycle diverges after 240 - 330 cycles for CFL 0.2 - 0.4; with one
level (plain forward-Euler smoothing) it converges like tools/c3_converge.py --stages 1."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import immersedboundary_jl_b200 as ib
from oracle import cfd, cpu_ref
F32 = np.float32
RAE = os.path.join(ROOT, "tests", "golden", "rae2822.dat")
NLEV = int(sys.argv[1]); NCYC = int(sys.argv[2]); NIT = int(sys.argv[3]); CFL = F32(float(sys.argv[4]))
stl = ib.merge_points(ib.Stereolitography(RAE))
feat = ib.DistanceField(ib.feature_regions(stl, radius=0.05))
msh = ib.Mesh(np.array([-25, -25], F32), np.array([50, 50], F32), ("wall", stl, F32(1e-2)), refinement_regions=[(feat, F32(5e-3))])
fams = [("farfield", [(0, False), (0, True), (1, False), (1, True)])]
dom = ib.Domain(msh, max_partition_size=10_000, hypercube_families=fams, upload=False)
cd, pro, coa = ib.multigrid(dom)
doms = ([dom] + cd)[:NLEV]
def csr(acc):
    p, i, w = acc.tables()
    import scipy.sparse as sp
    n = len(p) - 1
    return sp.csr_matrix((w if w is not None else np.ones(len(i), F32), i, p), shape=(n, int(i.max()) + 1))
Cs = [csr(a) for a in coa[:NLEV - 1]]; Ps = [csr(a) for a in pro[:NLEV - 1]]
refs = [cpu_ref.CpuRef.from_builder(d) for d in doms]
fl = cfd.Fluid()
a_inf = np.sqrt(1.4 * 283.0 * 288.15); al = np.radians(2.31)
Pinf = np.array([101325.0, 288.15, 0.73 * a_inf * np.cos(al), 0.73 * a_inf * np.sin(al)], F32)
wall = np.array([101325.0, 288.15, 0.0], F32)
bcs = [("wall", cfd.FlowBC(fl, wall, normal_flow=True)), ("farfield", cfd.FlowBC(fl, Pinf))]
lives = []
for d in doms:
    g = np.zeros(len(d), bool)
    for ch in d.boundaries.values():
        for b in ch.values(): g[b.ghost_indices] = True
    lives.append((~g).astype(F32)[:, None])
print("levels", [len(d) for d in doms], flush=True)
nev = [0] * NLEV
def f(l, Q):
    Q = np.asfortranarray(Q)
    refs[l].ghost_update(fl, Q, bcs, 8)          # in place on the level's state
    R = np.zeros((len(Q), 4), F32, order="F"); cf = np.zeros(len(Q), F32)
    refs[l].residual(fl, Q, R, cf, 8)
    nev[l] += 1
    return (R * (CFL / cf)[:, None] * lives[l]).astype(F32), Q
def FAS(Q, l, prescribed=None):
    fQ, Q = f(l, Q)
    source = None if prescribed is None else prescribed - fQ
    r = fQ if source is None else fQ + source
    if l < NLEV - 1:
        Qc = (Cs[l] @ Q).astype(F32); Qcold = Qc.copy()
        Qc = FAS(Qc, l + 1, (Cs[l] @ r).astype(F32))
        Q = (Q + (Ps[l] @ (Qc - Qcold)) * lives[l]).astype(F32)
    for _ in range(NIT):
        r, Q = f(l, Q)
        if source is not None: r = r + source
        Q = (Q + r).astype(F32)
    return Q
s = dom.surfaces["wall"]
def coeffs(Qs):
    P = cfd.state2primitive(fl, Qs)
    cp = cfd.pressure_coefficient(fl, P[:, 0], Pinf[0], 0.73)
    cps = np.array([(cp[s.idx[a:b]] * s.w[a:b]).sum() for a, b in zip(s.ptr[:-1], s.ptr[1:])], F32)
    Fxy = (cps[:, None] * s.normals * s.areas[:, None]).sum(axis=0)
    return float(-Fxy[0] * np.sin(al) + Fxy[1] * np.cos(al)), float(Fxy[0] * np.cos(al) + Fxy[1] * np.sin(al))
P0 = np.tile(Pinf, (len(dom), 1)); P0[ib.synthetic.inside_polygon(np.loadtxt(RAE), dom.cells()[0]), 2:] = 0
Q = np.asfortranarray(ib.synthetic.primitive2state_host(P0))
t0 = time.time()
for cyc in range(NCYC + 1):
    Q = FAS(Q, 0)
    if not np.isfinite(Q).all(): print("diverged at cycle", cyc); break
    if cyc % max(NCYC // 20, 1) == 0:
        r, _ = f(0, Q.copy()); nev[0] -= 1
        print(f"cycle {cyc:6d} fine evals {nev[0]:7d} coarse {nev[1:]}  |r_rho| {np.linalg.norm(r[:,0]):.4e}  Cl {coeffs(Q)[0]:+.6f} Cd {coeffs(Q)[1]:+.6f}  t={time.time()-t0:.0f}s", flush=True)
