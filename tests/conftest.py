import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

F32 = np.float32


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


def _ensure_built():
    import __graft_entry__ as g
    lib = os.path.join(ROOT, "immersedboundary.jl_b200", "libibx.so")
    if not os.path.exists(lib):
        g.build()


@pytest.fixture(scope="session")
def ib():
    _ensure_built()
    import immersedboundary_jl_b200 as ib
    return ib


@pytest.fixture(scope="session")
def oracle():
    import oracle
    return oracle


RAE = os.path.join(ROOT, "tests", "golden", "rae2822.dat")


class Case:
    """One configuration built twice: through the product (C++ builder) and through the oracle."""

    def __init__(self, name, ib, oracle, mps=100_000, upload=False):
        self.name = name
        M, OM = ib, oracle.mesher
        if name in ("advection", "dissipation"):
            h = F32(1e-2) if name == "advection" else F32(2e-2)
            seg = lambda mod, a, b: mod.Stereolitography(np.array([a, b], dtype=np.float64))
            regs = lambda mod: [(mod.Line([0.0, 0.0], [1.0, 1.0]), F32(2) * h), (mod.Line([0.0, 0.0], [0.5, 0.5]), h)]
            self.fams = [("outlet" if name == "advection" else "neumann", [(0, True), (1, True)])]
            self.msh = M.Mesh([0.0, 0.0], [1.0, 1.0], ("lower", seg(M, [0., 0.], [1., 0.]), h),
                              ("upper", seg(M, [0., 0.], [0., 1.]), h), refinement_regions=regs(M))
            self.omsh = OM.Mesh([0.0, 0.0], [1.0, 1.0], ("lower", seg(OM, [0., 0.], [1., 0.]), h),
                                ("upper", seg(OM, [0., 0.], [0., 1.]), h), refinement_regions=regs(OM))
        elif name == "rae2822":
            self.fams = [("farfield", [(0, False), (0, True), (1, False), (1, True)])]
            stl = M.merge_points(M.Stereolitography(RAE))
            feat = M.DistanceField(M.feature_regions(stl, radius=0.05))
            self.msh = M.Mesh(np.array([-25, -25], F32), np.array([50, 50], F32), ("wall", stl, F32(1e-2)),
                              refinement_regions=[(feat, F32(5e-3))])
            ostl = OM.merge_points(OM.Stereolitography(RAE))
            ofeat = OM.DistanceField(OM.feature_regions(ostl, radius=0.05))
            self.omsh = OM.Mesh(np.array([-25, -25], F32), np.array([50, 50], F32), ("wall", ostl, F32(1e-2)),
                                refinement_regions=[(ofeat, F32(5e-3))])
        elif name == "sphere3d_np2":
            # root box widths that are NOT powers of two: the kernels' exact power-of-two shortcuts must switch off
            self.fams = [("farfield", [(d, s) for d in range(3) for s in (False, True)])]
            self.msh = M.Mesh([-1.5, -1.5, -1.5], [3, 3, 3], ("wall", M.Sphere([0, 0, 0], 0.4), F32(0.05)),
                              refinement_regions=[(M.Ball([0, 0, 0], 0.7), F32(0.1))])
            self.omsh = OM.Mesh([-1.5, -1.5, -1.5], [3, 3, 3], ("wall", OM.AnalyticSphere([0, 0, 0], 0.4), F32(0.05)),
                                refinement_regions=[(OM.Ball([0, 0, 0], 0.7), F32(0.1))])
        elif name in ("sphere3d", "sphere3d_stl"):
            self.fams = [("farfield", [(d, s) for d in range(3) for s in (False, True)])]
            if name == "sphere3d":
                surf, osurf, h = M.Sphere([0, 0, 0], 0.5), OM.AnalyticSphere([0, 0, 0], 0.5), F32(0.06)
            else:
                pts, tri = ib.synthetic.icosphere(1, 0.5)
                surf, osurf, h = M.Stereolitography(pts, tri), OM.Stereolitography(pts, tri), F32(0.12)
            self.msh = M.Mesh([-2, -2, -2], [4, 4, 4], ("wall", surf, h), refinement_regions=[(M.Ball([0, 0, 0], 0.9), F32(0.12))])
            self.omsh = OM.Mesh([-2, -2, -2], [4, 4, 4], ("wall", osurf, h),
                                refinement_regions=[(OM.Ball([0, 0, 0], 0.9), F32(0.12))])
        else:
            raise KeyError(name)
        self.dom = ib.Domain(self.msh, max_partition_size=mps, hypercube_families=self.fams, upload=upload)
        self.odom = oracle.domain.Domain(self.omsh, max_partition_size=mps, hypercube_families=self.fams)


_cache = {}


@pytest.fixture(scope="session")
def get_case(ib, oracle):
    def get(name, mps=100_000, upload=False):
        key = (name, mps, upload)
        if key not in _cache:
            _cache[key] = Case(name, ib, oracle, mps, upload)
        return _cache[key]
    return get
