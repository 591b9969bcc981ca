"""CPU oracle: a NumPy restatement of ImmersedBoundary.jl's residual-evaluation path.

THIS PACKAGE IS TEST INFRASTRUCTURE. It is imported only by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` -- as the checker, never as the thing that is measured or shipped.
The product (``immersedboundary.jl_b200``) never imports it and has no CPU
fallback.

Parity status: **parity unpinned** for everything except ``Accumulator``.  The
reference (Julia) cannot run in this image (no ``julia`` binary, no network) and
its own test scripts contain no assertion, golden output or tolerance.  The one
known-answer vector the reference holds (``src/accumulator.jl:25-34``,
``[3.0, 38.0]``) is checked in ``tests/test_oracle_golden.py``; every other
function is pinned through analytic invariants (linear exactness, partition
independence, unit sums, round trips) listed in SURVEY.md section 8(c).

Third-party arithmetic restated here because it is not vendored in the reference:
NearestNeighbors.jl ^0.4.21 (``KDTree``/``knn``/``nn``/``inrange``: exact Euclidean
nearest neighbours; ties broken here by (squared float distance, lower index)) and
LinearAlgebra ``pinv`` (SVD pseudo-inverse, singular values below
``eps(T)*min(m,n)*smax`` dropped).

Every function cites the reference file:line it follows (paths relative to the
reference checkout root).
"""

from . import accumulator, nninterp, mesher, cfd, domain, solver, mgrid, point_implicit, euler, turbulence  # noqa: F401
