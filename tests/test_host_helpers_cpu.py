"""Host-side helpers of the Python mirror that need no GPU."""
import numpy as np

F32 = np.float32


def test_inside_polygon_even_odd(ib):
    sq = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    pts = np.array([[0.5, 0.5], [1.5, 0.5], [-0.1, 0.2], [0.99, 0.01], [0.5, 1.01]], F32)
    assert ib.synthetic.inside_polygon(sq, pts).tolist() == [True, False, False, True, False]
    # a concave "L": the notch is outside
    L = np.array([[0, 0], [2, 0], [2, 1], [1, 1], [1, 2], [0, 2]], float)
    pts = np.array([[0.5, 0.5], [1.5, 0.5], [1.5, 1.5], [0.5, 1.5]], F32)
    assert ib.synthetic.inside_polygon(L, pts).tolist() == [True, True, False, True]


def test_inside_polygon_on_the_rae2822_mesh(get_case, ib):
    """The cells started at rest by the C3 march are the ones enclosed by the airfoil: every one of them lies within the
    chord and the thickness of the section, and their total area is the section's area (shoelace) to within the cell size."""
    import os
    c = get_case("rae2822", 10_000)
    poly = np.loadtxt(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rae2822.dat"))
    cen, wid = c.dom.cells()
    inside = ib.synthetic.inside_polygon(poly, cen)
    assert 200 < inside.sum() < len(cen) // 4
    assert cen[inside, 0].min() > 0 and cen[inside, 0].max() < 1 and np.abs(cen[inside, 1]).max() < 0.07
    x, y = poly[:, 0], poly[:, 1]
    area = 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
    assert abs((wid[inside, 0] * wid[inside, 1]).sum() - area) < 0.05 * area


def test_multistage_coefficients(ib):
    for n, a in ib.RK_STAGES.items():
        assert len(a) == n and a[-1] == 1.0 and all(0 < x <= 1 for x in a)
