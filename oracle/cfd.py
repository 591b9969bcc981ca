"""Oracle (test infrastructure) restatement of the hot-path subset of ``src/cfd.jl``.

Arrays are (points, nv) with columns ``[p T u v (w)]`` (primitive) or
``[rho E rho*u ...]`` (state), float32 unless noted.  Type promotions of the
reference are kept: HLL returns float64.
"""
import numpy as np

F32 = np.float32


class Fluid:
    """``Fluid``, ``src/cfd.jl:14-53`` (defaults: R = 283, gamma = 1.4)."""

    def __init__(self, R=F32(283.0), gamma=F32(1.4), k=(F32(0.00646), F32(6.468e-5)), mu_ref=F32(1.716e-5),
                 T_ref=F32(273.15), S=F32(110.4)):
        self.R = F32(R)
        self.gamma = F32(gamma)
        self.k = [F32(x) for x in (k if np.ndim(k) else [k])]
        self.mu_ref = F32(mu_ref)
        self.T_ref = F32(T_ref)
        self.S = F32(S)


def _clampT(T):
    return np.maximum(T, F32(10.0))  # clamp(T, 10f0, Inf32)


def speed_of_sound(fld, T):
    """``speed_of_sound``, ``src/cfd.jl:62-64``: sqrt((gamma * R) * clamp(T))."""
    return np.sqrt(fld.gamma * fld.R * _clampT(T))


def pow32(x, y):
    """Float32 ``x ^ y`` as Julia evaluates it: in Float64 through exp2(log2|x| * y), rounded once
    (``Base.Math.pow_body`` for Float16/Float32); integer exponents by repeated Float64 multiplication."""
    x = np.asarray(x, dtype=F32).astype(np.float64)
    if isinstance(y, (int, np.integer)):
        r = np.ones_like(x)
        for _ in range(int(y)):
            r = r * x
        return r.astype(F32)
    with np.errstate(divide="ignore"):
        return np.exp2(np.log2(np.abs(x)) * np.float64(F32(y))).astype(F32)


def dynamic_viscosity(fld, T):
    """``dynamic_viscosity``, ``src/cfd.jl:71-77`` (note the exponent 2/3 of the reference)."""
    T = _clampT(T)
    return fld.mu_ref * pow32(T / fld.T_ref, F32(2.0) / F32(3)) * (fld.T_ref + fld.S) / (T + fld.S)


def heat_conductivity(fld, T):
    """``heat_conductivity``, ``src/cfd.jl:84-90``."""
    k = F32(0) * T
    for i, ki in enumerate(fld.k):
        k = k + ki * pow32(T, i)
    return k


def pressure_coefficient(fld, p, p_inf, M_inf):
    """``pressure_coefficient``, ``src/cfd.jl:420-426``."""
    return F32(2) * (p / F32(p_inf) - F32(1.0)) / (F32(M_inf) * F32(M_inf) * fld.gamma)


def reynolds_number(fld, P_inf, L_ref):
    """``Reynolds_number``, ``src/cfd.jl:626-637``."""
    P_inf = np.asarray(P_inf, dtype=F32)
    V = F32(np.sqrt(np.sum(P_inf[2:].astype(np.float64) ** 2)))
    rho = P_inf[0] / (fld.R * P_inf[1])
    mu = dynamic_viscosity(fld, P_inf[1:2])[0]
    return V * F32(L_ref) * rho / mu


def adjust_reynolds(fld, P_inf, L_ref, Re):
    """``adjust_Reynolds``, ``src/cfd.jl:645-654``: a new fluid whose reference viscosity gives ``Re``."""
    mu_ref = fld.mu_ref * reynolds_number(fld, P_inf, L_ref) / F32(Re)
    return Fluid(fld.R, fld.gamma, fld.k, mu_ref, fld.T_ref, fld.S)


def _half_sq(u):
    acc = u[:, 0] ** 2
    for d in range(1, u.shape[1]):
        acc = acc + u[:, d] ** 2
    return acc / F32(2)


def primitive2state(fld, P):
    """``primitive2state``, ``src/cfd.jl:106-123``."""
    p = P[:, 0]
    T = _clampT(P[:, 1])
    u = P[:, 2:]
    k = _half_sq(u)
    rho = p / (fld.R * T)
    E = rho * (fld.R / (fld.gamma - F32(1.0)) * T + k)
    return np.concatenate([rho[:, None], E[:, None], rho[:, None] * u], axis=1)


def state2primitive(fld, Q):
    """``state2primitive``, ``src/cfd.jl:137-151``."""
    rho = Q[:, 0]
    E = Q[:, 1]
    u = Q[:, 2:] / rho[:, None]
    k = _half_sq(u)
    p = (fld.gamma - F32(1.0)) * (E - rho * k)
    T = _clampT(p / (rho * fld.R))
    return np.concatenate([p[:, None], T[:, None], u], axis=1)


class FlowBC:
    """``FlowBC``, ``src/cfd.jl:160-300``."""

    def __init__(self, fluid, P, normal_flow=False):
        P = np.asarray(P, dtype=F32)
        self.fluid = fluid
        self.p_inf = P[0]
        self.T_inf = P[1]
        self.u_inf = P[2:].copy()
        self.normal_flow = normal_flow

    def __call__(self, P, normals, image_distances=None, du_dn=None, transpiration=F32(0.0)):
        p, T, u = P[:, 0], P[:, 1], P[:, 2:]
        if self.normal_flow:
            assert self.u_inf.size == 1, "Only 3 parcels in P (p, T and normal flow) allowed for normal_flow = true BC"
            un = np.full(P.shape[0], self.u_inf[0], dtype=P.dtype)
        else:
            un = normals[:, 0] * self.u_inf[0]
            for d in range(1, normals.shape[1]):  # matrix-vector product, left to right
                un = un + normals[:, d] * self.u_inf[d]
        cur = (u * normals)[:, 0]
        for d in range(1, normals.shape[1]):
            cur = cur + (u * normals)[:, d]
        a = speed_of_sound(self.fluid, T)
        M = np.abs(un) / a
        sup, sub = (M > 1.0).astype(F32), (M <= 1.0).astype(F32)
        pb = (un >= 0.0) * (sup * self.p_inf + sub * p) + (un < 0.0) * (sup * p + sub * self.p_inf)
        Tb = (un > 0.0) * self.T_inf + (un <= 0.0) * T
        if self.normal_flow:
            ub = u + normals * (un - cur + transpiration)[:, None]
        else:
            ub = (un < 0.0)[:, None] * u + (un >= 0.0)[:, None] * self.u_inf[None, :]
        if (du_dn is None) != (image_distances is None):
            raise ValueError("du!dn and image_distances must be passed together for BC imposition")
        if du_dn is not None:
            eps = np.finfo(ub.dtype).eps
            V = np.sqrt(np.sum(ub ** 2, axis=1)) + eps
            ub = ub * ((V - du_dn * image_distances) / V)[:, None]
        return np.concatenate([pb[:, None], Tb[:, None], ub], axis=1).astype(P.dtype)


def _flux_side(fld, P, dim):
    Q = primitive2state(fld, P)
    F = Q.copy()
    p = P[:, 0]
    F[:, 1] = F[:, 1] + p
    if np.ndim(dim) == 0:
        un = P[:, 2 + dim]
    else:
        un = np.sum(dim * P[:, 2:], axis=1)
    a = speed_of_sound(fld, P[:, 1])
    F = F * un[:, None]
    if np.ndim(dim) == 0:
        F[:, 2 + dim] = F[:, 2 + dim] + p
    else:
        F[:, 2:] = F[:, 2:] + p[:, None] * dim
    return Q, F, un, a


def inviscid_fluxes_hll(fld, PL, PR, dim):
    """HLL ``inviscid_fluxes``, ``src/cfd.jl:459-508``; ``dim`` 0-based int or (faces, nd) normals.

    Returns float64 (the ``0.0`` literals at :504-505 promote).  The wave-speed
    estimates deliberately use the reference's (opposite-state) formula.
    """
    QL, FL, uL, aL = _flux_side(fld, PL, dim)
    QR, FR, uR, aR = _flux_side(fld, PR, dim)
    SR = np.minimum((uR - aR).astype(np.float64), 0.0)[:, None]
    SL = np.maximum((uL + aL).astype(np.float64), 0.0)[:, None]
    with np.errstate(invalid="ignore", divide="ignore"):
        return (SL * FL - SR * FR + SR * SL * (QR - QL)) / (SL - SR)


def inviscid_fluxes_sensor(fld, PL, PR, nuL, nuR, dim):
    """Sensor-Rusanov ``inviscid_fluxes``, ``src/cfd.jl:516-554``."""
    UL = primitive2state(fld, PL)
    UL[:, 1] = UL[:, 1] + PL[:, 0]
    UR = primitive2state(fld, PR)
    UR[:, 1] = UR[:, 1] + PR[:, 0]
    P = (PL + PR) / F32(2)
    p, T = P[:, 0], P[:, 1]
    u = P[:, 2 + dim] if np.ndim(dim) == 0 else np.sum(dim * P[:, 2:], axis=1)
    a = speed_of_sound(fld, T)
    F = (UL + UR) * u[:, None] / F32(2)
    if np.ndim(dim) == 0:
        F[:, 2 + dim] = F[:, 2 + dim] + p
    else:
        F[:, 2:] = F[:, 2:] + p[:, None] * dim
    nu = np.maximum(nuL, nuR)
    if nu.ndim == 1:
        nu = nu[:, None]
    F = F + (UL - UR) * (nu * (a + np.abs(u))[:, None] / F32(2))
    return F


def jst_sensor_3pt(Pim1, Pi, Pip1):
    """Pointwise ``JST_sensor``, ``src/cfd.jl:563-573``."""
    e = F32(1e-14)
    return (np.abs(Pim1 + Pip1 - 2 * Pi) + e) / (np.abs(Pim1 - Pi) + np.abs(Pip1 - Pi) + e)


def shock_sensor(vg):
    """``shock_sensor``, ``src/cfd.jl:589-617``; ``vg[i][j]`` = d u_i / d x_j (vectors).  In 2-D the reference's
    cyclic loop visits the single vorticity component twice; kept."""
    e = F32(1e-14)
    nd = len(vg)
    vort = np.zeros_like(vg[0][0])
    divu = np.zeros_like(vg[0][0])
    for i in range(nd):
        i1 = (i + 1) % nd
        i2 = (i1 + 1) % nd
        divu = divu + vg[i][i]
        vort = vort + (vg[i2][i1] - vg[i1][i2]) ** 2
    divu = divu * divu
    return (divu + e) / (divu + vort + e)


def viscous_fluxes(fld, P, Pgrad, dim, mu_t=F32(0.0)):
    """``viscous_fluxes``, ``src/cfd.jl:664-736``: ``dim`` a 0-based axis or an (N, nd) direction matrix."""
    T = P[:, 1]
    mu = dynamic_viscosity(fld, T) + mu_t
    k = heat_conductivity(fld, T)
    nd = P.shape[1] - 2
    vg = lambda i, j: Pgrad[j][:, 2 + i]
    divu = np.zeros_like(T)
    for i in range(nd):
        divu = divu + vg(i, i)
    tau = lambda i, j: ((vg(i, j) + vg(j, i)) - (F32(2.0) / F32(3) if i == j else F32(0.0)) * divu) * mu
    f = lambda i: Pgrad[i][:, 1] * k
    F = np.zeros_like(P)
    if np.ndim(dim) == 0:
        F[:, 1] = F[:, 1] + f(dim)
        for j in range(nd):
            F[:, 1] = F[:, 1] + tau(dim, j) * P[:, 2 + j]
        for j in range(nd):
            F[:, 2 + j] = F[:, 2 + j] + tau(dim, j)
    else:
        taud = []
        for i in range(nd):
            s = np.zeros_like(T)
            for j in range(nd):
                s = s + tau(i, j) * dim[:, j]
            taud.append(s)
        for j in range(nd):
            F[:, 1] = F[:, 1] + f(j) * dim[:, j]
            F[:, 1] = F[:, 1] + taud[j] * P[:, 2 + j]
        for j in range(nd):
            F[:, 2 + j] = F[:, 2 + j] + taud[j]
    return F
