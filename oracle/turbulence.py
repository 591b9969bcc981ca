"""Oracle (test infrastructure) restatement of ``src/turbulence.jl``: pointwise closures, float32, the reference's
operation order.  ``x ^ y`` with Float32 operands follows Julia (Float64 inside, rounded once: ``cfd.pow32``);
``log`` / ``exp`` are taken in float64 and rounded (Julia's Float32 kernels are accurate to < 1 ulp).
Velocity-gradient arguments are nested lists ``g[i][j]`` = d u_i / d x_j of vectors."""
import numpy as np

from .cfd import pow32

F32 = np.float32
EPS = np.finfo(F32).eps


def _log(x):
    return np.log(x.astype(np.float64)).astype(F32)


def _exp(x):
    return np.exp(x.astype(np.float64)).astype(F32)


def von_karman(yp, kappa=F32(0.41), C=F32(4.9)):
    """``von_Karman``, ``src/turbulence.jl:11-16``."""
    return np.minimum(_log(np.maximum(yp, F32(1.0))) / kappa + C, yp)


def wall_function_rey(Rey, kappa=F32(0.41), C=F32(4.9), A=F32(19.0), beta=F32(0.075), beta_star=F32(0.09), D=F32(4.2),
                      A_plus=F32(360.0), omega=F32(0.5), n_iter=20):
    """``wall_function(Rey)``, ``src/turbulence.jl:27-72`` -> dict y+, u+, mu+, k+, du+/dy+."""
    Rey = np.maximum(np.abs(np.asarray(Rey, dtype=F32)), EPS)
    yp = np.sqrt(Rey)
    up = np.empty_like(yp)
    for _ in range(n_iter):
        up = von_karman(yp, kappa, C)
        yp = omega * (Rey / up) + (F32(1.0) - omega) * yp
    up = Rey / yp
    t = F32(1.0) - _exp(-yp / A)
    mup = kappa * yp * (t * t)
    dudy = F32(1.0) / (F32(1.0) + mup)
    kp = np.minimum(yp * yp / (F32(6.0) * beta_star / beta - F32(2.0)), D * _exp(-yp / A_plus))
    return {"y_plus": yp, "u_plus": up, "mu_plus": mup, "k_plus": kp, "dudy_plus": dudy}


def wall_function(y, u, nu, beta_star=F32(0.09), **kw):
    """``wall_function(y, u, nu)``, ``src/turbulence.jl:74-98`` -> dict u_tau, nu_t, k, omega, eps, dudn."""
    nt = wall_function_rey(u * y / nu, beta_star=beta_star, **kw)
    ut = u / nt["u_plus"]
    nut = nt["mu_plus"] * nu
    k = nt["k_plus"] * (ut * ut)
    om = k / nut
    return {"u_tau": ut, "nu_t": nut, "k": k, "omega": om, "eps": beta_star * om * k,
            "dudn": nt["dudy_plus"] * (ut * ut) / nu}


def shear_rate(g):
    """``shear_rate``, ``src/turbulence.jl:110-124``: sqrt(2 S_ij S_ij)."""
    s = np.zeros_like(g[0][0])
    for i in range(len(g)):
        for j in range(len(g)):
            e = (g[i][j] + g[j][i]) / F32(2)
            s = s + e * e
    return np.sqrt(F32(2) * s)


def smagorinsky(Delta, S, Cs=F32(0.17)):
    """``Smagorinsky_νSGS``, ``src/turbulence.jl:134-137``."""
    t = Cs * Delta
    return t * t * S


def standard_keps(k, eps, S, Cmu=F32(0.09), sigma_k=F32(1.0), sigma_eps=F32(1.3), C1=F32(1.44), C2=F32(1.92)):
    """``standard_kϵ``, ``src/turbulence.jl:175-194``."""
    nut = Cmu * (k * k) / eps
    Pk = nut * (S * S)
    return {"nu_k": nut / sigma_k, "nu_eps": nut / sigma_eps, "Sk": Pk - eps,
            "Seps": C1 * Pk * eps / k - C2 * (eps * eps) / k, "nu_t": nut}


def wray_agarwal(R, S, gradR, gradS, sigma_R=F32(0.72), C1=F32(0.0829), kappa=F32(0.41)):
    """``Wray_Agarwal``, ``src/turbulence.jl:222-241``; gradients are (N, nd)."""
    C2 = sigma_R + C1 / (kappa * kappa)
    dot = gradR[:, 0] * gradS[:, 0]
    for d in range(1, gradR.shape[1]):
        dot = dot + gradR[:, d] * gradS[:, d]
    src = C1 * R * S + C2 * dot * (R / (S + EPS))
    return {"nu_t": R, "nu_R": R * sigma_R, "S": np.minimum(src, F32(10.0) * R)}


def ducros_sensor(g):
    """``Ducros_sensor``, ``src/turbulence.jl:253-283``."""
    nd = len(g)
    div2 = np.zeros_like(g[0][0])
    for i in range(nd):
        div2 = div2 + g[i][i]
    div2 = div2 * div2
    if nd == 2:
        curl2 = (g[1][0] - g[0][1]) ** 2
    elif nd == 3:
        curl2 = (g[2][1] - g[1][2]) ** 2 + (g[0][2] - g[2][0]) ** 2 + (g[1][0] - g[0][1]) ** 2
    else:
        raise ValueError("Ducros sensor only implemented for 2D and 3D")
    return (div2 + EPS) / (div2 + curl2 + EPS)


def wale(Delta, g, Cw=F32(0.325)):
    """``WALE_νSGS``, ``src/turbulence.jl:292-337`` (3-D).  ``δ / 3`` is a Float64 scalar in the reference, so the
    S^d_ij terms are Float64 and the Float32 running sum is rounded once per term."""
    assert len(g) == 3, "WALE model only implemented for 3D"
    g2 = [[None] * 3 for _ in range(3)]
    for i in range(3):
        for j in range(3):
            s = np.zeros_like(g[i][j])
            for k in range(3):
                s = s + g[i][k] * g[k][j]
            g2[i][j] = s
    SS = np.zeros_like(g[0][0])
    for i in range(3):
        for j in range(3):
            e = (g[i][j] + g[j][i]) / F32(2)
            SS = SS + e * e
    SD = np.zeros_like(g[0][0])
    for i in range(3):
        for j in range(3):
            term = ((g2[i][j] + g2[j][i]) / F32(2)).astype(np.float64) - g2[i][j].astype(np.float64) * ((1.0 if i == j else 0.0) / 3.0)
            SD = (SD.astype(np.float64) + term * term).astype(F32)
    return Cw * (Delta * Delta) * pow32(SD, F32(1.5)) / (pow32(SS, F32(2.5)) + pow32(SD, F32(1.25)) + EPS)
