"""Host-side helpers for tests and bench: a stand-in for Step 1 (SPA_G_Get_Resid,
R/SPAGxECCT.R:48-96, which stays in R in a real deployment) and synthetic-data recipes.
numpy only; nothing here is on the Step-2 hot path."""
import numpy as np


def logistic_fit(X, y, maxit=25, eps=1e-8):
    """glm.fit(binomial) style IRLS; X includes the intercept column. Returns (beta, mu)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    mu = (y + 0.5) / 2
    eta = np.log(mu / (1 - mu))
    dev_old = _dev(y, mu)
    beta = np.zeros(X.shape[1])
    for _ in range(maxit):
        w = mu * (1 - mu)
        z = eta + (y - mu) / w
        sw = np.sqrt(w)
        beta, *_ = np.linalg.lstsq(X * sw[:, None], z * sw, rcond=None)
        eta = X @ beta
        mu = 1 / (1 + np.exp(-eta))
        dev = _dev(y, mu)
        if abs(dev - dev_old) / (abs(dev) + 0.1) < eps:
            break
        dev_old = dev
    return beta, mu


def _dev(y, mu):
    with np.errstate(divide="ignore", invalid="ignore"):
        a = np.where(y > 0, y * np.log(y / mu), 0.0)
        b = np.where(y < 1, (1 - y) * np.log((1 - y) / (1 - mu)), 0.0)
    return 2 * float(np.sum(a + b))


def get_resid(traits, y, covars):
    """Residuals of the genotype-independent model: y - mu (binary) or OLS residuals."""
    n = len(y)
    X = np.column_stack([np.ones(n)] + [np.asarray(c, dtype=np.float64) for c in covars])
    if traits == "binary":
        _, mu = logistic_fit(X, y)
        return np.asarray(y, dtype=np.float64) - mu
    beta, *_ = np.linalg.lstsq(X, np.asarray(y, dtype=np.float64), rcond=None)
    return np.asarray(y, dtype=np.float64) - X @ beta


def cox_fit_breslow(time, event, X, maxit=50, tol=1e-10):
    """Cox proportional-hazards fit by Newton-Raphson with Breslow's treatment of ties (stand-in for the Step-1
    `coxph` of SPA_G_Get_Resid, R/SPAGxECCT.R:72-80, used only to give test / bench residuals the right shape).
    Returns (beta, martingale residuals)."""
    time = np.asarray(time, dtype=np.float64); event = np.asarray(event, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64).reshape(len(time), -1)
    Xc = X - X.mean(axis=0)
    order = np.argsort(-time, kind="stable")                # descending: risk set of i = prefix ending at i's tie group
    t = time[order]; d = event[order]; Z = Xc[order]
    # index of the last element of each tie group (in descending order) for every position
    last = np.r_[np.nonzero(np.diff(t))[0], len(t) - 1]
    grp_end = last[np.searchsorted(last, np.arange(len(t)))]
    beta = np.zeros(Z.shape[1])
    for _ in range(maxit):
        r = np.exp(Z @ beta)
        S0 = np.cumsum(r)[grp_end]
        S1 = np.cumsum(Z * r[:, None], axis=0)[grp_end]
        ZZ = Z[:, :, None] * Z[:, None, :] * r[:, None, None]
        S2 = np.cumsum(ZZ, axis=0)[grp_end]
        m1 = S1 / S0[:, None]
        grad = (d[:, None] * (Z - m1)).sum(axis=0)
        hess = (d[:, None, None] * (S2 / S0[:, None, None] - m1[:, :, None] * m1[:, None, :])).sum(axis=0)
        step = np.linalg.solve(hess, grad)
        beta = beta + step
        if np.max(np.abs(step)) < tol:
            break
    r = np.exp(Z @ beta)
    S0 = np.cumsum(r)[grp_end]
    # cumulative baseline hazard at each subject's time: sum over events with time <= t_i, i.e. a reversed cumsum
    inc = d / S0
    H = np.cumsum(inc[::-1])[::-1]
    first = np.r_[0, np.nonzero(np.diff(t))[0] + 1]
    grp_start = first[np.searchsorted(first, np.arange(len(t)), side="right") - 1]
    H = H[grp_start]
    mart_sorted = d - r * H
    mart = np.empty_like(mart_sorted)
    mart[order] = mart_sorted
    return beta, mart


def cox_martingale_resid(time, event, X):
    return cox_fit_breslow(time, event, X)[1]
