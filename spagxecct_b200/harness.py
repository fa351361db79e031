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
