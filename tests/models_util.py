"""Small seeded instances of the five model families, as (engine model, oracle model) pairs."""
import numpy as np

from oracle import densities as od
from pymc3_b200 import model as pm


def glm_data(n, k, seed=0, alpha=0.3):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, k)).astype("f4")
    beta = rng.normal(0, 0.5, size=k)
    p = 1.0 / (1.0 + np.exp(-(alpha + X.astype("f8") @ beta)))
    y = (rng.random(n) < p).astype("f4")
    return X, y


def hier_data(n, g, seed=0):
    rng = np.random.default_rng(seed)
    w = rng.random(g) + 0.1
    idx = rng.choice(g, size=n, p=w / w.sum())
    floor = (rng.random(n) < 0.17).astype(np.uint8)
    a = 1.5 + 0.3 * rng.normal(size=g)
    b = -0.7 + 0.3 * rng.normal(size=g)
    y = (a[idx] + b[idx] * floor + 0.7 * rng.normal(size=n)).astype("f4")
    return idx, floor, y


def pairs(seed=3, glm_n=300, glm_k=7, hier_n=500, hier_g=6, sv_t=60):
    X, y = glm_data(glm_n, glm_k, seed)
    idx, floor, yy = hier_data(hier_n, hier_g, seed)
    ret = pm.sp500_log_returns()[:sv_t]
    sig = np.arange(1, 6, dtype="f8")
    return {
        "std_normal": (pm.StdNormal(5, sigma=sig), od.StdNormal(5, sig)),
        "eight_schools": (pm.EightSchoolsNCP(), od.EightSchoolsNCP()),
        "glm": (pm.LogisticGLM(X, y), od.LogisticGLM(X, y)),
        "hier": (pm.HierLinearNCP(idx, floor, yy, hier_g), od.HierLinearNCP(idx, floor, yy, hier_g)),
        "stoch_vol": (pm.StochVol(ret), od.StochVol(ret)),
    }
