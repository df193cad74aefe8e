"""pm.glm.GLM front-end (reference: pymc3/glm/linear.py:29-160, glm/utils.py:20-120) -- host logic only."""
import numpy as np
import pandas as pd
import pytest

import pymc3_b200 as pm
from pymc3_b200 import _capi


def _data(n=50, k=3, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, k))
    y = (rng.random(n) < 0.5).astype("f8")
    return X, y


def test_defaults_follow_the_reference():
    X, y = _data()
    m = pm.glm.GLM(X, y, family="binomial")
    assert m.free_RVs == ["Intercept", "x0", "x1", "x2"]            # glm/linear.py:72-99 creation order
    assert m.prior_tau == 1e-6                                       # glm/linear.py:49
    assert m.family == _capi.B2_GLM_LOGIT and m.ndim == 4


def test_labels_from_dataframe_series_and_dict():
    X, y = _data()
    df = pd.DataFrame(X, columns=["age", "dose", "bmi"])
    assert pm.glm.GLM(df, pd.Series(y), family=pm.glm.families.Binomial()).free_RVs == ["Intercept", "age", "dose", "bmi"]
    assert pm.glm.GLM(df, y, labels=["a", "b", "c"], family="binomial").free_RVs == ["Intercept", "a", "b", "c"]
    assert pm.glm.GLM(pd.Series(X[:, 0], name="z"), y, family="binomial").free_RVs == ["Intercept", "z"]
    assert pm.glm.GLM({"u": X[:, 0], "v": X[:, 1]}, y, family="binomial").free_RVs == ["Intercept", "u", "v"]
    with pytest.raises(ValueError):
        pm.glm.GLM(X, y, labels=["only_one"], family="binomial")


def test_name_prefixes_the_variables():
    X, y = _data()
    assert pm.glm.GLM(X, y, family="binomial", name="sub").free_RVs == ["sub_Intercept", "sub_x0", "sub_x1", "sub_x2"]


def test_regressor_prior_and_sigma_tau_conversion():
    X, y = _data()
    m = pm.glm.GLM(X, y, family="binomial", priors={"Regressor": pm.glm.Normal.dist(mu=0, sigma=2.0),
                                                     "Intercept": pm.glm.Flat.dist()})
    assert m.prior_tau == pytest.approx(0.25)
    assert pm.glm.Normal.dist(mu=0, tau=4.0).tau == 4.0 and pm.glm.Normal.dist(sd=0.5).tau == pytest.approx(4.0)
    with pytest.raises(ValueError):
        pm.glm.Normal.dist(tau=1.0, sigma=1.0)


def test_what_the_engine_does_not_fuse_is_refused_loudly():
    X, y = _data()
    with pytest.raises(NotImplementedError):
        pm.glm.GLM(X, y)                                   # reference default family is 'normal'
    with pytest.raises(NotImplementedError):
        pm.glm.GLM(X, y, family="poisson")
    with pytest.raises(KeyError):
        pm.glm.GLM(X, y, family="no_such_family")
    with pytest.raises(NotImplementedError):
        pm.glm.GLM(X, y, family="binomial", intercept=False)
    with pytest.raises(NotImplementedError):
        pm.glm.GLM(X, y, family="binomial", offset=1.0)
    with pytest.raises(NotImplementedError):
        pm.glm.GLM(X, y, family="binomial", priors={"x1": pm.glm.Normal.dist(0, 1)})
    with pytest.raises(NotImplementedError):
        pm.glm.GLM(X, y, family="binomial", priors={"Intercept": pm.glm.Normal.dist(0, 1)})
    with pytest.raises(TypeError):
        pm.glm.GLM(X, np.stack([y, y], axis=1), family="binomial")          # glm/linear.py:55-58


@pytest.mark.gpu
def test_glm_frontend_samples_like_logistic_glm():
    X, y = _data(400, 3, seed=3)
    kw = dict(draws=60, tune=60, chains=4, random_seed=5, progressbar=False)
    with pm.glm.GLM(pd.DataFrame(X, columns=list("abc")), y, family="binomial"):
        t1 = pm.sample(**kw)
    with pm.LogisticGLM(X, y, labels=list("abc")):
        t2 = pm.sample(**kw)
    assert t1.varnames == t2.varnames
    assert np.array_equal(t1["b"], t2["b"])
