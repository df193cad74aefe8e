"""`pm.glm.GLM` front-end for the fused Bernoulli-logit family (reference: pymc3/glm/linear.py:29-160,
glm/families.py:82-119, glm/utils.py:20-120).

The reference builds one scalar RV per design-matrix column on a Theano graph; here the same call
describes the model to the engine (family id + data pointers).  Only what the engine fuses is accepted --
`family` binomial/logit, Flat intercept, one Normal(0, tau) prior shared by the regressors -- anything
else raises NotImplementedError naming the missing piece, never a silent approximation.
"""
import numpy as np

from .model import LogisticGLM

__all__ = ["GLM", "families", "Normal", "Flat"]


class _Prior:
    """stand-in for `Distribution.dist(...)` objects passed in `priors=` (glm/linear.py:47-48)"""

    @classmethod
    def dist(cls, *args, **kwargs):
        return cls(*args, **kwargs)


class Normal(_Prior):
    def __init__(self, mu=0.0, sigma=None, tau=None, sd=None):
        if sd is not None:                         # continuous.py:472-476: `sd` is an alias of `sigma`
            sigma = sd
        if tau is None:
            tau = 1.0 if sigma is None else float(sigma) ** -2.0      # get_tau_sigma, continuous.py:105-144
        elif sigma is not None:
            raise ValueError("Can't pass both tau and sd")            # continuous.py:123-124
        self.mu, self.tau = float(mu), float(tau)


class Flat(_Prior):
    pass


class _Family:
    link = None


class Binomial(_Family):
    """glm/families.py:115-119: Binomial(n=1) likelihood, logit link"""
    link = "logit"


class _Families:
    Binomial = Binomial

    class Normal(_Family):
        link = "identity"

    class StudentT(_Family):
        link = "identity"

    class Poisson(_Family):
        link = "log"

    class NegativeBinomial(_Family):
        link = "log"


families = _Families()
_BY_NAME = dict(normal=families.Normal, student=families.StudentT, binomial=families.Binomial,
                poisson=families.Poisson, negative_binomial=families.NegativeBinomial)      # glm/linear.py:139-145


def any_to_array_and_labels(x, labels=None):
    """glm/utils.py:20-120 for the inputs that make sense without Theano: DataFrame / Series / dict /
    ndarray / torch tensor -> (2-d array, column labels); default labels x0, x1, ..."""
    if isinstance(labels, str):
        labels = [labels]
    mod = type(x).__module__
    if mod.startswith("pandas"):
        if hasattr(x, "columns"):                   # DataFrame
            labels = list(labels) if labels else [str(c) for c in x.columns]
            x = x.to_numpy()
        else:                                       # Series
            labels = list(labels) if labels else [str(x.name)]
            x = x.to_numpy()[:, None]
    elif isinstance(x, dict):
        labels = [str(k) for k in x.keys()]          # dict keys are the labels whatever `labels` says (utils.py:70-85)
        x = np.stack([np.asarray(v) for v in x.values()], axis=1)
    elif mod.startswith("torch"):
        if x.dim() == 1:
            x = x[:, None]
    else:
        x = np.asarray(x)
        if x.ndim == 1:
            x = x[:, None]
    if x.ndim != 2:
        raise ValueError("x must be one- or two-dimensional")
    if labels is None:
        labels = ["x%d" % i for i in range(x.shape[1])]
    labels = list(labels)
    if len(labels) != x.shape[1]:
        raise ValueError("Please provide full list of labels for coefficients, got len(labels) != x.shape[1]")  # utils.py:112-117
    return x, labels


class GLM(LogisticGLM):
    """pm.glm.GLM(x, y, intercept=True, labels=None, priors=None, vars=None, family='normal', name='',
    model=None, offset=0.) -- glm/linear.py:130-160.  `family` must be binomial here."""

    def __init__(self, x, y, intercept=True, labels=None, priors=None, vars=None, family="normal", name="",
                 model=None, offset=0.0):
        if isinstance(family, str):
            if family not in _BY_NAME:
                raise KeyError(family)
            family = _BY_NAME[family]()
        elif isinstance(family, type):
            family = family()
        if not isinstance(family, Binomial):
            raise NotImplementedError("the engine fuses the binomial (logit) family only; got %s"
                                      % type(family).__name__)
        if vars:
            raise NotImplementedError("`vars=` (user-supplied random variables) needs the general model front-end")
        if not intercept:
            raise NotImplementedError("the fused GLM kernels always carry a Flat intercept (intercept=True)")
        if np.any(np.asarray(offset) != 0):
            raise NotImplementedError("`offset` is not supported by the fused GLM kernels")
        priors = dict(priors or {})
        icpt = priors.pop("Intercept", None)
        if icpt is not None and not isinstance(icpt, Flat):
            raise NotImplementedError("Intercept prior must be Flat (the reference's default, glm/linear.py:50)")
        reg = priors.pop("Regressor", None)
        if priors:
            raise NotImplementedError("per-coefficient priors %r need the general model front-end" % sorted(priors))
        tau = 1e-6                                                   # glm/linear.py:49
        if reg is not None:
            if not isinstance(reg, Normal) or reg.mu != 0.0:
                raise NotImplementedError("Regressor prior must be Normal(mu=0, tau=...)")
            tau = reg.tau
        if hasattr(y, "to_numpy"):
            y = y.to_numpy()
        x, labels = any_to_array_and_labels(x, labels)
        prefix = (name + "_") if name else ""                         # model.py:770-776 name mangling of sub-models
        super().__init__(x, y, labels=[prefix + l for l in labels], prior_tau=tau)
        if prefix:
            self.free = tuple((prefix + n if n == "Intercept" else n, s) for n, s in self.free)
        self.family_obj = family

    @classmethod
    def from_formula(cls, formula, data, priors=None, vars=None, family="normal", name="", model=None, offset=0.0,
                     eval_env=0):
        """glm/linear.py:101-127: needs `patsy` exactly like the reference"""
        import patsy
        eval_env = patsy.EvalEnvironment.capture(eval_env, reference=1)
        y, x = patsy.dmatrices(formula, data, eval_env=eval_env)
        labels = list(x.design_info.column_names)
        x = np.asarray(x)
        if labels and labels[0] == "Intercept":                      # patsy's own intercept column is ours
            x, labels = x[:, 1:], labels[1:]
        return cls(x, np.asarray(y)[:, -1], intercept=True, labels=labels, priors=priors, vars=vars, family=family,
                   name=name, model=model, offset=offset)

