"""CPU oracle for the B200 NUTS/HMC engine.  TEST INFRASTRUCTURE ONLY.

This package is a CPU (NumPy/SciPy, fp64) restatement of the reference's
(PyMC3 v3.8) NUTS / HamiltonianMC hot path.  It exists to *check* the CUDA
engine and to serve as the timed CPU baseline in ``bench.py``; it is never the
thing that is shipped or measured as the product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from here.  Nothing under
``pymc3_b200/`` imports it (a test enforces that).

Parity status ("pinned" = checked against the reference's own golden vectors):

* sampler logic (``hmc_cpu.py``, ``potentials.py``)  -- PINNED: reproduces the
  reference's 100-draw NUTS and HamiltonianMC known-answer traces
  (``pymc3/tests/test_step.py:163-266, 371-474``) to < 1e-8 in ``legacy`` RNG mode
  (``tests/test_oracle_golden.py``; vectors in ``tests/golden/``).
* densities (``densities.py``) -- PINNED for Normal / HalfNormal / HalfCauchy /
  Exponential / StudentT / Bernoulli-logit / Binomial against ``scipy.stats`` exactly
  as the reference's ``check_logp`` does (``pymc3/tests/test_distributions.py:456-465``),
  and against the developer-guide vectors
  (``docs/source/developer_guide.rst:151-155, 572-575, 715-737``).
  ``GaussianRandomWalk.logp`` is PARITY UNPINNED in the reference itself (it has no
  value test, only ``.random``); here it is pinned only through ``Normal.logp`` and
  the restatement of ``pymc3/distributions/timeseries.py:237-256``.
* gradients -- Theano's ``tt.grad`` is not available; analytic gradients are checked
  by central finite differences in fp64 and by the developer-guide 20-vector.
* bulk-ESS / R-hat (``diagnostics.py``) -- PARITY UNPINNED (arviz is not vendored by
  the reference and is absent here); restated from Vehtari et al. 2021.
"""
