"""CPU oracle for the Kessler Monte Carlo conjunction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / reference legs of
``bench.py`` may import or execute it, and there only as the checker or as the
timed CPU baseline -- never as the thing shipped.  ``kessler_b200`` never
imports this package.

What it restates (reference = /root/reference/kessler, v1.0.1):

* ``oracle.sgp4_ref``     -- SGP4 ``sgp4init`` / propagate as the un-vendored,
  un-pinned third-party dependency ``dsgp4`` (esa/dSGP4, ``setup.py:35``)
  computes it: Vallado et al. 2006 near-earth branch, WGS-84, opsmode 'i'
  (SURVEY.md Appendix A).  dsgp4's source is absent from /root/reference, so
  parity is anchored on the reference's own golden vector
  ``docs/notebooks/trace.pickle`` (28 init coefficients x 2 TLEs, mean
  elements at the last epoch, i_conj / d_min / time_min of a full forward()
  run) and on ``tests/test_util.py:105-131``.
* ``oracle.kessler_path`` -- Kessler's own code around it: ``util.py:541-580``
  (subsample + linear upsample), ``model.py:23-52`` (find_conjunction),
  ``model.py:445-551`` (Monte Carlo covariance), ``model.py:424-437``
  (all-pairs collision count), ``util.py:304-341`` (RTN rotation).
* ``oracle/csrc/oracle.c`` -- the same algorithms as scalar C, built by
  ``oracle/Makefile`` into ``oracle/_build/liboracle.so`` (used for larger
  parity cases and as a second CPU baseline).

Pinned by: tests/test_oracle_golden.py against tests/golden/*.json|npz
(fixtures made by tests/golden/make_golden.py in the build container, where
the reference's own non-dsgp4 functions were executed through module stubs).
"""
