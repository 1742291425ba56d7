"""CPU oracle for the reward-evaluation hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in Python 3 + numpy/scipy float64, the arithmetic of the
reference's hot path (``gpmodel_library.py`` / ``aq_library.py`` / the selection
rules of ``mcts_library.py``) so that the CUDA path can be checked against it.

Rules (enforced by tests/test_no_oracle_in_product.py):

* Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
  ``--impl reference`` legs may import anything from here.
* Nothing under ``informative_path_planning_b200/`` imports it; the product
  path has no CPU fallback and raises if the CUDA library is missing.

Parity status: PINNED TO THE REFERENCE'S OWN NUMPY LINES, GPy RESTATED.
The reference ships no tests or golden vectors and is
Python 2.7 + GPy, neither of which exists in this image.  ``tests/golden/
make_golden.py`` therefore executes the reference's *own source files* from
``/root/reference`` under Python 3 (in-memory print/xrange transform, nothing
copied) with the absent third-party ``GPy`` replaced by ``oracle.gpy_restated``
(a restatement of GPy 1.9.x's published RBF / pdinv / jitchol / GPRegression
algorithms; GPy is un-vendored and unpinned by the reference).  The vectors it
writes pin this oracle to the reference's numpy arithmetic; the GPy layer
itself remains "restated, unpinned" and is cross-checked by independent
formulations (Cholesky-form vs inverse-form, longdouble spot checks, analytic
cases) in tests/test_oracle.py.
"""
