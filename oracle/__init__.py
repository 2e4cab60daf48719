"""
CPU oracle for the quantum-css-codes hot path.  TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the reference algorithms (jimpo/quantum-css-codes,
``bin_matrix.py`` and the numeric part of ``css_code.py``).  It exists to *check* the CUDA
path; it is never the thing shipped or measured.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package (``quantum_css_codes_b200``) never imports ``oracle``.

Pinning status
--------------
* ``gf2.rref_literal``, ``vec_to_int``, ``int_to_vec``, ``weight_w_vectors``,
  ``css.normalize_parity_check``, ``css.syndrome_table``, ``css.build_css`` are pinned against
  (i) every known-answer test the reference holds for the path (test/test_bin_matrix.py:8-31,
  test/test_css_code.py:13-53,108-143) and (ii) outputs of the unmodified reference imported
  in the build container (``oracle/gen_golden.py`` -> ``tests/golden/*.npz``).
* The Monte-Carlo composition (``montecarlo.py``), the depolarising sampler (``philox.py``),
  and ``gf2.rank / null_space / solve`` have NO counterpart in the reference:
  **parity unpinned** by the reference for those; they are pinned by exact enumeration
  (SURVEY Appendix A.4 failure-weight enumerators) and by algebraic properties instead.
"""
