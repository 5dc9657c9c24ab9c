"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the thu-ml/stochastic_gcn hot path: a plain-C restatement (``sgcn_oracle.c``), the
unmodified reference C++ compiled behind a C shim (``_ref/libsgcn_ref.so``, built by
``oracle/Makefile`` when ``/root/reference`` is present) and NumPy restatements of the
TensorFlow-side aggregators (``aggregators.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``stochastic_gcn_b200`` never does.
"""
