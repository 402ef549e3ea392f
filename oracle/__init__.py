"""CPU oracle for the genome_resquiggle hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is imported by the product package ``nanoraw_b200``.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or as the timed CPU baseline.
"""
