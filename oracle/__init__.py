"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Fireflies hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline -- never as the thing shipped.  ``fireflies_b200`` never
imports this package.
"""
