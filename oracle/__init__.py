"""CPU oracle for the U-Net + Dice hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker or as the
timed CPU baseline -- never on the path that is measured or shipped.
"""
