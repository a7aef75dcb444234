"""CPU oracle for the bixverse co-expression correlation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``bixverse_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and
there only as the checker / the timed CPU baseline.
"""
