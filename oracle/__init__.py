"""CPU oracle for the apyib correlated-wavefunction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``apyib_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / CPU
baseline, never as the product path.
"""
