"""CPU oracle for the Achelous 5-task forward.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or as the CPU baseline being timed.  The product
(``achelous_b200``) never imports this package and fails loudly when its
CUDA library is missing.
"""
