"""CPU oracle for the categorical-memory hot path (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is part of the product. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and there only as the checker or the reported CPU baseline.
The product path (``pinthememory_b200``) never routes through it.
"""
