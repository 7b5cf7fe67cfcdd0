"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference GPTQ QuantLinear math.

Nothing under ``oracle/`` is part of the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker / the timed CPU baseline.
"""
from .awq_oracle import *  # noqa: F401,F403
from .gptq_oracle import *  # noqa: F401,F403
