"""CPU oracle for the holocron_b200 hot path — TEST INFRASTRUCTURE, not product code.

Every function here is a plain-PyTorch (CPU, fp32) restatement of the reference algorithm
(frgfm/Holocron, file:line cited per function). It exists only so that ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs can check or time the
CUDA path against it. Nothing under ``holocron_b200/`` imports this package, and the product path raises when
the CUDA library is missing instead of falling back to it.

Pinning: the oracle is checked against (a) the exact known-answer vectors of the reference's own
``tests/test_ops.py`` and (b) golden input/output fixtures under ``tests/golden/`` that were generated in the
build container by importing the unmodified reference from ``/root/reference`` (script:
``tests/golden/make_golden.py``). ``tests/test_oracle_golden.py`` runs both on CPU.
"""
from . import boxes, functional, models, optim  # noqa: F401
