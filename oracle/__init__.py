"""CPU oracle for the OTPose temporal-fusion-head hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``otpose_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or the
timed CPU baseline -- never as the product path.

Parity pin: the reference ships no tests / golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against the reference's own
modules executed in the build container (``oracle/make_golden.py`` imports
``model/ConvVideoTransformer.py``, ``model/blocks.py`` and ``model/RSB.py``
from ``/root/reference`` unmodified, plus ``torchvision.ops.deform_conv2d``),
and the resulting vectors are committed under ``tests/golden/``.
"""
