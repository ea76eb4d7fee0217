"""CPU oracle for the ViViT low-rank GGN hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  It may be imported by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` -- and nowhere else.  ``vivit_b200`` never imports it.

Parity status: **pinned by relation, not by golden files.**  The reference
(``/root/reference``, f-dangel/vivit v1.0.0) cannot execute in this image
(``backpack-for-pytorch`` is not installed and ``Tensor.symeig`` was removed
from torch 2.x), and it ships no golden vectors for this path.  Its own tests
pin the path against a brute-force ``torch.autograd`` GGN
(``test/implementation/autograd.py``); ``oracle/autograd_ggn.py`` restates that
ground truth and ``tests/test_oracle_*.py`` check ``oracle/reference_path.py``
against it with the reference's tolerances on the reference's test problems
(``test/settings.py``).  The one binary fixture the reference holds
(``test/utils/tensor_causes_symeig_error.pt``) is re-exported to
``tests/golden/`` by ``tests/golden/make_golden.py``.
"""
