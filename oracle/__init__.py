"""CPU oracle for the ViViT low-rank GGN hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  It may be imported by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` -- and nowhere else.  ``vivit_b200`` never imports it.

Parity status: **pinned against outputs of the reference itself, run in the build container,**
and by relation to the autograd GGN.

* The reference (``/root/reference``, f-dangel/vivit v1.0.0) cannot be imported as it stands
  (``backpack-for-pytorch`` is not installed and ``Tensor.symeig`` was removed from torch 2.x), but what
  BackPACK contributes to this path are per-parameter tensors; everything after them is the
  reference's own code.  ``tests/golden/make_reference_run.py`` computes those tensors by autograd,
  stubs the BackPACK import names, shims ``symeig`` to ``linalg.eigh`` and runs the reference's four
  unmodified Computations on every fixture of ``tests/problems.py``; the results are committed as
  ``tests/golden/reference_run.pt`` and ``tests/test_reference_run_cpu.py`` holds
  ``oracle/reference_path.py`` to them (agreement 1e-13 in float64).
* The reference's own tests pin the path against a brute-force ``torch.autograd`` GGN
  (``test/implementation/autograd.py``); ``oracle/autograd_ggn.py`` restates that ground truth and
  ``tests/test_oracle_*.py`` check ``oracle/reference_path.py`` against it with the reference's
  tolerances on the reference's test problems (``test/settings.py``).  This also covers the part the
  reference run cannot (BackPACK's factor back-propagation and ``param_mjp``, restated ``[external]``).
* The one binary fixture the reference holds (``test/utils/tensor_causes_symeig_error.pt``) is
  re-exported to ``tests/golden/`` by ``tests/golden/make_golden.py``.
"""
