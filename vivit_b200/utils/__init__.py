"""Small helpers shared by the computations."""


def delete_savefield(param, savefield: str, verbose: bool = False) -> None:
    """Drop the per-parameter buffer stored under ``savefield`` once a hook has consumed it
    (``vivit/utils/__init__.py:8-19``; raises ``AttributeError`` if it is not there)."""
    delattr(param, savefield)
    if verbose:
        print(f"Param {id(param)}: Delete '{savefield}'")


def keep_indices(keep, evals):
    """Positions selected by a group's ``criterion`` as an ``int64`` index tensor on ``evals``' device.

    The reference indexes with the criterion's result directly (``evals[keep]``, ``evecs[:, keep]``:
    ``vivit/linalg/eigh.py:252-265``, ``vivit/optim/directional_derivatives.py:293-300``), so everything
    that indexing accepts is accepted here with the same meaning: lists / tensors of positions, negative
    positions (counted from the end) and boolean masks of length ``len(evals)``.
    """
    import torch

    n = evals.numel()
    idx = torch.as_tensor(keep)  # lists stay on the host: the range checks below cost no device sync
    if idx.dtype == torch.bool:
        idx = idx.to(evals.device)
        if idx.dim() != 1 or idx.numel() != n:
            raise IndexError(f"boolean criterion mask of shape {tuple(idx.shape)} does not match {n} eigenvalues")
        return idx.nonzero(as_tuple=False).flatten()
    if idx.numel() == 0:
        return idx.to(device=evals.device, dtype=torch.int64).flatten()
    if idx.is_floating_point() or idx.is_complex():
        raise IndexError("criterion must return integer positions or a boolean mask")
    idx = idx.to(torch.int64).flatten()
    if bool(((idx < -n) | (idx >= n)).any()):
        raise IndexError(f"criterion returned a position outside [-{n}, {n})")
    return torch.where(idx < 0, idx + n, idx).to(evals.device)
