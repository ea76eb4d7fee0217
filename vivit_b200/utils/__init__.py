"""Small helpers shared by the computations."""


def delete_savefield(param, savefield: str, verbose: bool = False) -> None:
    """Drop the per-parameter buffer stored under ``savefield`` once a hook has consumed it
    (``vivit/utils/__init__.py:8-19``; raises ``AttributeError`` if it is not there)."""
    delattr(param, savefield)
    if verbose:
        print(f"Param {id(param)}: Delete '{savefield}'")
