from .correlation import corr, patchify, CorrLayer, PatchLayer, corr_pyramid2  # noqa: F401
