from .correlation import corr, patchify, CorrLayer, PatchLayer, corr_pyramid2, PyramidRing  # noqa: F401
