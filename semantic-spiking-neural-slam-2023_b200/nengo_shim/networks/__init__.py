from .ensemblearray import EnsembleArray  # noqa: F401
