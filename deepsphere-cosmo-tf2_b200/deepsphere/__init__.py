"""deepsphere — B200-native drop-in for the graph-convolution hot path of
deepsphere/deepsphere-cosmo-tf2 (same module and class names as the reference package,
reference __init__.py:1-9)."""
from deepsphere._logger import logger
from deepsphere.healpy_networks import HealpyGCNN

__version__ = "0.3.0"

__all__ = ["HealpyGCNN", "__version__", "logger"]
