"""Package logger; mirrors the reference's env-var contract (_logger.py:6-38): logger name
"sbi_flows", level from DEEPSPHERE_LOG_LEVEL (1 critical .. 5 debug, default info)."""
import logging
import os
import sys

logger = logging.getLogger("sbi_flows")
if not logger.handlers:
    _h = logging.StreamHandler(sys.stdout)
    _h.setFormatter(logging.Formatter("%(asctime)s %(name)10s %(levelname).3s   %(message)s ", "%y-%m-%d %H:%M:%S"))
    logger.addHandler(_h)
logger.propagate = False
logger.setLevel(logging.INFO)

_LEVELS = {1: logging.CRITICAL, 2: logging.ERROR, 3: logging.WARNING, 4: logging.INFO}
if "DEEPSPHERE_LOG_LEVEL" in os.environ:
    try:
        _lvl = int(os.environ["DEEPSPHERE_LOG_LEVEL"])
    except ValueError:
        logger.warning("Loglevel set in DEEPSPHERE_LOG_LEVEL is not an int, got %s. Using default INFO!",
                       os.environ["DEEPSPHERE_LOG_LEVEL"])
        _lvl = 4
    logger.setLevel(_LEVELS.get(max(_lvl, 1), logging.DEBUG))
