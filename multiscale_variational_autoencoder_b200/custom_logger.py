"""Package logger (mirrors mvae/custom_logger.py:7-14)."""
import logging

logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)-8s %(message)s")
logger = logging.getLogger("mvae")
