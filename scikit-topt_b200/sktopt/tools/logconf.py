"""Logging helper (reference ``tools/logconf.py:3-11``)."""
import logging

logging.basicConfig(level=logging.WARNING,
                    format="%(asctime)s %(name)s %(levelname)s %(message)s")


def mylogger(name: str) -> logging.Logger:
    return logging.getLogger(name)
