"""``utils.Logger`` of the reference (baseline/utils/Logger.py): the module-level ``LOG`` the reference's scripts import.
The reference opens ``Baseline.log`` in the current directory at import; here the file handler is added only when
``DCASE_LOG_FILE`` names a path, so importing the package has no side effect on the working directory."""
import logging
import os
import sys


def create_logger(logger_name, log_file=None):
    logger = logging.getLogger(logger_name)
    logger.setLevel(logging.DEBUG)
    if not any(h.get_name() == "stdout" for h in logger.handlers):
        terminal = logging.StreamHandler(sys.stdout)
        terminal.setLevel(logging.INFO)
        terminal.set_name("stdout")
        terminal.setFormatter(logging.Formatter(" %(levelname)s - %(message)s"))
        logger.addHandler(terminal)
    if log_file and not any(h.get_name() == "file_handler" for h in logger.handlers):
        to_file = logging.FileHandler(log_file)
        to_file.setLevel(logging.DEBUG)
        to_file.set_name("file_handler")
        to_file.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
        logger.addHandler(to_file)
    return logger


LOG = create_logger("baseline", os.environ.get("DCASE_LOG_FILE"))
