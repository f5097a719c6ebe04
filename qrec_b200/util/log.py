"""File logger, one file per run under ./log/ (reference: util/log.py:4-15)."""
import logging
import os


class Log(object):
    def __init__(self, module, filename, directory='./log/'):
        os.makedirs(directory, exist_ok=True)
        self.logger = logging.getLogger(module)
        self.logger.setLevel(logging.INFO)
        fh = logging.FileHandler(os.path.join(directory, filename + '.log'))
        fh.setFormatter(logging.Formatter('%(asctime)s - %(name)s - %(levelname)s - %(message)s'))
        self.logger.addHandler(fh)

    def add(self, text):
        self.logger.info(text)
