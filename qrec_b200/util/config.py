"""Configuration objects with the reference's file format and quirks (util/config.py:2-66).

ModelConf  : `key=value` lines; a line that does not split into exactly two fields on '=' is
             reported and skipped; a missing key prints a message and exits(-1) (config.py:7-11).
OptionConf : "on|off -flag value -flag2 v1 v2" strings.  A token is a flag when it starts with '-'
             and what follows the first character is not all digits, so "-1" is a value while
             "-0.5" is (as in the reference) taken for a flag.
"""
import os
import sys


def _die(msg):
    print(msg)
    sys.exit(-1)


class ModelConf(object):
    def __init__(self, fileName=None, entries=None):
        self.config = {}
        if fileName is not None:
            self.readConfiguration(fileName)
        if entries:
            self.config.update(entries)

    @classmethod
    def from_string(cls, text):
        conf = cls()
        conf._parse(text.splitlines())
        return conf

    def _parse(self, lines):
        for lineno, raw in enumerate(lines):
            text = raw.strip()
            if not text:
                continue
            fields = text.split('=')
            if len(fields) != 2:
                print('config file is not in the correct format! Error Line:%d' % lineno)
                continue
            self.config[fields[0]] = fields[1]

    def readConfiguration(self, path):
        if not os.path.exists(path):
            print('config file is not found!')
            raise IOError(path)
        with open(path) as fh:
            self._parse(fh)

    def contains(self, key):
        return key in self.config

    def __getitem__(self, key):
        if key not in self.config:
            _die('parameter ' + key + ' is invalid!')
        return self.config[key]


def _is_flag(token):
    return token.startswith('-') and not token[1:].isdigit()


class OptionConf(object):
    def __init__(self, content):
        self.line = content.strip().split(' ')
        self.mainOption = self.line[0] == 'on'
        self.options = {}
        toks = self.line
        for pos, tok in enumerate(toks):
            if not _is_flag(tok):
                continue
            end = pos + 1
            while end < len(toks) and not _is_flag(toks[end]):
                end += 1
            self.options[tok] = ' '.join(toks[pos + 1:end])

    def __getitem__(self, key):
        if key not in self.options:
            _die('parameter ' + key + ' is invalid!')
        return self.options[key]

    def keys(self):
        return self.options.keys()

    def isMainOn(self):
        return self.mainOption

    def contains(self, key):
        return key in self.options
