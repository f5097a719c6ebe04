"""Text loaders / writers with the reference's column and binarisation options (util/io.py:19-85)."""
import os
import re
import sys

from .config import OptionConf


class FileIO(object):
    @staticmethod
    def writeFile(dir, file, content, op='w'):
        os.makedirs(dir, exist_ok=True)
        with open(dir + file, op) as fh:
            fh.writelines(content)

    @staticmethod
    def deleteFile(filePath):
        if os.path.exists(filePath):
            os.remove(filePath)

    @staticmethod
    def loadDataSet(conf, file, bTest=False, binarized=False, threshold=3.0):
        """-> [[user, item, float(rating)], ...].  `-columns a b [c]`, `-header`, `-delim`; with
        binarisation rows whose rating is below `threshold` are dropped and the rest become 1."""
        setup = OptionConf(conf['ratings.setup'])
        print('loading test data...' if bTest else 'loading training data...')
        with open(file) as fh:
            lines = fh.readlines()
        if setup.contains('-header'):
            lines = lines[1:]
        cols = [int(c) for c in setup['-columns'].strip().split()]
        splitter = re.compile(setup['-delim'] if setup.contains('-delim') else ' |,|\t')
        records = []
        for lineno, line in enumerate(lines):
            fields = splitter.split(line.strip())
            if not bTest and len(cols) < 2:
                print('The rating file is not in a correct format. Error: Line num %d' % lineno)
                sys.exit(-1)
            try:
                user, item = fields[cols[0]], fields[cols[1]]
                rating = fields[cols[2]] if len(cols) >= 3 else 1
                if binarized:
                    if float(fields[cols[2]]) < threshold:
                        continue
                    rating = 1
                records.append([user, item, float(rating)])
            except ValueError:
                print('Error! Have you added the option -header to the rating.setup?')
                sys.exit(-1)
        return records

    @staticmethod
    def loadUserList(filepath):
        print('loading user List...')
        with open(filepath) as fh:
            return [line.strip().split()[0] for line in fh]

    @staticmethod
    def loadRelationship(conf, filePath):
        """-> [[user1, user2, weight], ...] (util/io.py:88-111 of the reference): `social.setup` gives `-columns a b [c]`
        and `-header`; the weight defaults to 1."""
        setup = OptionConf(conf['social.setup'])
        print('loading social data...')
        with open(filePath) as fh:
            lines = fh.readlines()
        if setup.contains('-header'):
            lines = lines[1:]
        cols = [int(c) for c in setup['-columns'].strip().split()]
        splitter = re.compile(' |,|\t')
        out = []
        for lineno, line in enumerate(lines):
            fields = splitter.split(line.strip())
            if len(cols) < 2:
                print('The social file is not in a correct format. Error: Line num %d' % lineno)
                sys.exit(-1)
            out.append([fields[cols[0]], fields[cols[1]], float(fields[cols[2]]) if len(cols) >= 3 else 1])
        return out
