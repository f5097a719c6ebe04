"""`python -m qrec_b200 path/to/model.conf [--seed S]`: run one QRec configuration file on the engine.

The reference's entry point is an interactive menu (main.py:1-70) that ends in
`QRec(ModelConf('./config/<name>.conf')).execute()`; this is that last line with the conf path taken
from the command line.  Paths inside the conf (ratings=, -testSet, output.setup -dir) are resolved
relative to the current directory, as in the reference."""
import argparse
import random
import sys
import time

import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser(prog='python -m qrec_b200', description=__doc__.split('\n')[0])
    ap.add_argument('conf', help='QRec .conf file (model.name must be one of the engine\'s models)')
    ap.add_argument('--seed', type=int, default=None, help='random.seed / np.random.seed before loading the data')
    args = ap.parse_args(argv)
    if args.seed is not None:
        random.seed(args.seed)
        np.random.seed(args.seed)
    from .QRec import QRec
    from .util.config import ModelConf
    t0 = time.time()
    measure = QRec(ModelConf(args.conf)).execute()
    print('Running time: %f s' % (time.time() - t0))
    return measure


if __name__ == '__main__':
    main()
    sys.exit(0)
