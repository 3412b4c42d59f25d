import sys; sys.path.insert(0,'.')
from pycnn_b200 import _lib as lib
c=lib.Context(0); print(c.info()); c.close()
