"""Stand-in for the reference's `pycnn` package: the SAME import path, served by pycnn_b200 with the B200 backend on
(see pycnn.py next to this file).  Only the test tier puts this directory on sys.path."""
