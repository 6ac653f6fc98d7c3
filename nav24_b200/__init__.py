"""nav24_b200 — B200-native ORB front end (detector + matchers) behind nav24's FtDt / FtAssoc operator
interface.  The product is nav24_b200/libnav24orb.so (C ABI in include/nav24_orb.h, CUDA sources in
nav24_b200/csrc); this package only holds the ctypes binding used by tests and bench.py and a numpy
frame generator.  No CPU fallback exists."""
__version__ = "0.1.0"
