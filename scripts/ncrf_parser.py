"""Drop-in replacement for the reference's scripts/ncrf_parser.py: copy this file (and keep
centroflye_b200 importable) over the original; every public name is re-exported."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroflye_b200.ncrf_parser import *  # noqa: E402,F401,F403
