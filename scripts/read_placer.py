"""Drop-in replacement for the reference's scripts/read_placer.py: copy this file (and keep centroflye_b200
importable) over the original; same ReadPlacer class, same command line, same read_positions.csv."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroflye_b200.read_placer import *  # noqa: E402,F401,F403
from centroflye_b200.read_placer import main  # noqa: E402

if __name__ == "__main__":
    main()
