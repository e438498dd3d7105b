"""Drop-in replacement for the reference's scripts/distance_based_kmer_recruitment.py: copy this file (and keep
centroflye_b200 importable) over the original; every public name is re-exported."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroflye_b200.distance_based_kmer_recruitment import *  # noqa: E402,F401,F403
from centroflye_b200.distance_based_kmer_recruitment import main, parse_args  # noqa: E402,F401

if __name__ == "__main__":
    main()
