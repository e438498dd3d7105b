"""Drop-in for the reference's scripts/read_recruitment/rr binary: `python rr.py unit.fasta reads.fasta[.gz] output.fasta
edit_distance_threshold` (run_read_recruitment.sh calls `$SCRIPT_DIR/rr` with exactly these four arguments)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroflye_b200.read_recruitment import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
