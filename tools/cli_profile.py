#!/usr/bin/env python
"""cProfile of the second (warm) call of the drop-in command line on configs[1]: where the host time goes.
    python tools/cli_profile.py [top_n]"""
import contextlib
import cProfile
import io
import os
import pstats
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from centroflye_b200 import distance_based_kmer_recruitment as dbkr  # noqa: E402
from centroflye_b200 import synth  # noqa: E402


def main():
    top = int(sys.argv[1]) if len(sys.argv) > 1 else 35
    P = bench.CONFIGS["cenx"]["params"]
    unit, batch, units, reads_list = bench.simulate("cenx", 1.0, 0, 1, keep_reads=True)
    with tempfile.TemporaryDirectory() as tmp:
        fn = os.path.join(tmp, "report.ncrf")
        synth.write_ncrf_report(fn, reads_list, unit)
        argv = ["--ncrf", fn, "--coverage", str(P["coverage"]), "--min-coverage", str(P["min_coverage"]), "--outdir",
                os.path.join(tmp, "out"), "-k", str(P["k"]), "--max-distance", str(P["max_d"]), "--bottom", str(P["bottom"]),
                "--top", str(P["top"]), "--kmer-survival-rate", str(P["kmer_survival_rate"]), "--max-nonuniq", str(P["max_nonuniq"])]
        with contextlib.redirect_stdout(io.StringIO()):
            dbkr.main(argv)
            dbkr.main(argv)
        prof = cProfile.Profile()
        with contextlib.redirect_stdout(io.StringIO()):
            prof.enable()
            dbkr.main(argv)
            prof.disable()
        print({k: round(v, 4) for k, v in dbkr.LAST_TIMINGS.items()})
        out = io.StringIO()
        pstats.Stats(prof, stream=out).sort_stats("cumulative").print_stats(top)
        print(out.getvalue())


if __name__ == "__main__":
    main()
