#!/usr/bin/env python
"""Wall clock of the drop-in command line on configs[1] (what bench.py reports as e2e_cli), without the rest of the bench:
    python tools/cli_time.py [calls]"""
import contextlib
import io
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from centroflye_b200 import distance_based_kmer_recruitment as dbkr  # noqa: E402
from centroflye_b200 import synth  # noqa: E402


def main():
    calls = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    P = bench.CONFIGS["cenx"]["params"]
    unit, batch, units, reads_list = bench.simulate("cenx", 1.0, 0, 1, keep_reads=True)
    with tempfile.TemporaryDirectory() as tmp:
        fn = os.path.join(tmp, "report.ncrf")
        synth.write_ncrf_report(fn, reads_list, unit)
        argv = ["--ncrf", fn, "--coverage", str(P["coverage"]), "--min-coverage", str(P["min_coverage"]), "--outdir",
                os.path.join(tmp, "out"), "-k", str(P["k"]), "--max-distance", str(P["max_d"]), "--bottom", str(P["bottom"]),
                "--top", str(P["top"]), "--kmer-survival-rate", str(P["kmer_survival_rate"]), "--max-nonuniq", str(P["max_nonuniq"])]
        for _ in range(calls):
            t = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                dbkr.main(argv)
            print(json.dumps({"seconds": round(time.perf_counter() - t, 4),
                              **{k: round(v, 4) for k, v in dbkr.LAST_TIMINGS.items()}}), flush=True)


if __name__ == "__main__":
    main()
