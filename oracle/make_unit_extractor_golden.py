#!/usr/bin/env python
"""Outputs of the reference's unit_extractor.py k-mer analysis on reads of the golden reports (TEST INFRASTRUCTURE).

    python oracle/make_unit_extractor_golden.py

scripts/unit_extractor.py imports matplotlib (absent here) at module level, so the five pure functions
get_repetitive_kmers, get_convolution, get_period_info, get_hook_kmer, split_by_hook are compiled out of the reference
file by name (ast) and run on the gap-free rows of the first three records of every golden report at k = 15 (the
script's default) and 19, bin size 10.  Stored per case: tests/golden/<case>/unit_extractor.json.
"""
import ast
import gzip
import hashlib
import json
import os
import sys
import tempfile
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "bio_shim"))
sys.path.insert(0, os.path.join(REF, "scripts"))
NAMES = ["get_repetitive_kmers", "get_convolution", "get_period_info", "get_hook_kmer", "split_by_hook"]


def reference_functions():
    path = os.path.join(REF, "scripts", "unit_extractor.py")
    tree = ast.parse(open(path).read())
    body = ast.parse("from collections import defaultdict\nfrom bisect import bisect_left, bisect_right\n").body
    body += [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in NAMES]
    ns = {}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns


def digest(obj):
    return hashlib.md5(json.dumps(obj, sort_keys=False, separators=(",", ":")).encode()).hexdigest()


def main():
    ref = reference_functions()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import ncrf_parser
    golden = os.path.join(ROOT, "tests", "golden")
    for case in sorted(os.listdir(golden)):
        d = os.path.join(golden, case)
        if not os.path.exists(os.path.join(d, "report.ncrf.gz")):
            continue
        with tempfile.TemporaryDirectory() as tmp:
            rep = os.path.join(tmp, "report.ncrf")
            with gzip.open(os.path.join(d, "report.ncrf.gz"), "rb") as f, open(rep, "wb") as g:
                g.write(f.read())
            report = ncrf_parser.NCRF_Report(rep)
        out = []
        for r_id, rec in list(report.records.items())[:3]:
            seq = rec.r_al.replace("-", "").upper()
            for k in (15, 19):
                rep_kmers = ref["get_repetitive_kmers"](seq, k)
                conv, union_conv = ref["get_convolution"](rep_kmers)
                periods, bin_convs, bin_left, bin_right = ref["get_period_info"](union_conv, 10)
                hook = ref["get_hook_kmer"](conv, bin_left, bin_right) if union_conv else None
                splits = ref["split_by_hook"](seq, hook) if hook else {}
                out.append(dict(r_id=r_id, k=k, seq_len=len(seq), n_rep_kmers=len(rep_kmers),
                                rep_kmers_md5=digest(list(rep_kmers.items())), conv_md5=digest(list(conv.items())),
                                n_union_conv=len(union_conv), union_conv_md5=digest(union_conv),
                                periods=list(periods)[:20], bin_convs=list(bin_convs)[:20], bin_left=bin_left,
                                bin_right=bin_right, hook=hook, split_ids=list(splits.keys()),
                                splits_md5=digest(list(splits.items()))))
                print(case, r_id, k, len(seq), len(rep_kmers), len(union_conv), list(periods)[:2], hook, len(splits), flush=True)
        with open(os.path.join(d, "unit_extractor.json"), "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
