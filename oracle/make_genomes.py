#!/usr/bin/env python
"""Genomes of the BASELINE.json configs, made by the UNMODIFIED reference simulator (TEST / BENCH INFRASTRUCTURE).

Runs only in the build container (where /root/reference exists):

    python oracle/make_genomes.py

For every entry of GENOMES it runs the reference's own command line

    scripts/simulate_tandem_repeat.py --unit supplementary_data/<unit>.fasta --multiplicity M --div-rate 0.01 --seed S -o <tmp>

(/root/reference/scripts/simulate_tandem_repeat.py:58-89, legacy np.random, bit-stable for a seed) with
oracle/bio_shim standing in for Biopython, reads flanked_tandem_repeat.fasta back and stores it 2-bit packed under
tests/golden/genomes/<name>.npz together with the unit, the array start / length and the md5 of the fasta the
reference wrote.  The fixtures travel to the GPU box; the reference does not.  bench.py and the full-size parity
tests draw their reads from these genomes with centroflye_b200.synth.simulate_reads.
"""
import hashlib
import os
import runpy
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "genomes")

GENOMES = {
    # configs[0] / configs[1] / configs[3]: the cenX-like array
    "cenx_dxz1_m1500_s1": dict(unit="supplementary_data/DXZ1_rc.fasta", multiplicity=1500, div_rate=0.01, seed=1),
    # weak scaling of configs[1] over N GPUs (bench.py --gpus N): N cenX-like arrays.  Array j >= 2 has its OWN unit --
    # DXZ1 with 30 % of its bases substituted (numpy default_rng(100 + j)) -- so that the arrays share no k-mers and the
    # work per GPU stays that of configs[1] (arrays of one unit share their error k-mers: at 8 x 50x they enter the
    # rare band and the pair increments per array nearly double); the array itself is made by the reference simulator
    **{f"cenx_like_u{s}_m1500_s{s}": dict(unit="supplementary_data/DXZ1_rc.fasta", unit_subst=(0.30, 100 + s),
                                         multiplicity=1500, div_rate=0.01, seed=s) for s in range(2, 9)},
    # configs[2]: the cen6-like array
    "cen6_d6z1_m1000_s4": dict(unit="supplementary_data/D6Z1.fasta", multiplicity=1000, div_rate=0.01, seed=4),
}


def read_fasta(path):
    with open(path) as f:
        return "".join(line.strip() for line in f if not line.startswith(">"))


def main():
    sys.path.insert(0, ROOT)
    from centroflye_b200.encode import ascii_to_codes
    os.makedirs(OUT, exist_ok=True)
    for name, g in GENOMES.items():
        if os.path.exists(os.path.join(OUT, name + ".npz")) and "--force" not in sys.argv:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            unit_path = os.path.join(REF, g["unit"])
            if "unit_subst" in g:
                from centroflye_b200.encode import codes_to_ascii
                rate, useed = g["unit_subst"]
                rng = np.random.default_rng(useed)
                u = ascii_to_codes(read_fasta(unit_path).upper())
                hit = rng.random(u.size) < rate
                u[hit] = (u[hit] + rng.integers(1, 4, size=int(hit.sum()), dtype=np.uint8)) & 3
                unit_path = os.path.join(tmp, "unit.fasta")
                with open(unit_path, "w") as f:
                    f.write(">unit\n" + codes_to_ascii(u) + "\n")
            argv = ["simulate_tandem_repeat.py", "--unit", unit_path, "--multiplicity",
                    str(g["multiplicity"]), "--div-rate", str(g["div_rate"]), "--seed", str(g["seed"]), "-o", tmp]
            old_argv, old_path = sys.argv, list(sys.path)
            sys.argv = argv
            sys.path[:0] = [os.path.join(HERE, "bio_shim"), os.path.join(REF, "scripts")]
            try:
                runpy.run_path(os.path.join(REF, "scripts", "simulate_tandem_repeat.py"), run_name="__main__")
            finally:
                sys.argv, sys.path[:] = old_argv, old_path
            fasta = os.path.join(tmp, "flanked_tandem_repeat.fasta")
            md5 = hashlib.md5(open(fasta, "rb").read()).hexdigest()
            flanked = read_fasta(fasta)
            tr = read_fasta(os.path.join(tmp, "tandem_repeat.fasta"))
            unit = read_fasta(unit_path).upper()
        start = (len(flanked) - len(tr)) // 2
        assert flanked[start:start + len(tr)] == tr and len(tr) == len(unit) * g["multiplicity"]
        codes = ascii_to_codes(flanked)
        pad = (-codes.size) % 4
        c = np.concatenate([codes, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
        packed = (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), packed=packed, n=np.int64(codes.size), unit=np.array(unit),
                            array_start=np.int64(start), array_len=np.int64(len(tr)), fasta_md5=np.array(md5),
                            command=np.array(" ".join(["python"] + argv[:-2])))
        print(f"{name}: {codes.size} bp (array {len(tr)} bp at {start}), fasta md5 {md5}, "
              f"{os.path.getsize(os.path.join(OUT, name + '.npz'))} bytes")


if __name__ == "__main__":
    main()
