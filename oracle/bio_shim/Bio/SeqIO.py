"""``Bio.SeqIO.parse(filename, format)`` for fasta files only (all the hot path's imports need)."""
from collections import namedtuple

_Rec = namedtuple("_Rec", ["id", "seq"])


def parse(filename, format="fasta"):
    if format != "fasta":
        raise NotImplementedError("bio_shim only reads fasta")
    name, chunks = None, []
    with open(filename) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if name is not None:
                    yield _Rec(name, "".join(chunks))
                name, chunks = line[1:].split()[0] if len(line) > 1 else "", []
            elif name is not None:
                chunks.append(line.strip())
    if name is not None:
        yield _Rec(name, "".join(chunks))
