"""Sequence helpers with the reference's names (scripts/utils/bio.py), Biopython-free."""
from ..ncrf_parser import RC  # noqa: F401  (scripts/utils/bio.py:27-29)


def read_bio_seqs(filename):
    """fasta -> {id: sequence} (scripts/utils/bio.py:16-24; fasta only)."""
    seqs, name, chunks = {}, None, []
    with open(filename) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if name is not None:
                    seqs[name] = "".join(chunks)
                name, chunks = (line[1:].split() or [""])[0], []
            elif name is not None:
                chunks.append(line.strip())
    if name is not None:
        seqs[name] = "".join(chunks)
    return seqs


def read_bio_seq(filename):
    return str(list(read_bio_seqs(filename).values())[0])


def write_bio_seqs(filename, seqs):
    with open(filename, "w") as f:
        for seq_id, seq in seqs.items():
            f.write(f">{seq_id}\n{seq}\n")
