"""Base / k-mer encodings shared by the host side of the recruitment path.

Layout decisions (see DESIGN.md "Data layout in HBM"):

* bases are 2-bit codes A=0 C=1 G=2 T=3, so the numeric order of a packed
  k-mer equals the lexicographic order Python's ``sorted`` gives the reference
  when it writes ``unique_kmers_*.txt`` (distance_based_kmer_recruitment.py:161-164);
* a k-mer is a u64 with its FIRST base in the most significant used bits
  (``sum(code[i] << 2*(k-1-i))``); k <= 31 so that ~0 is free as the empty-slot
  sentinel of the device hash tables;
* reads are packed 16 bases per little-endian u32 word (base j of a word at bits
  2j..2j+1) and every read starts on a 64-base (16-byte) boundary so device
  code can issue aligned 128-bit loads.

The reference treats every character as a legal k-mer symbol
(distance_based_kmer_recruitment.py:47-53 does not even upper-case).  A 2-bit
alphabet cannot represent that, so anything outside upper-case ACGT is rejected
loudly (``ValueError``) instead of being silently miscounted.
"""
import numpy as np

MAX_K = 31
READ_ALIGN_BASES = 64  # every read starts on a 16-byte boundary of the packed stream

_CODE_OF = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE_OF[_c] = _i
_ASCII_OF = np.frombuffer(b"ACGT", dtype=np.uint8)


def check_k(k):
    if not isinstance(k, (int, np.integer)) or k < 1 or k > MAX_K:
        raise ValueError(f"k must be in [1, {MAX_K}] for 64-bit packed k-mers, got {k!r}")
    return int(k)


def ascii_to_codes(seq):
    """str/bytes/uint8-array of upper-case ACGT -> uint8 codes 0..3 (ValueError otherwise)."""
    if isinstance(seq, str):
        seq = seq.encode("ascii", errors="replace")
    arr = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.asarray(seq, dtype=np.uint8)
    codes = _CODE_OF[arr]
    if codes.size and codes.max() == 255:
        bad = int(np.flatnonzero(codes == 255)[0])
        raise ValueError(
            f"non-ACGT symbol {chr(int(arr[bad]))!r} at offset {bad}: the 2-bit device path only accepts upper-case A/C/G/T"
        )
    return codes


def codes_to_ascii(codes):
    return _ASCII_OF[np.asarray(codes, dtype=np.uint8)].tobytes().decode("ascii")


def pack_codes(codes, out=None):
    """uint8 codes (len multiple of 16 after zero padding) -> u32 words, 16 bases per word."""
    codes = np.asarray(codes, dtype=np.uint8)
    n_words = (codes.size + 15) // 16
    if codes.size != n_words * 16:
        codes = np.concatenate([codes, np.zeros(n_words * 16 - codes.size, dtype=np.uint8)])
    q = codes.reshape(-1, 4)
    b = (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8)
    words = b.view("<u4")
    if out is not None:
        out[: words.size] = words
        return out
    return words


def unpack_codes(words, n):
    b = np.asarray(words, dtype="<u4").view(np.uint8)
    q = np.empty((b.size, 4), dtype=np.uint8)
    q[:, 0] = b & 3
    q[:, 1] = (b >> 2) & 3
    q[:, 2] = (b >> 4) & 3
    q[:, 3] = (b >> 6) & 3
    return q.reshape(-1)[:n]


def kmer_to_int(kmer):
    v = 0
    for c in ascii_to_codes(kmer):
        v = (v << 2) | int(c)
    return v


def kmers_to_ints(kmers, k):
    """iterable of equal-length ACGT strings -> uint64 array (vectorised)."""
    kmers = list(kmers)
    if not kmers:
        return np.empty(0, dtype=np.uint64)
    joined = "".join(kmers)
    if len(joined) != k * len(kmers):
        raise ValueError(f"all k-mers must have length k={k}")
    codes = ascii_to_codes(joined).reshape(len(kmers), k).astype(np.uint64)
    shifts = (2 * (k - 1 - np.arange(k, dtype=np.uint64))).astype(np.uint64)
    return (codes << shifts).sum(axis=1, dtype=np.uint64)


def ints_to_kmers(vals, k):
    """uint64 array -> list[str] (vectorised)."""
    vals = np.asarray(vals, dtype=np.uint64)
    if vals.size == 0:
        return []
    shifts = (2 * (k - 1 - np.arange(k, dtype=np.uint64))).astype(np.uint64)
    codes = ((vals[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8)
    flat = _ASCII_OF[codes].tobytes().decode("ascii")
    return [flat[i * k:(i + 1) * k] for i in range(vals.size)]


def kmers_of_codes(codes, k):
    """All k-mers of one code array as uint64 (host helper for tests / small inputs)."""
    codes = np.asarray(codes, dtype=np.uint64)
    n = codes.size - k + 1
    if n <= 0:
        return np.empty(0, dtype=np.uint64)
    v = np.zeros(n, dtype=np.uint64)
    for i in range(k):
        v = (v << np.uint64(2)) | codes[i:i + n]
    return v
