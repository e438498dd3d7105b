"""The k-mer counting step of the reference's ``scripts/better_consensus_unit_reconstruction.py`` on the device
(SURVEY.md §8f rank 3).  Only ``get_kmer_counts_reads`` (:127-135) lives here -- the De Bruijn graph polishing that
consumes the counts stays with the reference.

``get_kmer_counts_reads(ncrf_report, k)`` returns a mapping kmer -> total number of occurrences over the gap-free rows of
all records; it reads like the reference's ``defaultdict(int)`` (missing k-mers count 0) and iterates in sorted order."""
from .distance_based_kmer_recruitment import KmerFreqs
from .encode import check_k
from .engine import U32_MAX, default_engine, to_host_u32, to_host_u64
from .read_kmer_cloud import report_batch, report_device_reads


def get_kmer_counts_reads(ncrf_report, k=19):
    k = check_k(k)
    engine = default_engine()
    batch = report_batch(ncrf_report)
    reads = report_device_reads(ncrf_report, engine, k)
    table = engine.count_total(reads, batch, k)
    keys, counts, _ = engine.table_select(table, 0, U32_MAX, U32_MAX, with_counts=True)
    return KmerFreqs(to_host_u64(keys), to_host_u32(counts), k, table=table, engine=engine)


def get_canonical_kmer_counts(ncrf_report, k=19):
    """kmer -> occurrences over the gap-free rows of all records with the two strands of a k-mer merged: the count of
    `jellyfish count -m k -C` that tandemQUAST's select_kmers.py:131-133 dumps for the reads, keyed by the canonical
    (smaller) strand."""
    k = check_k(k)
    engine = default_engine()
    batch = report_batch(ncrf_report)
    reads = report_device_reads(ncrf_report, engine, k)
    table = engine.count_total(reads, batch, k, canonical=True)
    keys, counts, _ = engine.table_select(table, 0, U32_MAX, U32_MAX, with_counts=True)
    return KmerFreqs(to_host_u64(keys), to_host_u32(counts), k, table=table, engine=engine)
