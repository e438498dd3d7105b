"""B200-native unique-k-mer recruitment for centroFlye (one hot path, nothing else).

Drop-in modules (same names as the reference's scripts/):
    centroflye_b200.distance_based_kmer_recruitment
    centroflye_b200.read_kmer_cloud
    centroflye_b200.read
    centroflye_b200.ncrf_parser
The device work is in csrc/cfk.cu behind the C ABI of include/cfk.h; see DESIGN.md.
"""
__version__ = "0.1.0"
