"""Minimal stand-in for Biopython so the UNMODIFIED reference modules import
(scripts/utils/bio.py:5 does ``from Bio import SeqIO``).  Test infrastructure only."""
