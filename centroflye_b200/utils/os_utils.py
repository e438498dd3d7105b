"""Filesystem helpers with the reference's names (scripts/utils/os_utils.py:21-34)."""
import os


def smart_mkdir(dirname):
    """``mkdir`` that tolerates an existing directory (parents must exist)."""
    try:
        os.mkdir(dirname)
    except FileExistsError:
        pass


def smart_makedirs(dirname):
    """``mkdir -p``: an existing directory is not an error."""
    os.makedirs(dirname, exist_ok=True)
