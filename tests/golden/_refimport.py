"""Import of the UNMODIFIED reference for the golden generator and the live cross-check tests:
delegates to oracle/refimport.py (staged copy in oracle/_ref/ first, then /root/reference/src)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle.refimport import REF_ROOT, REF_SRC, available, import_reference  # noqa: E402,F401
