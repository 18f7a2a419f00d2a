"""Repository-level contracts that need no GPU: INTEGRATION.md lists every entry point include/mmseg_b200.h declares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "mmseg_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|int32_t|int64_t|const char\*)\s+(mmseg_[a-z0-9_]+)\s*\(", text, flags=re.M)))


def test_integration_doc_lists_every_entry_point():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    syms = _header_symbols()
    assert len(syms) >= 35
    families = {"mmseg_dicece_bwd": "mmseg_dicece_fwd/_bwd"}     # spelled as a family in the table
    missing = [s for s in syms if s not in doc and families.get(s, "\0") not in doc]
    assert not missing, missing
